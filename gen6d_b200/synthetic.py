"""The synthetic workload used by bench.py, smoke() and the estimator-level tests: seeded
random-weight checkpoints (BASELINE.json: "random weights", no network for pretrained ones), the
procedural object database, and detector score statistics matched to those weights so the
correlation maps are not saturated by the clip (SURVEY.md 8c caution (i))."""
import numpy as np

from .database import SyntheticObjectDatabase
from .network import name2network
from .weights import seeded_state_dict

WEIGHT_SEED = 0
# per-level (mean, std) of the raw detector correlation under WEIGHT_SEED weights
DET_SCORE_STATS = [[106600.0, 29270.0], [68870.0, 23050.0], [18050.0, 5970.0]]
DATABASE = {'n_views': 72, 'height': 480, 'width': 640, 'seed': 7}


def network_cfgs(angle_num=5):
    return {
        'detector': {'name': 'detector_synth', 'network': 'detector', 'vgg_score_stats': DET_SCORE_STATS},
        'selector': {'name': 'selector_synth', 'network': 'selector', 'selector_angle_num': angle_num},
        'refiner': {'name': 'refiner_synth', 'network': 'refiner'},
    }


def seeded_networks(device='cuda', angle_num=5, seed=WEIGHT_SEED):
    nets = {}
    for name, cfg in network_cfgs(angle_num).items():
        net = name2network[name](cfg)
        net.load_state_dict(seeded_state_dict(net, seed), strict=True)
        nets[name] = net.to(device).eval() if device != 'cpu' else net.eval()
    return nets


def seeded_state_dicts(angle_num=5, seed=WEIGHT_SEED):
    return {name: seeded_state_dict(name2network[name](cfg), seed) for name, cfg in network_cfgs(angle_num).items()}


def synthetic_database(**over):
    return SyntheticObjectDatabase(**{**DATABASE, **over})


def build_estimator(database=None, **cfg_over):
    from .estimator import Gen6DEstimator
    est = Gen6DEstimator({'refine_iter': 3, **cfg_over}, modules=seeded_networks())
    database = database or synthetic_database()
    est.build(database, 'all')
    return est, database
