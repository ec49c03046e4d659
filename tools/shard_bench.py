"""Throughput of the two shardings of SURVEY 8(e) (run under torchrun for N > 1):
  selector ref-shard  : BASELINE configs[3] shape: 36 rotation bins, --refs references in total
  refiner pose-shard  : BASELINE configs[4] shape: --poses poses in total, 32^3 volume
Prints one JSON line per workload from rank 0 (device time, max over ranks)."""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from golden import cases
from gen6d_b200 import dist as gdist, ops
from gen6d_b200.network import name2network
from gen6d_b200.weights import seeded_state_dict

ap = argparse.ArgumentParser()
ap.add_argument('--refs', type=int, default=64)
ap.add_argument('--bins', type=int, default=36)
ap.add_argument('--poses', type=int, default=32)
ap.add_argument('--iters', type=int, default=5)
a = ap.parse_args()
comm = gdist.init_from_env()
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))


def build(name, cfg):
    net = name2network[name](cfg)
    net.load_state_dict(seeded_state_dict(net, 0))
    return net.cuda().eval()


def timed(fn, iters):
    fn(); comm.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda', dtype=torch.float64)
    return float(comm.all_reduce_sum(t)[0]) / comm.world if False else float(t[0])


# ---- selector, references sharded
g = torch.Generator().manual_seed(1)
sel = gdist.shard_selector(build('selector', {'selector_angle_num': a.bins}), comm)
r0, r1 = comm.shard_range(a.refs)
poses = cases.sphere_poses(3, a.refs)
# synthesise only this rank's reference images (the full set would be refs*bins*49 KB of uint8 per rank)
imgs = (torch.rand(a.bins, a.refs if comm.world == 1 else a.refs, 8, 8, 3, generator=g) * 255).to(torch.uint8).numpy()
imgs = np.kron(imgs, np.ones((1, 1, 16, 16, 1), np.uint8))
sel.load_ref_imgs(imgs, poses, np.zeros(3, np.float32), np.array([0, 0, 1], np.float32))
que = cases.rand_images_u8(5, 1, 128, 128, 3)
ms = timed(lambda: sel.select_que_imgs(que), a.iters)
if comm.rank == 0:
    print(json.dumps({'workload': f'selector ref-shard: {a.refs} refs x {a.bins} bins over {comm.world} GPU(s)',
                      'ms_per_query': ms, 'queries_per_s': 1e3 / ms, 'slices_per_gpu': (r1 - r0) * a.bins,
                      'ref_stack_gb_per_gpu': (r1 - r0) * a.bins * 688128 / 1e9}))

# ---- refiner, poses sharded
rfr = build('refiner', {})
rc = cases.refiner_case(seed=7, qn=a.poses)
dev = lambda x: torch.from_numpy(x).cuda()
args = [ops.preprocess_u8(dev(rc['que_imgs']), 4, True), dev(rc['que_Ks']), dev(rc['que_poses']),
        ops.preprocess_u8(dev(rc['ref_imgs']), 4, True), dev(rc['ref_Ks']), dev(rc['ref_poses'])]
ms = timed(lambda: gdist.pose_shard(rfr._forward_nhwc, args, comm), a.iters)
if comm.rank == 0:
    print(json.dumps({'workload': f'refiner pose-shard: {a.poses} poses, 6 views, 32^3, over {comm.world} GPU(s)',
                      'ms_per_iteration': ms, 'pose_iterations_per_s': a.poses / ms * 1e3}))
