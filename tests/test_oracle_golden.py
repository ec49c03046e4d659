"""Pins oracle/gen6d_oracle.py (the CPU restatement) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU-only; runs in `-m "not gpu"`."""
import json
import os

import numpy as np
import pytest
import torch

from golden import cases
from gen6d_b200.network import name2network
from gen6d_b200.weights import seeded_state_dict
from oracle import gen6d_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'net_golden.npz'))


def sub(t, n=4096):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return f[::step][:n].numpy()


def close(a, b, rtol=1e-4, atol=1e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_state_dict_spec_matches_reference():
    spec = json.load(open(os.path.join(HERE, 'golden', 'state_dict_spec.json')))
    for name, entries in spec.items():
        cfg = {}
        ours = name2network[name](cfg).state_dict()
        assert sorted(e[0] for e in entries) == sorted(ours.keys())  # order is irrelevant to load_state_dict
        for k, shape, dtype in entries:
            assert list(ours[k].shape) == shape, k
            assert str(ours[k].dtype) == dtype, k


@pytest.fixture(scope='module')
def det():
    c = cases.detector_case()
    sd = seeded_state_dict(name2network['detector'](c['cfg']), cases.WEIGHT_SEED)
    ref_feats = O.det_load_refs(sd, cases.u8_to_nchw(c['ref_imgs']))
    return c, sd, ref_feats


def test_detector_reference_features(det):
    c, sd, ref_feats = det
    for l, f in enumerate(ref_feats):
        close(sub(f), G[f'det.ref_feats{l}.sub'], rtol=1e-4, atol=1e-4)


def test_detector_raw_correlation(det):
    c, sd, ref_feats = det
    qf = O.det_extract(sd, cases.u8_to_nchw(c['que_imgs']))
    for l, raw in enumerate(O.det_raw_correlation(qf, ref_feats)):
        close(raw, G[f'det.raw_corr{l}'], rtol=1e-5, atol=1e-2)


def test_detector_maps_and_argmax(det):
    c, sd, ref_feats = det
    out = O.det_detect(sd, c['cfg'], cases.u8_to_nchw(c['que_imgs']), ref_feats)
    close(out['scores'], G['det.scores'], atol=1e-4)
    close(out['select_pr_offset'], G['det.offset'], atol=1e-4)
    close(out['select_pr_scale'], G['det.scale'], atol=1e-4)
    pos, scl, idx = O.det_parse(out['scores'], out['select_pr_scale'], out['select_pr_offset'])
    assert idx.numpy().tolist() == G['det.argmax'].tolist()  # bit-exact cell selection
    close(pos, G['det.positions'], atol=1e-3)
    close(scl, G['det.scales'], atol=1e-4)
    close(pos, G['det.wrap.positions'], atol=1e-3)


@pytest.fixture(scope='module')
def sel():
    c = cases.selector_case()
    sd = seeded_state_dict(name2network['selector'](c['cfg']), cases.WEIGHT_SEED)
    feats, embed = O.sel_load_refs(sd, cases.u8_to_nchw(c['ref_imgs']), torch.from_numpy(c['ref_poses']),
                                   torch.from_numpy(c['object_center']), torch.from_numpy(c['object_vert']))
    return c, sd, feats, embed


def test_selector_reference_cache(sel):
    c, sd, feats, embed = sel
    for l, f in enumerate(feats):
        close(sub(f), G[f'sel.ref_feats{l}.sub'], atol=1e-6)
    close(embed, G['sel.ref_pose_embed'], atol=1e-5)


def test_selector_score_vps(sel):
    c, sd, feats, embed = sel
    que = O.sel_feats(sd, cases.u8_to_nchw(c['que_imgs']))
    close(O.sel_score_vps(que, feats), G['sel.score_vps'], rtol=1e-4, atol=1e-5)


def test_selector_logits_angles_argmax(sel):
    c, sd, feats, embed = sel
    logits, angles = O.sel_forward(sd, cases.u8_to_nchw(c['que_imgs']), feats, embed)
    close(logits, G['sel.logits'], atol=2e-4)
    close(angles, G['sel.angles'], atol=2e-4)
    idx, ang = O.sel_select(logits, angles)
    assert idx.numpy().tolist() == G['sel.wrap.ref_idx'].tolist()  # bit-exact viewpoint selection
    close(ang, G['sel.wrap.angles'], atol=2e-4)


def test_refiner_volume_and_regression():
    c = cases.refiner_case()
    sd = seeded_state_dict(name2network['refiner']({}), cases.WEIGHT_SEED)
    T = torch.from_numpy
    out = O.ref_forward(sd, cases.u8_to_nchw(c['que_imgs']), T(c['que_Ks']), T(c['que_poses']),
                        cases.u8_to_nchw(c['ref_imgs']), T(c['ref_Ks']), T(c['ref_poses']), 32, return_taps=True)
    close(sub(out['mean'], 16384), G['ref.mean.sub'], atol=1e-5)
    close(sub(out['std'], 16384), G['ref.std.sub'], atol=1e-5)
    close(sub(out['vin'], 16384), G['ref.in.sub'], atol=1e-5)
    close(out['rotation'], G['ref.rotation'], atol=1e-4)
    close(out['offset'], G['ref.offset'], atol=1e-4)
    close(out['scale'], G['ref.scale'], atol=1e-4)
    qf = O.ref_feature_net(sd, cases.u8_to_nchw(c['que_imgs']))
    close(sub(qf, 8192), G['ref.feat_que.sub'], atol=1e-5)
