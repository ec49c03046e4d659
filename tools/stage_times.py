"""Per-stage device timings of the three networks at the estimator's default sizes
(detector 480x640 frame / 32 refs, selector 64 refs x 5 angles, refiner 6 refs / 32^3).
Development aid; bench.py is the contract benchmark."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from golden import cases  # noqa: E402
from gen6d_b200 import _lib, ops  # noqa: E402
from gen6d_b200.network import name2network  # noqa: E402
from gen6d_b200.weights import seeded_state_dict  # noqa: E402


def timed(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (_lib.launch_count() - l0) // iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--det-refs', type=int, default=32)
    ap.add_argument('--sel-refs', type=int, default=64)
    ap.add_argument('--an', type=int, default=5)
    ap.add_argument('--qn', type=int, default=1)
    a = ap.parse_args()
    out = {}

    def build(name, cfg):
        net = name2network[name](cfg)
        net.load_state_dict(seeded_state_dict(net, 0))
        return net.cuda().eval()

    det = build('detector', {'vgg_score_stats': cases.DET_STATS})
    det.load_ref_imgs(cases.rand_images_u8(1, a.det_refs, 128, 128, 3))
    frame = cases.rand_images_u8(2, 1, 480, 640, 3)
    out['detector_ms'], out['detector_launches'] = timed(lambda: det.detect_que_imgs(frame))

    sel = build('selector', {'selector_angle_num': a.an})
    c = cases.selector_case(rfn=a.sel_refs, an=a.an)
    t0 = time.time()
    sel.load_ref_imgs(c['ref_imgs'], c['ref_poses'], c['object_center'], c['object_vert'])
    torch.cuda.synchronize()
    out['selector_load_s'] = time.time() - t0
    out['selector_ms'], out['selector_launches'] = timed(lambda: sel.select_que_imgs(c['que_imgs']))
    # the S2 kernel alone, per level
    S = a.sel_refs * a.an
    for l, ref in enumerate(sel.ref_feats_cache):
        P = ref.shape[1] * ref.shape[2]
        q = torch.rand(P, 512, device='cuda')
        ms, _ = timed(lambda: ops.sel_corr_score(ref.reshape(S, P, 512), q), iters=20)
        out[f'sel_corr_score_l{l}_us'] = ms * 1e3
        out[f'sel_corr_score_l{l}_GBs'] = (S * P * 512 * 4 + P * 512 * 4 + S * 4) / (ms * 1e-3) / 1e9

    rfr = build('refiner', {})
    rc = cases.refiner_case(qn=a.qn)
    dev = lambda x: torch.from_numpy(x).cuda()
    que = ops.preprocess_u8(dev(rc['que_imgs']), 4, True)
    ref = ops.preprocess_u8(dev(rc['ref_imgs']), 4, True)
    args = (que, dev(rc['que_Ks']), dev(rc['que_poses']), ref, dev(rc['ref_Ks']), dev(rc['ref_poses']))
    out['refiner_ms'], out['refiner_launches'] = timed(lambda: rfr._forward_nhwc(*args))
    feats = torch.rand(a.qn * 7, 32, 32, 128, device='cuda')
    rf, qf = feats[:a.qn * 6].reshape(a.qn, 6, 32, 32, 128), feats[a.qn * 6:]
    ms, _ = timed(lambda: ops.ref_volume_fill(rf, qf, args[4], args[5], args[1], args[2], 32, 128, 128), iters=20)
    out['ref_volume_fill_us'] = ms * 1e3
    out['ref_volume_fill_GBs'] = a.qn * (7 * 128 * 32 * 32 * 4 + 3 * 128 * 32 ** 3 * 4) / (ms * 1e-3) / 1e9
    out['pose_ms_est'] = out['detector_ms'] + out['selector_ms'] + 3 * out['refiner_ms']
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
