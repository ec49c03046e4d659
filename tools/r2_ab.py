"""A/B of the refiner volume-fill kernels (G6D_R2_V = 3: two-pass, every tap loaded; 4 / 5: tap footprints reused
along the k-runs, branchy / predicated) on a real refinement problem, 1 and 10 poses per launch: CUDA events around
each launch (ops profiling), and the largest deviation of the outputs from version 3."""
import os, sys
os.environ['G6D_BRANCH_STREAMS'] = '0'        # serialised: nothing overlaps the timed kernel
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import geometry as G, ops, synthetic as syn
est, db = syn.build_estimator()
ids = db.get_img_ids(); K = db.K
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for Q in (1, 10):
    probs = [G.refine_problem(db, ids, db.get_image(ids[(7 + 5 * q) % len(ids)]), K, db.get_pose(ids[(9 + 5 * q) % len(ids)]), 128, 6, True)
             for q in range(Q)]
    args = [dev(np.stack([p[k] for p in probs], 0)) for k in ('que_img', 'que_K', 'que_pose', 'ref_imgs', 'ref_Ks', 'ref_poses')]
    ref_out = None
    for ver in ('3', '4', '5'):
        os.environ['G6D_R2_V'] = ver
        with torch.no_grad():
            for _ in range(3):
                est.refiner._refine_u8(*args)
            torch.cuda.synchronize()
            prof = ops.enable_profiling()
            for _ in range(10):
                out = est.refiner._refine_u8(*args)
            st = ops.collect_profile(prof)
        r2 = st['g6d_ref_volume_fill']
        us = r2['ms'] / r2['n'] * 1e3
        dev_out = 0.0 if ref_out is None else float((out - ref_out).abs().max())
        if ref_out is None:
            ref_out = out.clone()
        print(f'Q={Q:2d} G6D_R2_V={ver}: {us:7.1f} us per launch, {us / Q:6.1f} us per pose, {r2["work"] / r2["n"] / us / 1e3:7.1f} GB/s, '
              f'max |d pose update| vs v3 {dev_out:.2e}', flush=True)
os.environ.pop('G6D_R2_V', None)
