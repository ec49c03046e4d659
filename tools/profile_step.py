"""One device-resident pose step between cudaProfilerStart/Stop, for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:<kernel> -c 3 -o gpurun_out/prof_<kernel> python tools/profile_step.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import geometry as G, ops, synthetic as syn  # noqa: E402

est, db = syn.build_estimator()
ids = db.get_img_ids()
K = db.K
pose0, inter = est.predict(db.get_image(ids[7]), K)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
frame_dev, crop_dev = dev(db.get_image(ids[7])[None]), dev(inter['det_que_img'][None])
gt = db.get_pose(ids[7])
pr = G.refine_problem(db, ids, db.get_image(ids[7]), K, gt, 128, 6, True)
prob = tuple(dev(pr[k][None]) for k in ('que_img', 'que_K', 'que_pose', 'ref_imgs', 'ref_Ks', 'ref_poses'))


def step(n_refine=3):
    with torch.no_grad():
        o = est.detector._detect_nhwc(ops.preprocess_u8(frame_dev, out_c=3, imagenet_norm=False))
        ops.det_parse(o['score_predict'], o['scale_predict'], o['offset_predict'], 8)
        lg, ang, _ = est.selector._select_nhwc(ops.preprocess_u8(crop_dev, out_c=4, imagenet_norm=True))
        ops.sel_parse(lg, ang)
        qi, qk, qp, ri, rk, rp = prob
        for _ in range(n_refine):
            est.refiner._forward_nhwc(ops.preprocess_u8(qi, 4, True), qk, qp, ops.preprocess_u8(ri, 4, True), rk, rp)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(n_refine=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled one step (detector + selector + 1 refinement iteration)')
