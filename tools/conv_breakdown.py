"""Per-call time / TFLOP/s of every tensor-core convolution in one pose step (eager, CUDA events)."""
import os, sys, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import geometry as G, ops, synthetic as syn
est, db = syn.build_estimator()
ids = db.get_img_ids(); K = db.K
img = db.get_image(ids[7])
pose0, inter = est.predict(img, K)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
frame, crop = dev(img[None]), dev(inter['det_que_img'][None])
pr = G.refine_problem(db, ids, img, K, inter['refine_poses'][0], 128, 6, True)
prob = [dev(pr[k][None]) for k in ('que_img', 'que_K', 'que_pose', 'ref_imgs', 'ref_Ks', 'ref_poses')]
def step():
    with torch.no_grad():
        est.detector._detect_u8(frame); est.selector._select_u8(crop); est.refiner._refine_u8(*prob)
step(); torch.cuda.synchronize()
prof = ops.enable_profiling(); step(); st = ops.collect_profile(prof)
calls = st['#calls']
tot = sum(c[0] for c in calls)
print(f'conv_tc calls {len(calls)} total {tot:.2f} ms, {sum(c[1] for c in calls)/tot/1e9:.1f} TFLOP/s')
agg = collections.OrderedDict()
for ms, work, name, tag in calls:
    a = agg.setdefault(tag, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += work
for tag, (n, ms, work) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'{ms:7.3f} ms x{n:2d} {work/ms/1e9:6.1f} TF/s  {tag}')
