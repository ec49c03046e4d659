"""Where does an end-to-end predict() spend its time?  (host geometry vs device stages)"""
import os, sys, time
import numpy as np, torch, cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import geometry as G, synthetic as syn
print('cv2 threads', cv2.getNumThreads(), 'torch threads', torch.get_num_threads(), 'cpus', len(os.sched_getaffinity(0)))
if len(sys.argv) > 1:
    cv2.setNumThreads(int(sys.argv[1])); torch.set_num_threads(int(sys.argv[1]))
est, db = syn.build_estimator()
ids = db.get_img_ids(); K = db.K
img = db.get_image(ids[3])
def T(fn, n=10):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
pose, inter = est.predict(img, K)
print('predict ms', T(lambda: est.predict(img, K)))
est.cfg['host_warps'] = True
print('predict ms (host warps, as the reference)', T(lambda: est.predict(img, K)))
est.cfg['host_warps'] = False
frame = est.detector.upload_frame(img)
print(' upload_frame ms', T(lambda: est.detector.upload_frame(img)))
print(' detect_que_imgs ms', T(lambda: est.detector.detect_que_imgs(img[None], que_dev=frame[None])))
_, M = G.crop_similarity(None, inter['det_position'], 1 / inter['det_scale_r2q'], 0, 128)
print(' select_from_frame ms', T(lambda: est.selector.select_from_frame(frame, M, 128)))
p0 = inter['refine_poses'][0]
print(' refine_que_imgs ms (device warps)', T(lambda: est.refiner.refine_que_imgs(img, K, p0, 128, 6, True, que_dev=frame)))
print(' refine_que_imgs ms (host warps)', T(lambda: est.refiner.refine_que_imgs(img, K, p0, 128, 6, True, host_warps=True)))
print('   refine_problem ms (host warps)', T(lambda: G.refine_problem(db, ids, img, K, p0, 128, 6, True)))
print('   refine_problem ms (no warps)', T(lambda: G.refine_problem(db, ids, None, K, p0, 128, 6, True, warp=False)))
pr = G.refine_problem(db, ids, None, K, p0, 128, 6, True, warp=False)
def jobs():
    srcs = [frame] + est.refiner._ref_images_dev(list(pr['ref_ids']))
    mats = [G.perspective_dst_to_src(pr['que_H'])] + [G.perspective_dst_to_src(H) for H in pr['ref_Hs']]
    return G.pack_warp_jobs(srcs, mats)
print('   pack_warp_jobs ms', T(jobs))
args = [est.refiner._to_dev(jobs())] + [est.refiner._to_dev(pr[k][None]) for k in ('que_K', 'que_pose', 'ref_Ks', 'ref_poses')]
print('   refine graph ms', T(lambda: est.refiner.stages.run('refine_warp128', est.refiner._refine_warped(128), args)))
print('   apply_refinement ms', T(lambda: G.apply_refinement(pr, np.array([1., 0, 0, 0]), np.array([0.1, 0.1]), 1.01)))
