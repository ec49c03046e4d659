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
print(' detect_que_imgs ms', T(lambda: est.detector.detect_que_imgs(img[None])))
crop = inter['det_que_img']
print(' crop_similarity ms', T(lambda: G.crop_similarity(img, inter['det_position'], 1 / inter['det_scale_r2q'], 0, 128)))
print(' select_que_imgs ms', T(lambda: est.selector.select_que_imgs(crop[None])))
p0 = inter['refine_poses'][0]
print(' refine_que_imgs ms', T(lambda: est.refiner.refine_que_imgs(img, K, p0, 128, 6, True)))
print('   refine_problem ms', T(lambda: G.refine_problem(db, ids, img, K, p0, 128, 6, True)))
prob = G.refine_problem(db, ids, img, K, p0, 128, 6, True)
args = [est.refiner._to_dev(prob[k][None]) for k in ('que_img', 'que_K', 'que_pose', 'ref_imgs', 'ref_Ks', 'ref_poses')]
print('   _to_dev x6 ms', T(lambda: [est.refiner._to_dev(prob[k][None]) for k in ('que_img', 'que_K', 'que_pose', 'ref_imgs', 'ref_Ks', 'ref_poses')]))
print('   refine graph ms', T(lambda: est.refiner.stages.run('refine', est.refiner._refine_u8, args)))
print('   refine eager ms', T(lambda: est.refiner._refine_u8(*args)))
print('   apply_refinement ms', T(lambda: G.apply_refinement(prob, np.array([1., 0, 0, 0]), np.array([0.1, 0.1]), 1.01)))
u8 = est.detector._to_dev(img[None])
print(' detect graph ms', T(lambda: est.detector.stages.run('detect', est.detector._detect_u8, [u8])))
print(' detect eager ms', T(lambda: est.detector._detect_u8(u8)))
c8 = est.selector._to_dev(crop[None])
print(' select graph ms', T(lambda: est.selector.stages.run('select', est.selector._select_u8, [c8])))
print(' select eager ms', T(lambda: est.selector._select_u8(c8)))
