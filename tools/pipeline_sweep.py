import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import synthetic as syn
est, db = syn.build_estimator()
ids = db.get_img_ids(); K = db.K
imgs = [db.get_image(ids[(7 + 3 * i) % len(ids)]) for i in range(8)]
n = 40
frames = [imgs[i % 8] for i in range(n)]
est.predict(frames[0], K)
t = time.perf_counter()
for f in frames[:20]: est.predict(f, K)
print('single', 20 / (time.perf_counter() - t), 'poses/s')
for sw in (0.005, 0.0005):
    sys.setswitchinterval(sw)
    for w in (2, 3, 4, 6):
        est.predict_many(frames[:w * 2], [K] * (w * 2), workers=w)
        best = 0
        for rep in range(2):
            t = time.perf_counter()
            est.predict_many(frames, [K] * n, workers=w)
            torch.cuda.synchronize()
            best = max(best, n / (time.perf_counter() - t))
        print(f'switch {sw} workers {w}: {best:.1f} poses/s')
