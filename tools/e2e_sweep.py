"""End-to-end throughput of Gen6DEstimator.predict_many over (workers, batch) settings (numpy frames in,
numpy poses out, host geometry + H2D + D2H inside the timed region)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import synthetic as syn
est, db = syn.build_estimator()
ids = db.get_img_ids()
imgs = [db.get_image(ids[(7 + 3 * i) % len(ids)]) for i in range(8)]
K = db.K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
import itertools
for (workers, batch), si in itertools.product(((1, 4), (2, 4), (3, 4), (2, 8), (3, 8)), (2e-4,)):
    frames = [imgs[i % len(imgs)] for i in range(N)]
    est.predict_many(frames[:workers * batch * 2], [K] * (workers * batch * 2), workers=workers, batch=batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    est.predict_many(frames, [K] * N, workers=workers, batch=batch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'workers {workers} batch {batch} switchinterval {si:g}: {N / dt:7.1f} poses/s ({dt / N * 1e3:.2f} ms/pose)', flush=True)
