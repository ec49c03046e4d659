"""Where the host time of one predict_batch call goes (cProfile, one thread, graphs already captured):
`cpu` of Tensor is the wait for the stage's kernels + D2H, everything else is host work that a single lane cannot
overlap with its own device work."""
import cProfile, os, pstats, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
est, db = syn.build_estimator()
ids = db.get_img_ids()
imgs = [db.get_image(ids[(7 + 3 * i) % len(ids)]) for i in range(B)]
Ks = [db.K] * B
for _ in range(3):
    est.predict_batch(imgs, Ks)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    est.predict_batch(imgs, Ks)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(22)
st.sort_stats('cumulative').print_stats(30)
