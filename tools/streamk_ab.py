"""A/B of the persistent convolution kernel's two schedules (tile schedule + uniform split-K vs stream-K) on the
layer shapes of a pose step that take the persistent kernel.  Each shape: a CUDA graph of REP launches, timed over
replays with CUDA events (no launch overhead in the numbers)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import ops

ops.require_cuda()
REP = 10
# (B, D, H, W, cin, cout, k, pro)   k = (kd, kh, kw)
SHAPES = [(320, 1, 8, 8, 512, 128, (1, 3, 3), 1), (320, 1, 8, 8, 128, 128, (1, 3, 3), 1), (320, 1, 8, 8, 64, 128, (1, 3, 3), 1),
          (320, 1, 4, 4, 256, 256, (1, 3, 3), 1), (320, 1, 4, 4, 512, 256, (1, 3, 3), 1), (320, 1, 4, 4, 128, 256, (1, 3, 3), 1),
          (1280, 1, 8, 8, 512, 128, (1, 3, 3), 1), (1280, 1, 4, 4, 256, 256, (1, 3, 3), 1),
          (64, 1, 1, 1, 512, 512, (1, 1, 1), 0), (20, 1, 4, 4, 768, 512, (1, 1, 1), 0), (4, 1, 4, 4, 512, 512, (1, 3, 3), 0)]


def time_graph(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(REP):
                fn()
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20):
            gr.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (20 * REP) * 1e3


for B, D, H, W, cin, cout, k, pro in SHAPES:
    shape = (B, D, H, W, cin) if k[0] > 1 or D > 1 else (B, H, W, cin)
    x = torch.randn(*shape, device='cuda')
    wshape = (cout, cin) + (k if len(shape) == 5 else k[1:])
    w = torch.randn(*wshape, device='cuda') * 0.02
    pc = ops.pack_conv(w, None, pad=tuple(kk // 2 for kk in (k if len(shape) == 5 else k[1:])))
    kw = {}
    if pro:
        ps = torch.rand(B, cin, device='cuda') + 0.5
        pb = torch.randn(B, cin, device='cuda') * 0.1
        kw = dict(prologue=ops.PRO_AFFINE_RELU, pro_scale=ps, pro_shift=pb, group_rows=1)
    out = []
    for sk in ('0', '1'):
        os.environ['G6D_CONV_STREAMK'] = sk
        out.append(time_graph(lambda: ops.conv(x, pc, **kw)))
    M = B * D * H * W
    K = cin * k[0] * k[1] * k[2]
    fl = 2.0 * M * K * cout
    print(f'M={M:6d} N={cout:3d} K={K:5d} pro={pro}  tile-schedule {out[0]:7.1f} us ({fl / out[0] / 1e6:6.1f} TF/s)   '
          f'stream-K {out[1]:7.1f} us ({fl / out[1] / 1e6:6.1f} TF/s)   x{out[0] / out[1]:.2f}', flush=True)
os.environ.pop('G6D_CONV_STREAMK', None)
