import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import ops
os.environ['G6D_CONV_FLAT'] = '0'
H, W, cin, cout = int(sys.argv[1]), int(sys.argv[2]), 512, 512
w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.02
pc = ops.pack_conv(w, torch.zeros(cout, device='cuda'), pad=1)
x = torch.randn(1, H, W, cin, device='cuda')
for _ in range(3): ops.conv(x, pc)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv(x, pc)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
