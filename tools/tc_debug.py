import ctypes, os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = {'bn32': (1, 16, 16, 32, 32), 'bn64': (2, 17, 23, 64, 64), 'bn128': (1, 32, 40, 128, 256), 'bn32k': (1, 12, 16, 512, 32)}
if len(sys.argv) > 1:
    import torch, torch.nn.functional as F
    from gen6d_b200 import ops, _lib
    B, H, W, cin, cout = CASES[sys.argv[1]]
    k = 15 if sys.argv[1] == 'bn32k' else 3
    x = torch.randn(B, cin, H, W); w = torch.randn(cout, cin, k, k) * 0.05
    ref = F.conv2d(x, w, padding=k // 2)
    pc = ops.pack_conv(w.cuda(), None, pad=k // 2)
    print(sys.argv[1], 'launching', flush=True)
    y = ops.conv(x.permute(0, 2, 3, 1).contiguous().cuda(), pc)
    dbg = (ctypes.c_int * 8)()
    rc = _lib.lib().g6d_conv_tc_debug(dbg)
    print(sys.argv[1], 'rc', rc, 'timeout record', list(dbg), flush=True)
    err = (y.permute(0, 3, 1, 2).cpu() - ref).abs().max() / ref.abs().max()
    print(sys.argv[1], 'rel err', float(err), flush=True)
else:
    for name in CASES:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=90)
            print(r.stdout[-600:], r.stderr[-600:], f'[{time.time()-t:.1f}s]', flush=True)
        except subprocess.TimeoutExpired as e:
            print(name, 'TIMEOUT', (e.stdout or b'')[-300:], flush=True)
