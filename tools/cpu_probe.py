import os, time, torch
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, 'n/a')
x = torch.rand(8, 256, 64, 64); w = torch.rand(256, 256, 3, 3)
for n in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    torch.nn.functional.conv2d(x, w, padding=1)
    t = time.perf_counter()
    for _ in range(3): torch.nn.functional.conv2d(x, w, padding=1)
    print(n, 'threads', (time.perf_counter() - t) / 3 * 1e3, 'ms')
