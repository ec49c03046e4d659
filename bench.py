"""Contract benchmark: poses/sec of the Gen6D inference hot path (detect -> select -> 3x refine).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pose: a synthetic 480x640 frame through the detector (32 reference views,
4 scales), the 128x128 crop through the selector (64 reference views x 5 in-plane angles), and
three refinement iterations (6 views, 32^3 volume) -- BASELINE.json's full-estimator config.

  value : poses/s of the three-network device path with every input already resident in HBM
          (frame, crop and the three refinement problems were uploaded before the timed region).
  e2e   : poses/s through the public API `Gen6DEstimator.predict(numpy frame, K) -> numpy pose`,
          host geometry (OpenCV warps), pinned H2D copies and D2H reads inside the timed region.
  roofline     : dominant kernel (the implicit-GEMM convolution) timed live with CUDA events.
  cpu_baseline : oracle/ (torch-CPU port of the reference path) timed on the host cores.

N > 1 (torchrun): one process per GPU, each rank runs an independent replica on its own frames
(weak scaling, no data-path collective; poses are all-gathered once at the end over NCCL).
`--impl reference` times the CPU oracle port instead (the reference itself cannot travel to the
GPU box: /root/reference does not exist there).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E2E_WORKERS = int(os.environ.get('G6D_E2E_WORKERS', '6'))      # frames in flight per GPU (predict_many / device streams)
METRIC = 'poses/sec end-to-end (128^2 crop, 64 refs, 3 refine iters)'
WORKLOAD = ('full estimator detect->select->3x refine: synthetic 480x640 frame, detector 32 refs x 4 scales, '
            'selector 64 refs x 5 angles, refiner 6 views 32^3 volume, seeded random weights')


def read_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1400.0, 'src': 'fallback'}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------
def usable_cpus():
    from gen6d_b200.geometry import usable_cpus as u
    return u()


def cpu_pose_fn():
    """Returns (fn, describe): fn() runs ONE network-only pose of the oracle port on the CPU."""
    from gen6d_b200 import geometry as G
    from gen6d_b200 import synthetic as syn
    from oracle import gen6d_oracle as O
    torch.set_num_threads(min(usable_cpus(), 64))
    sds = syn.seeded_state_dicts()
    db = syn.synthetic_database()
    g = torch.Generator().manual_seed(5)
    to01 = lambda u8: torch.from_numpy(u8.astype(np.float32) / 255)
    det_refs = torch.rand(32, 3, 128, 128, generator=g)
    sel_refs = torch.rand(5, 64, 3, 128, 128, generator=g)
    ids = db.get_img_ids()[:64]
    poses = torch.from_numpy(np.stack([db.get_pose(i) for i in ids]))
    t0 = time.perf_counter()
    with torch.no_grad():
        det_feats = O.det_load_refs(sds['detector'], det_refs)
        sel_feats, embed = O.sel_load_refs(sds['selector'], sel_refs, poses, torch.zeros(3), torch.tensor([0., 0., 1.]))
    load_s = time.perf_counter() - t0
    frame = to01(db.get_image('11')).permute(2, 0, 1)[None].contiguous()
    crop = torch.rand(1, 3, 128, 128, generator=g)
    rq, rr = torch.rand(1, 3, 128, 128, generator=g), torch.rand(1, 6, 3, 128, 128, generator=g)
    K = torch.tensor([[[304., 0, 64], [0, 304., 64], [0, 0, 1]]])
    qp = torch.from_numpy(db.get_pose('11'))[None]
    rp = torch.from_numpy(np.stack([db.get_pose(i) for i in ids[:6]]))[None]
    det_cfg = {'vgg_score_stats': syn.DET_SCORE_STATS}

    def one_pose():
        with torch.no_grad():
            o = O.det_detect(sds['detector'], det_cfg, frame, det_feats)
            O.det_parse(o['scores'], o['select_pr_scale'], o['select_pr_offset'])
            lg, ang = O.sel_forward(sds['selector'], crop, sel_feats, embed)
            O.sel_select(lg, ang)
            for _ in range(3):
                O.ref_forward(sds['refiner'], rq, K, qp, rr, K[:, None].repeat(1, 6, 1, 1), rp, 32)

    return one_pose, {'reference_set_load_s': round(load_s, 2)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    fn, info = cpu_pose_fn()
    steps, warm = args.steps, args.warmup
    note = None
    if steps + warm > 40:   # keep the whole run within a few minutes (one CPU pose is ~4-5 s)
        steps = max(1, 40 - warm)
        note = f'steps capped from {args.steps} to {steps} (one CPU pose takes seconds)'
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    v = steps / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'poses/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warm, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'path': 'network-only pose (no host warps), torch CPU'},
            'cpu_baseline': {'value': v, 'unit': 'poses/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                             'sample': f'{steps} poses of the oracle port (torch {torch.__version__} CPU kernels)', **info},
            'e2e': {'value': v, 'unit': 'poses/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if note:
        line['note'] = note
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from gen6d_b200 import _lib, graphs, ops
    from gen6d_b200 import geometry as G
    from gen6d_b200 import synthetic as syn
    from gen6d_b200.network import base as nbase

    torch.cuda.set_device(local_rank)
    ops.require_cuda()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    est, db = syn.build_estimator()
    ids = db.get_img_ids()
    frames = [ids[(7 + rank * 13 + i * 3) % len(ids)] for i in range(8)]   # different frames per rank
    K = db.K

    # ---- stage inputs for the device-resident measurement (captured from one real prediction)
    pose0, inter = est.predict(db.get_image(frames[0]), K)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    frame_dev = dev(db.get_image(frames[0]))                     # uint8 [h,w,3], resident
    _, M_crop = G.crop_similarity(None, inter['det_position'], 1 / inter['det_scale_r2q'], 0, 128)
    crop_jobs = dev(G.pack_warp_jobs([frame_dev], [G.affine_dst_to_src(M_crop)]))
    probs = []
    for p in inter['refine_poses'][:3]:
        pr = G.refine_problem(db, ids, None, K, p, 128, 6, True, warp=False)
        srcs = [frame_dev] + est.refiner._ref_images_dev(list(pr['ref_ids']))
        mats = [G.perspective_dst_to_src(pr['que_H'])] + [G.perspective_dst_to_src(H) for H in pr['ref_Hs']]
        probs.append((dev(G.pack_warp_jobs(srcs, mats)),) + tuple(dev(pr[k][None]) for k in ('que_K', 'que_pose', 'ref_Ks', 'ref_poses')))
    det, sel, rfr = est.detector, est.selector, est.refiner

    # W independent frames in flight, each on its own stream with its own captured stage graphs
    # (shared weights / reference features): at batch 1 many kernels launch fewer CTAs than SMs,
    # so concurrent frames are what fills the machine.  Results stay on the device.
    W = E2E_WORKERS
    nets = [(det, sel, rfr)] + [(det.worker_clone(), sel.worker_clone(), rfr.worker_clone()) for _ in range(W - 1)]
    lanes = [torch.cuda.Stream() for _ in range(W)]

    def device_step(i=0, eager=False):
        """The three-network path on device-resident inputs (through the captured stage graphs,
        exactly what predict() launches, minus host geometry and copies)."""
        d, sl, r = nets[0] if eager else nets[i % W]
        run = (lambda m, name, fn, a: fn(*a)) if eager else (lambda m, name, fn, a: m.stages.run(name, fn, a))
        with torch.no_grad():
            run(d, 'detect', d._detect_u8, [frame_dev[None]])
            run(sl, 'select_warp128', sl._select_warped(128), [crop_jobs])          # detection crop + selector
            for pr in probs:
                run(r, 'refine_warp128', r._refine_warped(128), list(pr))            # 7 look-at crops + refiner

    def device_steps(n):
        main = torch.cuda.current_stream()
        for st in lanes:
            st.wait_stream(main)
        for i in range(n):
            with torch.cuda.stream(lanes[i % W]):
                device_step(i)
        for st in lanes:
            main.wait_stream(st)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm, takes_index=False, batched=False):
        if batched:
            fn(warm)
        else:
            for _ in range(warm):
                fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count() + graphs.REPLAYED_KERNELS[0]
        w0 = time.perf_counter()
        e0.record()
        if batched:
            fn(steps)
        else:
            for i in range(steps):
                fn(i) if takes_index else fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        ms = max(e0.elapsed_time(e1), 0.0)
        launches = _lib.launch_count() + graphs.REPLAYED_KERNELS[0] - l0
        barrier()
        t = torch.tensor([ms, wall * 1e3], device='cuda', dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, _, launches = timed(device_steps, args.steps, max(args.warmup, 2 * W), batched=True)

    # ---- end to end through the public API (numpy in, numpy out)
    imgs = [db.get_image(f) for f in frames]
    nbase.IO_BYTES['h2d'] = nbase.IO_BYTES['d2h'] = 0
    out_poses = []

    def e2e_step(i=0):
        pose, _ = est.predict(imgs[i % len(imgs)], K)
        out_poses.append(pose)

    _, e2e_wall_ms, _ = timed(e2e_step, args.steps, args.warmup, takes_index=True)
    io = dict(nbase.IO_BYTES)
    n_calls = args.steps + args.warmup

    # the throughput API: the same per-frame predict(), two frames in flight on one GPU
    def pipelined(n):
        res = est.predict_many([imgs[i % len(imgs)] for i in range(n)], [K] * n, workers=E2E_WORKERS)
        out_poses.extend(r[0] for r in res)

    pipelined(2 * E2E_WORKERS)                         # builds the worker clones, captures their graphs
    pipelined(max(args.warmup, E2E_WORKERS))           # untimed warm-up of the whole pipelined path
    barrier()
    t0 = time.perf_counter()
    pipelined(args.steps)
    torch.cuda.synchronize()
    pipe_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([pipe_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        pipe_ms = float(tt[0])
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        mine = torch.from_numpy(np.stack(out_poses[-args.steps:], 0)).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)       # the only collective: results, once, at the end

    # ---- live kernel timing (CUDA events around every launch of the three kernels of interest)
    os.environ['G6D_BRANCH_STREAMS'] = '0'      # per-kernel timing: one kernel at a time, no co-scheduling
    device_step(0, eager=True)
    torch.cuda.synchronize()
    prof = ops.enable_profiling()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(2):
        device_step(0, eager=True)
    pe1.record()
    torch.cuda.synchronize()
    eager_ms = pe0.elapsed_time(pe1)
    stats = ops.collect_profile(prof)
    os.environ.pop('G6D_BRANCH_STREAMS', None)
    peaks = read_peaks()
    conv = stats.get('g6d_conv_tc', {'ms': 0, 'work': 0, 'n': 1})
    ffma = stats.get('g6d_conv', {'ms': 0, 'work': 0, 'n': 0})
    roof = {'kernel': 'conv_tc_kernel (tcgen05 implicit-GEMM convolution, 3xTF32 fp32-faithful mode)', 'bound': 'tensor',
            'achieved': conv['work'] / max(conv['ms'], 1e-9) / 1e9, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
            'traffic': None,
            'peak_source': f"{peaks['src']} bf16 dense GEMM (sustained); achieved counts fp32-equivalent flops 2MNK, each "
                           "issued as 3 TF32 MMAs, so 1/6 of this peak is the ceiling of the parity mode",
            'launches_per_step': conv['n'] // 2, 'ms_per_step': conv['ms'] / 2,
            'share_of_step': conv['ms'] / max(eager_ms, 1e-9),
            'ffma_fallback': {'launches_per_step': ffma['n'] // 2, 'ms_per_step': ffma['ms'] / 2,
                              'tflops': ffma['work'] / max(ffma['ms'], 1e-9) / 1e9}}
    roof['frac'] = roof['achieved'] / roof['peak']
    roof['frac_of_3xtf32_ceiling'] = roof['achieved'] / (roof['peak'] / 6.0)
    extra = []
    for key, label in (('g6d_sel_corr_score3', 'selector correlation + rotated-similarity score, 3 levels (S2)'),
                       ('g6d_ref_volume_fill', 'refiner unproject-and-aggregate volume fill (R2)')):
        if key in stats:
            s = stats[key]
            ach = s['work'] / max(s['ms'], 1e-9) / 1e6
            extra.append({'kernel': label, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                          'frac': ach / peaks['hbm_gbs'], 'us_per_launch': s['ms'] / s['n'] * 1e3, 'traffic': None})

    if rank != 0:
        return
    value = world * args.steps / (dev_ms * 1e-3)
    e2e_v = world * args.steps / (e2e_wall_ms * 1e-3)
    line = {'metric': METRIC, 'value': value, 'unit': 'poses/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'parallelism': f'replica x{world} (independent frames per GPU), {E2E_WORKERS} frames in flight per GPU on separate streams',
                       'l2': 'per-step working set (220 MB selector reference stack + 300 MB weights + detector '
                             'activations) exceeds the 126 MB L2; no explicit flush'},
            'e2e': {'value': world * args.steps / (pipe_ms * 1e-3), 'unit': 'poses/s', 'ms_per_step': pipe_ms / args.steps,
                    'h2d_bytes_per_step': io['h2d'] // n_calls, 'd2h_bytes_per_step': io['d2h'] // n_calls,
                    'api': f'Gen6DEstimator.predict_many(numpy frames, Ks, workers={E2E_WORKERS}) -> numpy poses: per frame the same '
                           'predict() (pinned H2D of the frame once, crops cut from it on the device, camera geometry on the host, D2H of every stage '
                           f'result), {E2E_WORKERS} frames in flight per GPU',
                    'single_frame_latency': {'value': e2e_v, 'unit': 'poses/s', 'ms_per_step': e2e_wall_ms / args.steps,
                                             'api': 'Gen6DEstimator.predict(numpy frame, K), one frame at a time'}},
            'gpu_launches': int(launches), 'roofline': roof, 'kernels': extra, 'clocks': clocks}
    if world == 1:
        fn, info = cpu_pose_fn()        # sets torch threads to the usable-CPU count for the CPU baseline
        fn()
        n = 2
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': n / dt, 'unit': 'poses/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': f'{n} network-only poses of the oracle port after 1 warm-up '
                                          f'(torch {torch.__version__} CPU kernels)', **info}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        args.steps = 5 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference_arm(args, rank, world)
        return
    args.steps = 20 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else max(3, args.warmup)
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
