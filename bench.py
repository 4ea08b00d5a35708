#!/usr/bin/env python
"""Benchmark of the hot path: matched frame-pairs/s at 1241x376 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload flow|quad|mono]

A step = one new frame for every one of the S independent sequences of this rank (S frame pairs):
Matcher::pushBack + matchFeatures [+ bucketFeatures + RANSAC/pose for `mono`], results back on the host.
Sequences are sharded over GPUs (one rank per GPU under torchrun, no data-path collective; weak scaling: S is
per GPU).  `value` is measured with the frames already resident in HBM (pushBackDevice); `e2e` repeats the run
through the public host-buffer API (Matcher::pushBack from pinned host memory, H2D inside the timed region).
Timing: CUDA events on a context stream around the K steps, barrier + synchronise on both sides, max over ranks.
The rank-0 line also carries the roofline of the fused filter+NMS kernel (timed alone with events, 128 frames per
launch so the input exceeds L2) and the reference CPU path timed on this box's host cores (`cpu_baseline`).
"""
import argparse
import ctypes as C
import json
import os
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')   # one CUDA stream per host worker: see csrc/ctx.cu, visocu_create
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'opencl-structure-from-motion_b200'), os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

W, H = 1241, 376
METRIC = 'matched frame-pairs/sec at 1241x376'
UNIT = 'frame-pairs/s'
KITTI = dict(f=645.2, cu=635.9, cv=194.1, height=1.6, pitch=-0.08)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(',')])

    def mark(self):
        """start (first call) / end (second call) of the timed region"""
        self.marks = getattr(self, 'marks', []) + [time.time()]

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        marks = getattr(self, 'marks', [])
        rows = self.rows
        if len(marks) == 2:          # samples taken during the timed region (the 200 ms sampling grid is padded by one tick)
            inside = [r for r in rows if marks[0] - 0.25 <= r[0] <= marks[1] + 0.25]
            rows = inside or rows[-1:]
        rows = [r[1:] for r in rows]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def make_frames(n_frames, n_pool=4):
    """n_pool independent synthetic blob fields (generator A of SURVEY.md 8d), n_frames crops each, 3x1 px apart."""
    import synth
    return [synth.blob_sequence(n_frames, W, H, seed=1234 + s) for s in range(n_pool)]


def corridor_frames(n_frames, n_pool=2):
    import synth
    return [synth.corridor_sequence(n_frames, W, H, seed=1234 + s) for s in range(n_pool)]


def params_for(workload):
    import visocu_py as V
    import host_py as Hh
    if workload == 'quad':
        mp = V.Params(nms_n=2)                         # matlab/demo_matching_quad.m:12-21 of the reference
    else:
        mp = V.Params()
    kw = dict(KITTI)
    kw['bucket_max_features'] = 2
    return Hh.MonoParams(match=mp, **kw)


# ------------------------------------------------------------------------------------------------ reference arm
def reference_run(workload, steps, warmup, pairs_per_step=6):
    """The reference's own CPU implementation (oracle/_ref, unmodified sources) on all host cores: one sequence per
    thread, every step = pairs_per_step frame pairs per thread."""
    import pyref
    ref = pyref.RefLib()
    cores = os.cpu_count() or 1
    nfr = (steps + warmup) * pairs_per_step + 1
    if workload == 'mono':
        pools = corridor_frames(min(nfr, 40), 1)
    else:
        pools = make_frames(min(nfr, 66), 2)
    import synth
    rp = pyref.MatcherParams(nms_n=2) if workload == 'quad' else pyref.MatcherParams()
    results = {}

    def worker(tid, lo, n, out):
        seq = pools[tid % len(pools)]
        idx = [(lo + k) % len(seq) for k in range(n + 1)]
        imgs = np.ascontiguousarray(seq[idx])
        if workload == 'mono':
            mp = pyref.MonoParams(match=rp, bucket_max_features=2, **KITTI)
            tot, per, ok, _ = pyref.time_mono_sequence(ref, mp, imgs)
        elif workload == 'quad':
            right = np.ascontiguousarray(np.roll(imgs, -12, axis=2))
            tot, per, nm = pyref.time_matcher_sequence(ref, rp, 2, imgs, right, bucket=(2, 50.0, 50.0))
        else:
            tot, per, nm = pyref.time_matcher_sequence(ref, rp, 0, imgs)
        out[tid] = tot

    def run_step(step):
        out = {}
        th = [threading.Thread(target=worker, args=(t, step * pairs_per_step, pairs_per_step, out)) for t in range(cores)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for s in range(warmup):
        run_step(s)
    t0 = time.perf_counter()
    for s in range(steps):
        run_step(warmup + s)
    secs = time.perf_counter() - t0
    pairs = cores * pairs_per_step * steps
    return pairs / secs, secs, cores, '%d threads x %d frame pairs per step, %d steps (oracle/_ref = unmodified reference sources, g++ -O3 -DUSE_SIMD)' % (
        cores, pairs_per_step, steps)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import visocu_py as V
    import host_py as Hh
    workload = args.workload
    S = args.sequences
    K, Wm = args.steps, args.warmup
    # host workers per GPU: one per core of the rank's share, but at least 8 (a worker mostly waits for its stream and
    # yields the core while it does, see visocu_stream_wait; measured on 8 GPUs / 32 cores: 4 workers 132 k, 8 workers 150 k pairs/s)
    threads = args.threads or max(8, (os.cpu_count() or 1) // max(1, min(world, 8)))
    if threads * world > (os.cpu_count() or 1):
        # more workers than cores: waiting workers sleep between polls instead of yielding (8 GPUs / 32 cores: 157 k -> 172 k)
        os.environ.setdefault('VISOCU_WAIT_SLEEP_US', '100')
    mp = params_for(workload)
    dims = np.array([W, H, W], np.int32)
    n_frames = K + Wm + 1
    pools = corridor_frames(n_frames, 2) if workload == 'mono' else make_frames(n_frames, 4)
    method = 2 if workload == 'quad' else 0
    mode = 1 if workload == 'mono' else 0
    Hh.set_device(local_rank)
    tctx = V.Context(local_rank)                       # timing / staging context
    info = tctx.device_info()

    # stage every frame once: pinned host copies (e2e) and device copies (value)
    fb = W * H
    npool = len(pools)
    pinned = tctx.host_alloc(npool * n_frames * fb)
    pin_arr = np.ctypeslib.as_array((C.c_uint8 * (npool * n_frames * fb)).from_address(pinned)).reshape(npool, n_frames, H, W)
    for p in range(npool):
        pin_arr[p] = pools[p]
    dev = tctx.device_alloc(npool * n_frames * fb)
    tctx.memcpy_h2d(dev, pin_arr)
    pinned_r = dev_r = None
    if workload == 'quad':
        # right camera = the same field seen 12 px further left (positive disparity), as in the reference arm
        pinned_r = tctx.host_alloc(npool * n_frames * fb)
        pin_r = np.ctypeslib.as_array((C.c_uint8 * (npool * n_frames * fb)).from_address(pinned_r)).reshape(npool, n_frames, H, W)
        for p in range(npool):
            pin_r[p] = np.roll(pools[p], -12, axis=2)
        dev_r = tctx.device_alloc(npool * n_frames * fb)
        tctx.memcpy_h2d(dev_r, pin_r)

    def frame_ptrs(base, k, right=False):
        # sequence s walks through pool s % npool
        if right:
            base = dev_r if base == dev else pinned_r
        return [base + ((s % npool) * n_frames + k % n_frames) * fb for s in range(S)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def timed(on_device):
        runner = Hh.Runner(local_rank, S, threads, mode, method, mp)
        sampler = ClockSampler(local_rank) if (rank == 0 and on_device) else None     # sampling starts before the warm-up
        base = dev if on_device else pinned
        stereo = workload == 'quad'
        k = 0
        runner.step(frame_ptrs(base, k), dims, frame_ptrs(base, k, True) if stereo else None, on_device=on_device)
        k += 1
        for _ in range(Wm):
            runner.step(frame_ptrs(base, k), dims, frame_ptrs(base, k, True) if stereo else None, on_device=on_device, bucket=stereo)
            k += 1
        l0 = runner.launches(); b0 = runner.transfer_bytes()
        Hh.stage_times(reset=True)
        tctx.sync(); barrier()
        if sampler:
            sampler.mark()
        tctx.timer_start()
        t0 = time.perf_counter()
        # the K timed steps run back to back inside the host library: every worker walks its own sequences through
        # the K frames, no barrier between steps (the sequences are independent)
        steps1 = [frame_ptrs(base, k + j) for j in range(K)]
        steps2 = [frame_ptrs(base, k + j, True) for j in range(K)] if stereo else None
        secs, nm, ok = runner.run(steps1, dims, steps2, on_device=on_device, bucket=stereo)
        total_matches = int(nm.sum()); oks = int(ok.sum())
        k += K
        tctx.sync()
        ms = tctx.timer_stop()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.mark()
        clocks = sampler.stop() if sampler else None
        barrier()
        l1 = runner.launches(); b1 = runner.transfer_bytes()
        ro = runner.outlier_stats()
        stages = {k: round(1e3 * v[0] / max(v[1], 1), 4) for k, v in Hh.stage_times().items()}
        runner.close()
        return dict(ms=ms, wall=wall, launches=l1 - l0, h2d=(b1[0] - b0[0]) / K, d2h=(b1[1] - b0[1]) / K,
                    matches_per_pair=total_matches / float(S * K), ok_frac=oks / float(S * K), clocks=clocks, stages=stages, outliers=ro)

    res_dev = timed(True)
    res_e2e = timed(False)

    # ---- roofline leg: the fused filter+NMS kernel alone, 128 frames per launch (input 60 MB read + 119 MB written)
    roof = None
    if rank == 0:
        nb = 128
        rctx = V.Context(local_rank)
        vp = V.Params(**{f: getattr(mp.match, f) for f, _ in V.Params._fields_})
        vp.half_resolution = 0      # full-resolution matching mode: 128 x 1248 x 376 x 3 B = 180 MB per launch > L2
        rctx.configure(vp, W, H, nb)
        ptrs = [dev + ((i % npool) * n_frames + (i // npool) % n_frames) * fb for i in range(nb)]
        frames = list(range(nb))
        for _ in range(3):
            rctx.push_frames(frames, ptrs=ptrs, bpl_in=W, on_device=True)
        rctx.profile(True)
        for _ in range(max(5, K)):
            rctx.push_frames(frames, ptrs=ptrs, bpl_in=W, on_device=True)
        fms, nl, nfr = rctx.profile_read()
        rctx.profile(False)
        bpl = W + 15 - (W - 1) % 16
        # algorithmic bytes per frame: the fused kernel reads the matching-resolution image and
        # writes du and dv at that resolution; blob/checkerboard responses stay on chip, maxima leave as 4 B codes
        if vp.half_resolution:
            wm, hm = W // 2, H // 2
            bplm = wm + 15 - (wm - 1) % 16
        else:
            wm, hm, bplm = W, H, bpl
        alg = 3.0 * bplm * hm
        peaks, which = measured_peaks()
        achieved = alg * nfr / (fms * 1e-3) / 1e9 if fms > 0 else 0.0
        roof = {'bound': 'hbm', 'kernel': 'k_filter_nms', 'achieved': round(achieved, 2), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': round(achieved / peaks['hbm_gbs'], 5),
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload, from the committed
                # `ncu --set full` capture profiles/r1_v2_filter_nms_raw.csv (60.72 MB + 110.33 MB)
                'traffic': 171055104, 'traffic_source': 'profiles/r1_v2_filter_nms_summary.md', 'peak_source': which + ' (MEASURED_PEAKS.json hbm_gbs)',
                'launch_ms': round(fms / max(nl, 1), 5), 'frames_per_launch': nb,
                'algorithmic_bytes_per_launch': int(alg * nb),
                'workload': '128 frames of 1241x376 per launch, half_resolution=0 (180 MB of algorithmic traffic per launch, larger than L2)',
                'note': '3 B per matching-resolution pixel (1 read + du + dv written); timed alone with CUDA events on the launching stream'}
        rctx.close()

    tctx.host_free(pinned); tctx.device_free(dev)
    if pinned_r:
        tctx.host_free(pinned_r); tctx.device_free(dev_r)
    tctx.close()
    return res_dev, res_e2e, roof, threads, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='flow', choices=['flow', 'quad', 'mono'])
    ap.add_argument('--sequences', type=int, default=64, help='independent sequences per GPU (config 5 uses 64)')
    ap.add_argument('--threads', type=int, default=0, help='host worker threads per rank (default: cores / ranks)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dry-run', action='store_true', help='no GPU work: fake per-rank timings over gloo (tests of the N>1 plumbing)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    workload_name = {'flow': 'flow_1241x376_matcher_defaults (configs[0]: pushBack + matchFeatures(0), half-res, multi-stage, pixel refinement)',
                     'quad': 'quad_1241x376_nms2 (configs[1]: stereo pushBack + matchFeatures(2) + bucketFeatures)',
                     'mono': 'mono_vo_1241x376 (configs[2]: VisualOdometryMono::process, 2000 RANSAC iterations)'}[args.workload]

    if args.impl == 'reference':
        if rank != 0:
            return
        val, secs, cores, sample = reference_run(args.workload, args.steps, args.warmup)
        print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': round(val, 2), 'unit': UNIT, 'n_gpus': args.gpus,
                          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(1e3 * secs / args.steps, 3),
                          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                          'config': {'workload': workload_name},
                          'cpu_baseline': {'value': round(val, 2), 'unit': UNIT, 'cores': cores, 'kind': 'reference', 'sample': sample},
                          'e2e': {'value': round(val, 2), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    tdev = 'cpu' if args.dry_run else 'cuda'
    if world > 1:
        import torch
        import torch.distributed as dist
        if args.dry_run:
            dist.init_process_group('gloo')
        else:
            torch.cuda.set_device(local_rank)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if args.dry_run:
        ms = 10.0 * (rank + 1) * args.steps
        fake = dict(ms=ms, wall=ms * 1e-3, launches=0, h2d=0, d2h=0, matches_per_pair=0.0, ok_frac=0.0, clocks=None)
        res_dev, res_e2e, roof, threads, info = fake, dict(fake), None, 1, {'name': 'none (dry run)'}
    else:
        res_dev, res_e2e, roof, threads, info = run_ours(args, rank, world, local_rank)
    ms_dev, ms_e2e = res_dev['ms'], res_e2e['ms']
    launches = res_dev['launches']
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
        l = torch.tensor([launches], dtype=torch.int64, device=tdev)
        dist.all_reduce(l, op=dist.ReduceOp.SUM)
        launches = int(l[0])
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    S, K = args.sequences, args.steps
    pairs = S * world * K
    line = {'metric': METRIC, 'value': round(pairs / (ms_dev * 1e-3), 2), 'unit': UNIT, 'n_gpus': world, 'steps': K,
            'warmup': args.warmup, 'ms_per_step': round(ms_dev / K, 4), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8', 'data': 'dry-run' if args.dry_run else 'synthetic',
            'config': {'workload': workload_name, 'sequences_per_gpu': S, 'frame_pairs_per_step': S * world,
                       'host_threads_per_gpu': threads, 'inputs': 'frames resident in HBM (pushBackDevice); every step uses new frames, '
                       '%d sequences x 0.47 MB per step, L2 not flushed between steps (working set of a step = %d MB of planes and records)' % (S, int(S * 3.3)),
                       'device': info['name'], 'matches_per_pair': round(res_dev['matches_per_pair'], 1),
                       'process_ok_fraction': res_dev['ok_frac'],
                       'host_ms_per_call': res_dev.get('stages'),
                       'outlier_removal': dict(where='device (csrc/outliers.cu; lists it declines go to the host implementation)',
                                               **res_dev.get('outliers', {}))},
            'clocks': res_dev['clocks'],
            'e2e': {'value': round(pairs / (ms_e2e * 1e-3), 2), 'unit': UNIT, 'h2d_bytes_per_step': int(res_e2e['h2d'] * world),
                    'd2h_bytes_per_step': int(res_e2e['d2h'] * world), 'api': 'Matcher::pushBack(host image) + matchFeatures + getMatches'},
            'gpu_launches': launches, 'roofline': roof}
    if not args.no_cpu_baseline and not args.dry_run:
        try:
            val, secs, cores, sample = reference_run(args.workload, 5, 1)
            line['cpu_baseline'] = {'value': round(val, 2), 'unit': UNIT, 'cores': cores, 'kind': 'reference', 'sample': sample}
        except Exception as e:     # the checker library is prebuilt; say so instead of inventing a number
            line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'reference', 'sample': 'unavailable: %s' % e}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
