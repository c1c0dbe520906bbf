#!/usr/bin/env python
"""bench.py -- chunks/s and real-time factor of the full encode -> stage 1 -> stage 2 -> vocode path.

  python bench.py --gpus 1 --steps K --warmup W            (driver; N > 1 via torch.distributed.run)
  python bench.py --impl reference ...                      (the CPU implementation of the same path)

A "step" is one 0.3 s @ 24 kHz chunk of one audio stream pushed through the device-resident session
(BASELINE.json configs[1]: single stream, buffer_time = 0.3 s, extras (0, 0.5, 0), frame 5 ms;
stage-1 / stage-2 U-Nets at base width 64 with seeded synthetic weights; synthetic speech).
  value : chunks/s with the chunk's samples already in HBM and the output left there
          (ryk_session_push_device), K consecutive chunks, CUDA events, max over ranks
  e2e   : the same K chunks through ryk_session_push with HOST buffers (H2D + kernels + D2H per step)
Multi-GPU: one independent stream per rank ("weak"); NCCL only broadcasts the weights at init.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BUFFER_TIME = 0.3
EXTRA = (0.0, 0.5, 0.0)
FS = 24000
THRESHOLD_DB = 60.0
METRIC = 'chunks_per_s_0.3s_24kHz_encode_stage1_stage2_vocode'
WORKLOAD = ('single stream per GPU, buffer_time=0.3 s, extras (0,0.5,0), frame_period 5 ms, 24 kHz in/out, '
            'convert window 260 -> 384 frames, stage-1 1-D U-Net base 64 (13.6 M params), '
            'stage-2 2-D U-Net base 64 on 384x512 (54.4 M params, 142 GFLOP/chunk), WORLD DIO+StoneMask/CheapTrick/D4C + realtime synthesis')

# DRAM traffic of one stage-2 k4 block (14 k_conv_tc + 9 split-K reduce launches, 384x512, batch 1) from the committed ncu
# capture profiles/r01c_ncu_full_one_step.csv: 284.67 MB read + 0.11 MB written (algorithmic: 108.8 MB fp16 weights +
# ~100 MB activations written/read once; under ncu every replay starts with cold caches, so activations are re-read).
STAGE2_BLOCK_DRAM_BYTES = 284.77e6

# algorithmic work of the stage-2 k4 layers (the tcgen05 kernel launches) for one 384x512 forward, base 64
STAGE2_TC_FLOP = None


def stage2_tc_flop(Tp=384, base=64):
    enc = [1, 2, 4, 8, 8, 8, 8, 8]
    dec = [8, 8, 8, 8, 4, 2, 1]
    fl = 0.0
    for i in range(1, 8):
        cin, cout = base * enc[i - 1], base * enc[i]
        fl += 2.0 * 16 * cin * cout * (Tp >> i) * (512 >> i)
    for d in range(7):
        cin = base * enc[7] if d == 0 else base * dec[d - 1] + base * enc[7 - d]
        cout = base * dec[d]
        fl += 2.0 * 16 * cin * cout * (Tp >> (7 - d)) * (512 >> (7 - d))
    return fl


def measured_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        # the stage-2 block is timed INSIDE the long pipelined step (not alone), so the sustained cuBLAS figure is the denominator
        sus = d.get('bf16_tflops_sustained')
        if sus:
            return dict(tflops=float(sus), hbm=float(d['hbm_gbs']), burst=float(d['bf16_tflops']),
                        source='measured (MEASURED_PEAKS.json, cuBLAS bf16 sustained: the kernel is timed inside a long step; burst figure in peak_burst)')
        return dict(tflops=float(d['bf16_tflops']), hbm=float(d['hbm_gbs']), burst=float(d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json, cuBLAS bf16 burst)')
    return dict(tflops=1590.0, hbm=6650.0, burst=1590.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.ready.set()
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
                getattr(pynvml, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                getattr(pynvml, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                getattr(pynvml, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
            }
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as exc:        # clocks are best-effort; never fail the bench for them
            self.reasons.add(f'unavailable:{type(exc).__name__}')

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def make_models(rank, world):
    """rank 0 writes the seeded synthetic model files; with N > 1 the arrays are NCCL-broadcast so that
    every rank uploads identical weights (the only collective of the whole job)."""
    from realtime_yukarin_b200 import synthetic
    d = Path(tempfile.mkdtemp(prefix=f'ryk_bench_r{rank}_'))
    if world == 1:
        return synthetic.write_synthetic_models(d, seed=0)
    from realtime_yukarin_b200.distributed import broadcast_params
    p1 = broadcast_params(synthetic.make_stage1_params(0) if rank == 0 else None, src=0, device='cuda')
    p2 = broadcast_params(synthetic.make_stage2_params(0) if rank == 0 else None, src=0, device='cuda')
    paths = synthetic.write_synthetic_models(d, seed=0, base1=8, base2=8)     # configs / statistics files
    np.savez(paths['stage1_model_path'], **p1)
    np.savez(paths['stage2_model_path'], **p2)
    return paths


class CpuPath:
    """The CPU implementation of the same path (oracle port: C WORLD/SPTK + torch-CPU convs), one stream."""

    def __init__(self, paths, stream=0, threads=None, n_chunks=64):
        import torch
        from oracle import nets as onets
        from oracle import pipeline as opipe
        from realtime_yukarin_b200 import synthetic
        if threads:
            torch.set_num_threads(threads)
        self.cores = torch.get_num_threads()
        p1, p2 = onets.load_npz(paths['stage1_model_path']), onets.load_npz(paths['stage2_model_path'])
        stats = (float(np.log(150.0)), 0.2, float(np.log(250.0)), 0.2)
        cfg = opipe.PathConfig(threshold_db=THRESHOLD_DB)
        self.orc = opipe.StreamOracle(cfg, p1, p2, stats, buffer_time=BUFFER_TIME, extra=EXTRA, backend='torch')
        self.n = round(BUFFER_TIME * FS)
        self.x = synthetic.synthetic_speech((n_chunks + 1) * BUFFER_TIME, stream=stream)
        self.k = 0

    def step(self):
        k = self.k % (len(self.x) // self.n)
        self.orc.push(self.x[k * self.n:(k + 1) * self.n])
        self.k += 1

    def rate(self, n_chunks):
        t0 = time.perf_counter()
        for _ in range(n_chunks):
            self.step()
        return n_chunks / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from realtime_yukarin_b200 import synthetic
    d = Path(tempfile.mkdtemp(prefix='ryk_ref_'))
    paths = synthetic.write_synthetic_models(d, seed=0)
    t0 = time.perf_counter()
    cpu = CpuPath(paths, n_chunks=args.steps + args.warmup + 1)
    for _ in range(max(1, args.warmup)):
        cpu.step()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        cpu.step()
    dt = time.perf_counter() - t1
    value = args.steps / dt
    line = dict(
        impl='reference', metric=METRIC, value=value, unit='chunks/s', rtf=value * BUFFER_TIME, n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=1000.0 * dt / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f64 (WORLD/SPTK) + f32 (U-Nets)', data='synthetic',
        config=dict(workload=WORKLOAD, note='reference arm = CPU restatement of the path on the host cores (the reference itself cannot run: its '
                                            'arithmetic lives in un-vendored pyworld/pysptk/chainer, SURVEY 8c); one step = one 0.3 s chunk of one stream'),
        cpu_baseline=dict(value=value, unit='chunks/s', cores=cpu.cores, kind='port',
                          sample=f'{args.steps} consecutive 0.3 s chunks of one stream after {max(1, args.warmup)} warm-up chunk(s)'),
        e2e=dict(value=value, unit='chunks/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        wall_s=time.perf_counter() - t0)
    print(json.dumps(line))


def run_gpu(args):
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from realtime_yukarin_b200 import synthetic
    from realtime_yukarin_b200.engine import Engine, SessionConfig
    from realtime_yukarin_b200.models import AcousticConverter, F0Converter, SuperResolution
    from realtime_yukarin_b200.params import create_from_json, create_sr_from_json

    paths = make_models(rank, world)
    eng = Engine(device=local_rank)
    f0c = F0Converter(paths['input_statistics_path'], paths['target_statistics_path'])
    AcousticConverter(create_from_json(paths['stage1_config_path']), paths['stage1_model_path'], f0_converter=f0c, engine=eng)
    SuperResolution(create_sr_from_json(paths['stage2_config_path']), paths['stage2_model_path'], engine=eng)
    eng.set_precision('fp16')

    T, B = args.buffer_time, args.streams_per_gpu
    default_workload = (B == 1 and abs(T - BUFFER_TIME) < 1e-9)
    Tw = round((T + 2 * EXTRA[1]) * 200)
    Tp = Tw + (128 - Tw % 128)

    def new_session():
        cfg = SessionConfig(fs=FS, frame_period_ms=5.0, f0_floor=71.0, f0_ceil=800.0, fft_length=1024, order=8, alpha=0.466,
                            buffer_time=T, encode_extra_time=EXTRA[0], convert_extra_time=EXTRA[1], decode_extra_time=EXTRA[2],
                            threshold_db=THRESHOLD_DB, vocoder_buffer_size=1024)
        return eng.session_create(cfg)

    def new_streams():
        """B sessions of this rank; B > 1: grouped so that stage 2 runs once per step at batch B (BASELINE config 5)."""
        sids = [new_session() for _ in range(B)]
        return sids, (eng.group_create(sids) if B > 1 else None)

    def free_streams(sids, gid):
        if gid is not None:
            eng.group_destroy(gid)
        for sid in sids:
            eng.session_destroy(sid)

    n = round(T * FS)
    total = args.warmup + args.steps
    xs = [synthetic.synthetic_speech((total + 1) * T, stream=rank * B + j) for j in range(B)]
    chunks = [[np.ascontiguousarray(x[k * n:(k + 1) * n]) for x in xs] for k in range(total)]     # [step][stream]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- leg 1: device-resident ("value") ----
    sids, gid = new_streams()
    d_in = torch.from_numpy(np.stack([np.stack(c) for c in chunks])).cuda()       # [step][stream][n]
    out_cap = (n // 1024 + 5) * 1024 + 8192
    RING = 8                                     # distinct output slots: consecutive chunks are in flight together
    d_out = torch.empty((RING, B, out_cap), dtype=torch.float64, device='cuda')
    d_n = torch.zeros((RING, B), dtype=torch.int32, device='cuda')

    def push_dev(k):
        r = k % RING
        if gid is None:
            eng.session_push_device(sids[0], d_in[k, 0].data_ptr(), n, d_out[r, 0].data_ptr(), out_cap, d_n[r, 0:].data_ptr())
        else:
            eng.group_push_device(gid, [d_in[k, j].data_ptr() for j in range(B)], n, [d_out[r, j].data_ptr() for j in range(B)], out_cap,
                                  [d_n[r, j:].data_ptr() for j in range(B)])

    eng.profile(True)                            # warm up the same (event-instrumented) graphs the timed region replays
    for k in range(args.warmup):
        push_dev(k)
    barrier()
    eng.profile_read()                           # discard the warm-up timings
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(timeout=5)
    launches0 = eng.launch_count
    eng.profile(True)
    eng.timer_start()
    t_host0 = time.perf_counter()
    trace = []
    for k in range(args.warmup, total):
        push_dev(k)
        trace.append(time.perf_counter())
    if os.environ.get('RYK_BENCH_TRACE') == '1':
        print('host us per push:', [round((b - a) * 1e6) for a, b in zip([t_host0] + trace[:-1], trace)], file=sys.stderr)
    t_host = time.perf_counter() - t_host0          # host time to queue the K steps (launch overhead view)
    t_dev = eng.timer_stop() * 1e-3          # CUDA events on the stream the kernels are launched on
    barrier()
    s2_ms, s2_runs = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    t_dev = max_over_ranks(t_dev)
    stage_times = None
    if os.environ.get('RYK_STAGE_TIMES') == '1':       # diagnostics: device time of each stage of the last pipelined step
        st, en = eng.session_stage_times(sids[0])
        stage_times = dict(stages=['gate_slides', 'world_analysis', 'stage1', 'stage2', 'synthesis'],
                           start_ms=np.round(st, 3).tolist(), end_ms=np.round(en, 3).tolist())
    free_streams(sids, gid)

    # ---- leg 2: end to end with host buffers ("e2e") ----
    sids, gid = new_streams()
    host_out = [np.empty(out_cap, dtype=np.float64) for _ in range(B)]
    produced = 0
    DEPTH = 3                                    # steps in flight (submit k, collect k - DEPTH): host buffers both ways

    def submit(k):
        return eng.session_submit(sids[0], chunks[k][0]) if gid is None else eng.group_submit(gid, chunks[k])

    def collect(t):
        if gid is None:
            return len(eng.session_collect(sids[0], t, host_out[0]))
        return sum(len(o) for o in eng.group_collect(gid, t, host_out))

    for k in range(args.warmup):
        collect(submit(k))
    barrier()
    t0 = time.perf_counter()
    tickets = []
    for k in range(args.warmup, total):
        tickets.append(submit(k))
        if len(tickets) > DEPTH:
            produced += collect(tickets.pop(0))
    while tickets:
        produced += collect(tickets.pop(0))
    t_e2e = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    free_streams(sids, gid)

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    value = world * B * args.steps / t_dev
    e2e = world * B * args.steps / t_e2e
    peaks = measured_peaks()
    fl = stage2_tc_flop(Tp) * B
    ach = fl * s2_runs / (s2_ms * 1e-3) / 1e12 if s2_ms > 0 else None
    cpu_rate = cores = None
    if world == 1 and default_workload:                      # reported baseline: rank 0 at N = 1 only
        cpu = CpuPath(paths, n_chunks=8)
        cpu.step()
        cpu_rate, cores = cpu.rate(3), cpu.cores
    metric = METRIC if default_workload else f'chunks_per_s_{T:g}s_24kHz_encode_stage1_stage2_vocode'
    workload = WORKLOAD if default_workload else (
        f'{B} stream(s) per GPU' + (' grouped: one batched stage-2 forward per step' if B > 1 else '') +
        f', buffer_time={T:g} s, extras (0,0.5,0), frame_period 5 ms, 24 kHz in/out, convert window {Tw} -> {Tp} frames, '
        f'stage-2 input ({B},1,{Tp},512), same models as the default workload; one step = one chunk of every stream')
    line = dict(
        metric=metric, value=value, unit='chunks/s', rtf=value * T, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=1000.0 * t_dev / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f64 (WORLD analysis/synthesis), f32 (stage 1), f16 in / f32 accumulate (stage 2 tcgen05)', data='synthetic',
        config=dict(workload=workload, timing='CUDA events on the engine stream (forked to / joined from the session streams) around the K pushes, max over ranks', pipeline='gate | analysis (2 chunks in flight) | stage 1 | stage 2 | synthesis of consecutive chunks overlap on 6 CUDA streams per audio stream, each stage a CUDA graph (the reference overlaps its 3 worker processes); e2e keeps 3 steps in flight',
                    l2='per-step footprint (109 MB fp16 stage-2 weights + 54 MB stage-1 weights + ~100 MB activations) exceeds the 126 MB L2; no explicit flush',
                    streams_per_gpu=B, silence_threshold_db=THRESHOLD_DB),
        e2e=dict(value=e2e, unit='chunks/s', rtf=e2e * T, h2d_bytes_per_step=B * n * 4,
                 d2h_bytes_per_step=int(produced / max(1, args.steps)) * 8 + B * (4 + 8)),
        gpu_launches=int(launches), host_enqueue_ms_per_step=1000.0 * t_host / args.steps,
        clocks=clocks,
        roofline=dict(bound='tensor', kernel='k_conv_tc (stage-2 k4 layers 1..14, incl. split-K memset/finalize)' + ('' if B == 1 else ' + the two 3x3 edge layers (group forward timed as a whole)'), achieved=ach, peak=peaks['tflops'],
                      unit='TFLOP/s', frac=(ach / peaks['tflops']) if ach else None, traffic=STAGE2_BLOCK_DRAM_BYTES if default_workload else None,
                      traffic_source='ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum summed over the 23 launches of one 384x512 stage-2 k4 block (cold caches per replay): profiles/r01c_ncu_full_one_step.csv', peak_source=peaks['source'], peak_burst=peaks['burst'],
                      flop_per_step=fl, ms_per_step_in_kernel=(s2_ms / s2_runs) if s2_runs else None),
    )
    if stage_times is not None:
        line['stage_timeline'] = stage_times
    if cpu_rate is not None:
        line['cpu_baseline'] = dict(value=cpu_rate, unit='chunks/s', cores=cores, kind='port',
                                    sample='3 chunks of 0.3 s after a warm-up chunk, C WORLD/SPTK restatement + torch-CPU U-Nets, same models/audio')
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--streams-per-gpu', type=int, default=1, help='B > 1: BASELINE config 5 style, B grouped streams per GPU')
    ap.add_argument('--buffer-time', type=float, default=BUFFER_TIME, help='seconds per chunk (default workload: 0.3)')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == 'b200' else 6
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
