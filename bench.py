#!/usr/bin/env python
"""Headline benchmark: SNAC-24k audio-seconds/sec of the token->waveform hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A *step* is one decode tick: every stream of this rank's partition contributes one 28-token
(4-frame) sliding window, the tick is decoded in one engine call and emits 2048 samples (85.33 ms
of 24 kHz audio) per stream.  Workload = BASELINE config 4's per-GPU figure: 1024 concurrent
streams per GPU (weak scaling: every added GPU brings its own 1024-stream partition; no data-path
collective, streams are independent).  ``value`` is measured with tokens already resident in HBM;
``e2e`` goes through the host-buffer C-ABI call the Python ``convert_to_audio_batch`` makes (pinned
host tokens -> H2D -> kernels -> D2H PCM + status inside the timed region).

One JSON line on stdout (rank 0).  Only the ``cpu_baseline`` leg and ``--impl reference`` touch
``oracle/``.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAMPLES_PER_WINDOW = 2048
SAMPLE_RATE = 24000.0
AUDIO_S_PER_WINDOW = SAMPLES_PER_WINDOW / SAMPLE_RATE
FLOP_PER_WINDOW_REFERENCE = 3.312e9   # what model.decode executes for a 4-frame window
FLOP_PER_WINDOW_CONE = 1.363e9        # exact dependency cone of samples [2048,4096)
METRIC = "snac24k_audio_seconds_per_second"
UNIT = "audio-s/s"
PROFILE_STEP = os.environ.get("SNACB_PROFILE_STEP", "0") == "1"  # scripts/gpu_profile_r02.sh


def synth_tokens(first_stream: int, n: int, frames: int) -> np.ndarray:
    """SURVEY 8(d) token recipe: stream s -> PCG64(1234+s), ids U{1..4095}; [n, 7*frames] int32."""
    out = np.empty((n, 7 * frames), dtype=np.int32)
    for i in range(n):
        rng = np.random.Generator(np.random.PCG64(1234 + first_stream + i))
        out[i] = rng.integers(1, 4096, size=7 * frames, dtype=np.int64)
    return out


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tensor_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "tensor_tflops_burst": float(d.get("bf16_tflops", 1590.0)),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tensor_tflops": 1400.0, "tensor_tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 8:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(names, p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(n_windows: int, frames: int, steps: int, warmup: int, budget_s: float | None = None):
    """The reference's CPU algorithm (oracle port: verbatim speechpipe semantics + restated SNAC decode),
    one B=1 ``convert_to_audio``-equivalent per window exactly like the reference serialises them,
    torch intra-op threads = all host cores.  Returns (windows_per_s, seconds_per_step list, cores)."""
    import torch

    from oracle import snac_ref, speechpipe_ref as sp
    from project_morpheus_b200 import weights

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_grad_enabled(False)
    model = snac_ref.SNAC.from_state_dict(weights.random_state_dict(0, "w1")).eval()
    model.set_noise("randn")  # reference behaviour: fresh torch.randn inside decode
    tok = synth_tokens(0, n_windows, frames)

    def decode(c0, c1, c2):
        codes = [torch.from_numpy(c.astype(np.int64))[None] for c in (c0, c1, c2)]
        return model.decode(codes)[0, 0].numpy()

    def one_step():
        t0 = time.perf_counter()
        for row in tok:
            out = sp.window_to_pcm(row.tolist(), decode)
            assert out is not None and len(out) == 4096
        return time.perf_counter() - t0

    for _ in range(warmup):
        one_step()
    times, spent = [], 0.0
    for _ in range(steps):
        dt = one_step()
        times.append(dt); spent += dt
        if budget_s is not None and spent >= budget_s and len(times) >= 2:
            break
    wps = n_windows * len(times) / sum(times)
    return wps, times, cores


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_windows
    wps, times, cores = cpu_reference_run(n, args.frames, args.steps, args.warmup)
    value = wps * AUDIO_S_PER_WINDOW
    sample = f"{n} of the {args.streams} windows of one tick per step, B=1 decode per window (the reference has no batching)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "streams_per_gpu": args.streams, "frames_per_window": args.frames,
                   "tokens_per_window": 7 * args.frames, "weights": "random-init seed 0 variant w1"},
        "windows_per_s": wps,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args) -> str:
    return (f"cfg4-per-gpu: {args.streams} concurrent streams per GPU, one {args.frames}-frame "
            f"({7 * args.frames}-token) sliding window per stream per tick, emits samples [2048,4096) as int16")


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from project_morpheus_b200 import _lib, weights
    from project_morpheus_b200.engine import SnacEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    S, F, K, W = args.streams, args.frames, args.steps, max(3, args.warmup)
    eng = SnacEngine(weights.random_state_dict(0, "w1"), device=local, precision=args.precision, trim=not args.no_trim,
                     chunk_items=args.chunk, lanes=args.lanes, persistent_ru=args.persistent_ru)
    my_streams = rank + world * np.arange(S, dtype=np.int64)     # this rank's partition: stream s -> rank s mod G (partition.py)
    tok_host = np.concatenate([synth_tokens(int(s_), 1, F) for s_ in my_streams])
    keys = my_streams.astype(np.uint64)                          # Philox stream keys
    tok_dev = torch.from_numpy(tok_host).to(dev)
    pcm_dev = torch.empty((S, 2048), dtype=torch.int16, device=dev)
    st_dev = torch.empty((S,), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def tick_device(step):
        eng.decode_windows_device(tok_dev, noise="philox", seed=step, keys=keys, pcm=pcm_dev, status=st_dev)

    def tick_host(step):
        return eng.decode_windows(tok_host, noise="philox", seed=step, keys=keys)

    for i in range(W):
        tick_device(i)
        tick_host(i)
    torch.cuda.synchronize(dev)
    assert int((st_dev != _lib.WIN_OK).sum()) == 0

    # ---- timed region 1: device-resident inputs (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = eng.launch_count
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()                                  # untimed L2 flush between steps
        evs[i][0].record()
        prof = PROFILE_STEP and i == K - 1  # `ncu --profile-from-start off` then captures exactly one tick
        if prof:
            torch.cuda.synchronize(dev)
            torch.cuda.profiler.start()
        tick_device(100 + i)
        if prof:
            torch.cuda.synchronize(dev)
            torch.cuda.profiler.stop()
        evs[i][1].record()
    barrier()
    wall_dev = time.perf_counter() - wall0
    launches = eng.launch_count - launches0
    dev_ms = [a.elapsed_time(b) for a, b in evs]

    # ---- timed region 2: host buffers through the C-ABI (e2e)
    evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    e2e_wall = []
    for i in range(K):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        evs2[i][0].record()
        pcm, st = tick_host(200 + i)
        evs2[i][1].record()
        evs2[i][1].synchronize()
        e2e_wall.append(time.perf_counter() - t0)
        checksum = int(pcm[:, ::257].astype(np.int64).sum())  # touch the result on the host
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    e2e_ms = [max(a.elapsed_time(b), 1e3 * w) for (a, b), w in zip(evs2, e2e_wall)]

    # ---- timed region 3: the same host-buffer ticks through the pipelined C-ABI pair (submit t+1, then wait t): every
    # step still copies its tokens host -> device and its PCM + statuses device -> host inside the timed region, the
    # copy-back and the host work of step t run under the kernels of step t+1.  Per-tick activations (GBs) exceed L2.
    for i in range(max(W, 2)):  # untimed: sizes the two slots' pinned / device staging
        eng.wait_windows(eng.submit_windows(tok_host, noise="philox", seed=150 + i, keys=keys))
    barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    pending, checksum_p = [], 0
    for i in range(K):
        pending.append(eng.submit_windows(tok_host, noise="philox", seed=200 + i, keys=keys))
        if len(pending) == 2:
            pcm_p, st_p = eng.wait_windows(pending.pop(0))
            checksum_p += int(pcm_p[:, ::257].astype(np.int64).sum())
    while pending:
        pcm_p, st_p = eng.wait_windows(pending.pop(0))
        checksum_p += int(pcm_p[:, ::257].astype(np.int64).sum())
    pipe_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    t_dev = torch.tensor([sum(dev_ms), sum(e2e_ms), pipe_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    tot_dev_ms, tot_e2e_ms, tot_pipe_ms = float(t_dev[0]), float(t_dev[1]), float(t_dev[2])

    # ---- sustained: back-to-back device ticks for >= 2 s (no flush, no host work in between: the power-capped figure;
    # the per-tick activation traffic is GBs, far beyond L2), with its own clock samples
    sus_sampler = ClockSampler(local)
    if rank == 0:
        sus_sampler.start()
    barrier()
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_sus = max(8, int(args.sustained_s / max(1e-4, (sum(dev_ms) / K) * 1e-3)))
    ev_a.record()
    for i in range(n_sus):
        tick_device(400 + i)
    ev_b.record()
    barrier()
    sus_ms = torch.tensor([ev_a.elapsed_time(ev_b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sus_ms, op=dist.ReduceOp.MAX)
    sus_clocks = sus_sampler.stop() if rank == 0 else {}
    sustained = {"value": world * S * n_sus * AUDIO_S_PER_WINDOW / (float(sus_ms[0]) * 1e-3), "unit": UNIT, "ticks": n_sus,
                 "seconds": float(sus_ms[0]) * 1e-3, "ms_per_step": float(sus_ms[0]) / n_sus, "clocks": sus_clocks,
                 "note": "device-resident inputs, ticks issued back to back, max over ranks"}

    # ---- BASELINE config 4 AS WRITTEN (strong scaling): 1024 streams in total, stream s on rank s mod G, one tick = every
    # rank decodes its 1024/G windows through the host API and the PCM of all streams is gathered on rank 0 (host-side
    # gather over gloo) INSIDE the timed region.
    gloo = dist.new_group(backend="gloo") if world > 1 else None
    from project_morpheus_b200.partition import PartitionedDecoder, local_streams, partition_of
    cfg4 = None
    if args.cfg4_streams > 0:
        tot4 = args.cfg4_streams
        mine4 = local_streams(tot4, rank, world)
        tok4 = np.concatenate([synth_tokens(s_, 1, F) for s_ in mine4]) if mine4 else np.zeros((0, 7 * F), np.int32)
        keys4 = np.asarray(mine4, dtype=np.uint64)
        pd = PartitionedDecoder(lambda w: [], rank=rank, world_size=world, group=gloo) if world > 1 else None
        equal = (tot4 % world == 0)

        def tick4(step):
            pcm4, st4 = eng.decode_windows(tok4, noise="philox", seed=step, keys=keys4)
            if pd is not None and equal:
                return pd.gather_pcm(pcm4, dst=0)
            return pcm4

        # device-side gather (N > 1): pinned host tokens -> H2D -> kernels -> NCCL gather over NVLink onto rank 0 -> ONE
        # D2H of the whole tick there; every rank also reads its window statuses back
        use_nccl4 = world > 1 and equal and len(mine4) > 0
        if use_nccl4:
            tok4_pin = torch.from_numpy(tok4).pin_memory()
            tok4_dev = torch.empty(tok4.shape, dtype=torch.int32, device=dev)
            pcm4_dev = torch.empty((len(mine4), 2048), dtype=torch.int16, device=dev)
            st4_dev = torch.empty((len(mine4),), dtype=torch.int32, device=dev)
            st4_pin = torch.empty((len(mine4),), dtype=torch.int32).pin_memory()
            out4_pin = torch.empty((tot4, 2048), dtype=torch.int16).pin_memory() if rank == 0 else None

            def tick4_dev(step):
                tok4_dev.copy_(tok4_pin, non_blocking=True)
                eng.decode_windows_device(tok4_dev, noise="philox", seed=step, keys=keys4, pcm=pcm4_dev, status=st4_dev)
                st4_pin.copy_(st4_dev, non_blocking=True)
                g = pd.gather_pcm_device(pcm4_dev, dst=0, out_host=out4_pin)
                torch.cuda.current_stream(dev).synchronize()
                assert int(st4_pin.sum()) == 0
                return g

        def timed4(fn):
            for i in range(3):
                fn(i)
            barrier()
            t0 = time.perf_counter()
            g = None
            for i in range(n4):
                g = fn(500 + i)
            torch.cuda.synchronize(dev)
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return g, float(dt[0])

        n4 = max(4, K)
        g4, dt4 = timed4(tick4)
        g4d, dt4d = timed4(tick4_dev) if use_nccl4 else (None, None)
        if rank == 0:
            assert g4 is not None and g4.shape == (tot4 if (world == 1 or equal) else len(mine4), 2048)
            host = {"ms_per_tick": 1e3 * dt4 / n4, "audio_s_per_s": tot4 * n4 * AUDIO_S_PER_WINDOW / dt4,
                    "gather": ("gloo host gather of every stream's PCM onto rank 0 inside the timed region" if world > 1
                               else "single GPU: no gather"),
                    "api": "SnacEngine.decode_windows (host buffers in and out) per rank, PartitionedDecoder.gather_pcm"}
            cfg4 = {"streams_total": tot4, "streams_per_gpu": len(mine4), "ticks": n4, "scaling": "strong"}
            if use_nccl4:
                assert g4d is not None and np.array_equal(g4d, g4)  # same seed on the last tick: both gathers, same bytes
                cfg4.update({"ms_per_tick": 1e3 * dt4d / n4, "audio_s_per_s": tot4 * n4 * AUDIO_S_PER_WINDOW / dt4d,
                             "gather": "NCCL gather of every rank's device PCM onto rank 0 over NVLink + one D2H of the whole tick there, "
                                       "inside the timed region; bytes equal to the host-gather path",
                             "api": "pinned host tokens -> H2D -> SnacEngine.decode_windows_device -> PartitionedDecoder.gather_pcm_device",
                             "host_gather": host})
            else:
                cfg4.update(host)

    # ---- BASELINE config 5 on the whole box: 512 streams per GPU, ragged ticks (0/1/2 pending windows of 1/4/7 frames per
    # stream), scene lifetimes with evict + slot refill (barge_in) and re-homing of a stream to another partition after
    # three chunks (mid_stream_swap), every tick through PartitionedDecoder.decode_tick (host gather inside the timing).
    cfg5p = None
    if world > 1 and args.cfg5_streams_per_gpu > 0:
        ns5 = args.cfg5_streams_per_gpu * world
        rng5 = np.random.default_rng(2024)  # same seed on every rank: all ranks build the same global tick
        scene = rng5.integers(0, 4, size=ns5)  # breathing_room / long_read / barge_in / mid_stream_swap, 25 % each
        life = np.where(scene == 0, 2, np.where(scene == 1, 60, np.where(scene == 2, 2, 6)))
        age = np.zeros(ns5, dtype=np.int64)
        sid = np.arange(ns5, dtype=np.int64)  # current stream id of slot i (changes on refill / re-homing)
        next_id = ns5
        ticks5 = []
        stats5 = {"evicted": 0, "rehomed": 0, "windows": 0}
        for t in range(10):
            tick = []
            for i in range(ns5):
                for _ in range(int(rng5.choice([0, 1, 2], p=[0.2, 0.6, 0.2]))):
                    fr = 1 if age[i] == 0 else (7 if age[i] >= 4 else 4)
                    tick.append((int(sid[i]), synth_tokens(int(90000 + sid[i] * 13 + t), 1, 7)[0][: 7 * fr].tolist()))
                    age[i] += 1
                if scene[i] == 3 and age[i] == 3:      # mid_stream_swap: the stream moves to another adapter / partition
                    sid[i] = next_id + ((partition_of(int(sid[i]), world) + 1 - next_id) % world)
                    next_id += world
                    stats5["rehomed"] += 1
                if age[i] >= life[i]:                    # scene over (barge_in: evicted after chunk 2): slot refilled
                    sid[i], age[i] = next_id, 0
                    next_id += 1
                    stats5["evicted"] += 1
            stats5["windows"] += len(tick)
            ticks5.append(tick)

        def dec5(wins):
            n5 = len(wins)
            tok5 = np.zeros((n5, 49), dtype=np.int32)
            lens5 = [len(w) for w in wins]
            for i, w in enumerate(wins):
                tok5[i, : len(w)] = w
            pcm5, st5 = eng.decode_windows(tok5, ntok=lens5, noise="philox", seed=9, keys=np.arange(n5, dtype=np.uint64))
            return [pcm5[i].tobytes() if st5[i] == _lib.WIN_OK else (b"" if st5[i] == _lib.WIN_EMPTY else None) for i in range(n5)]

        pd5 = PartitionedDecoder(dec5, rank=rank, world_size=world, group=gloo)
        for tick in ticks5[:2]:
            pd5.decode_tick(tick, dst=0)
        barrier()
        t0 = time.perf_counter()
        emitted5 = 0
        for tick in ticks5:
            merged = pd5.decode_tick(tick, dst=0)
            if rank == 0:
                assert len(merged) == len({s_ for s_, _ in tick})
                emitted5 += sum(1 for v in merged.values() if v)
        torch.cuda.synchronize(dev)
        dt5 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt5, op=dist.ReduceOp.MAX)
        # the same ticks with the gather on the device side: NCCL over NVLink onto rank 0, one D2H of the tick there
        cap5 = max(1, max(sum(1 for s_, _ in tick if partition_of(s_, world) == rank) for tick in ticks5))
        tok5_pin = torch.zeros((cap5, 49), dtype=torch.int32).pin_memory()
        tok5_dev = torch.zeros((cap5, 49), dtype=torch.int32, device=dev)

        def dec5_dev(wins):
            n5 = len(wins)
            lens5 = [len(w) for w in wins]
            t5 = tok5_pin.numpy()
            t5[:n5] = 0
            for i, w in enumerate(wins):
                t5[i, : len(w)] = w
            tok5_dev[:n5].copy_(tok5_pin[:n5], non_blocking=True)
            return eng.decode_windows_device(tok5_dev[:n5], ntok=lens5, noise="philox", seed=9, keys=np.arange(n5, dtype=np.uint64))

        for tick in ticks5[:2]:
            pd5.decode_tick_device(tick, dec5_dev, dst=0)
        barrier()
        t0 = time.perf_counter()
        emitted5d, last_dev = 0, None
        for tick in ticks5:
            last_dev = pd5.decode_tick_device(tick, dec5_dev, dst=0)
            if rank == 0:
                emitted5d += sum(1 for v in last_dev.values() if v)
        torch.cuda.synchronize(dev)
        dt5d = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt5d, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert emitted5d == emitted5 and last_dev == merged  # both gathers: the same bytes for every stream
            cfg5p = dict(stats5, streams=ns5, streams_per_gpu=args.cfg5_streams_per_gpu, ticks=len(ticks5),
                         ms_per_tick=1e3 * float(dt5d[0]) / len(ticks5), audio_s_per_s=emitted5d * AUDIO_S_PER_WINDOW / float(dt5d[0]),
                         note="whole box, PartitionedDecoder.decode_tick_device per tick: ragged decode per rank (pinned host tokens -> "
                              "H2D -> kernels), NCCL gather of every rank's PCM + statuses onto rank 0 over NVLink, one D2H of the tick "
                              "there, {stream: bytes} built on rank 0 - all inside the timed region; evict / refill / re-homing are "
                              "host bookkeeping (the decoder is stateless across windows); bytes equal to the host-gather path",
                         host_gather={"ms_per_tick": 1e3 * float(dt5[0]) / len(ticks5),
                                      "audio_s_per_s": emitted5 * AUDIO_S_PER_WINDOW / float(dt5[0]),
                                      "note": "PartitionedDecoder.decode_tick: host decode per rank + gloo gather_object"})

    # ---- per-kernel-class device time (CUDA events around each launch), same tick, right after
    stats, extra = {}, {}
    if rank == 0:
        eng.profile(True)
        for i in range(2):
            tick_device(300 + i)
        torch.cuda.synchronize(dev)
        stats = eng.profile_read()
        eng.profile(False)
        # BASELINE config 2 (64 streams per tick) and single-window latency through the host API
        for name, n in (("cfg2_64_streams", 64), ("single_window", 1)):
            t = tok_host[:n]
            k = keys[:n]
            for i in range(5):
                eng.decode_windows(t, noise="philox", seed=i, keys=k)
            lat = []
            for i in range(args.latency_reps):
                t0 = time.perf_counter()
                eng.decode_windows(t, noise="philox", seed=i, keys=k)
                lat.append(time.perf_counter() - t0)
            p50 = statistics.median(lat)
            extra[name] = {"p50_ms": 1e3 * p50, "p95_ms": 1e3 * sorted(lat)[int(0.95 * (len(lat) - 1))],
                           "audio_s_per_s": n * AUDIO_S_PER_WINDOW / p50, "reps": len(lat)}

        # p50 of the reference-shaped Python call itself: speechpipe.convert_to_audio(28 tokens) -> bytes
        try:
            os.environ.setdefault("SNACB_RANDOM_INIT", "0:w1")
            os.environ.setdefault("SNACB_PRECISION", args.precision)
            import importlib
            sp_mod = importlib.import_module("project_morpheus_b200.speechpipe")
            win = tok_host[0].tolist()
            for i in range(10):
                sp_mod.convert_to_audio(win, i)
            lat = []
            for i in range(args.latency_reps * 5):
                t0 = time.perf_counter()
                out_b = sp_mod.convert_to_audio(win, i)
                lat.append(time.perf_counter() - t0)
            assert out_b is not None and len(out_b) == 4096
            extra["convert_to_audio_call"] = {"p50_ms": 1e3 * statistics.median(lat),
                                              "p95_ms": 1e3 * sorted(lat)[int(0.95 * (len(lat) - 1))], "reps": len(lat),
                                              "note": "project_morpheus_b200.speechpipe.convert_to_audio, Python list in -> bytes out"}
        except Exception as exc:  # noqa: BLE001
            extra["convert_to_audio_call"] = {"error": repr(exc)}

        # Through the reference's seam: N requests, one SnacB200Adapter + one pull loop each (orchestrator/core.py:89-117) under
        # ONE event loop; all of them decode through the shared DecodeTicker (one batched engine call per tick).
        if args.pull_streams:
            try:
                import asyncio
                from oracle import speechpipe_ref as sp_ref2
                from project_morpheus_b200.adapter import SnacB200Adapter
                sp_mod = importlib.import_module("project_morpheus_b200.speechpipe")
                res_pull = {}
                for n_ad in [int(x) for x in args.pull_streams.split(",") if x]:
                    fr_ad = 12
                    strs = [sp_ref2.synth_token_strings(60000 + i, fr_ad) for i in range(n_ad)]

                    def source(strings):
                        async def gen(**_):
                            for j, s_ in enumerate(strings):
                                if j % 7 == 0:
                                    await asyncio.sleep(0)
                                yield s_
                        return gen

                    async def pull_all(gpu_ring=True):
                        ads = [SnacB200Adapter("bench", "tara", token_source=source(st_), seed=i, gpu_ring=gpu_ring)
                               for i, st_ in enumerate(strs)]

                        async def loop(ad):
                            nb = 0
                            while True:
                                c = await ad.pull(4096)
                                nb += len(c.pcm)
                                if c.eos:
                                    return nb
                        return sum(await asyncio.gather(*[loop(a) for a in ads]))

                    tkr = sp_mod.get_ticker()
                    asyncio.run(pull_all())  # warm-up (workspace, graphs)
                    t_before = dict(tkr.stats())
                    t0 = time.perf_counter()
                    nbytes = asyncio.run(pull_all())
                    dt = time.perf_counter() - t0
                    t_after = tkr.stats()
                    asyncio.run(pull_all(False))
                    t0 = time.perf_counter()
                    nbytes_h = asyncio.run(pull_all(False))
                    dt_h = time.perf_counter() - t0
                    res_pull[f"{n_ad}_streams"] = {
                        "audio_s_per_s": nbytes / 2 / SAMPLE_RATE / dt, "seconds": dt, "bytes": nbytes,
                        "ticks": t_after["ticks"] - t_before["ticks"], "windows": t_after["windows"] - t_before["windows"],
                        "frames_per_stream": fr_ad,
                        "host_bytes_ring": {"audio_s_per_s": nbytes_h / 2 / SAMPLE_RATE / dt_h, "seconds": dt_h, "bytes": nbytes_h}}
                res_pull["note"] = ("token strings -> SnacB200Adapter.pull(4096) per request, per-token Python of tokens_decoder "
                                    "included; engine calls = ticks, not windows.  Main figure: the tick's PCM is written by the GPU "
                                    "into pinned per-stream rings (N3, csrc/egress_ring.cu) and pull() is one native ring read; "
                                    "host_bytes_ring: PCM matrix D2H + Python bytes per window (gpu_ring=False)")
                extra["ticker_pull"] = res_pull
            except Exception as exc:  # noqa: BLE001
                extra["ticker_pull"] = {"error": repr(exc)[:300]}

        # BASELINE config 5 pattern: ragged ticks (per stream 0/1/2 pending windows, 1 / 4 / 7 frames each)
        if args.ragged_streams > 0:
            rng = np.random.default_rng(2024)
            ns = args.ragged_streams
            ticks = []
            for t in range(12):
                wins, lens = [], []
                for sidx in range(ns):
                    for _ in range(int(rng.choice([0, 1, 2], p=[0.2, 0.6, 0.2]))):
                        fr = int(rng.choice([1, 4, 7], p=[0.05, 0.35, 0.6]))
                        wins.append(synth_tokens(90000 + sidx * 13 + t, 1, 7)[0][: 7 * fr])
                        lens.append(7 * fr)
                tokr = np.zeros((len(wins), 49), dtype=np.int32)
                for i, w in enumerate(wins):
                    tokr[i, : len(w)] = w
                ticks.append((tokr, lens))
            for tokr, lens in ticks[:3]:
                eng.decode_windows(tokr, ntok=lens, noise="philox", seed=1, keys=np.arange(len(lens), dtype=np.uint64))
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            emitted = 0
            for tokr, lens in ticks:
                pcm_r, st_r = eng.decode_windows(tokr, ntok=lens, noise="philox", seed=2, keys=np.arange(len(lens), dtype=np.uint64))
                emitted += int((st_r == _lib.WIN_OK).sum())
            dt = time.perf_counter() - t0
            extra["cfg5_ragged"] = {"streams": ns, "ticks": len(ticks), "windows": int(sum(len(l) for _, l in ticks)),
                                    "ms_per_tick": 1e3 * dt / len(ticks), "audio_s_per_s": emitted * AUDIO_S_PER_WINDOW / dt,
                                    "note": "host API, mixed 1/4/7-frame windows grouped by frame count inside one call"}

        # Next row N2: token strings -> PCM chunks for every stream of a tick, through the Python tick scheduler
        # (per-token Python like the reference's tokens_decoder, one batched decode) and through the native ingress
        # (csrc/ingest.cpp).  Steady state: 7 new token strings per stream per tick -> one 49-token (7-frame) window per stream.
        if args.ingest_streams > 0:
            from oracle import speechpipe_ref as sp_ref
            from project_morpheus_b200.ingest import NativeTickScheduler, decode_arrays_with
            from project_morpheus_b200.scheduler import TickScheduler
            ns, warm_f, timed_ticks = args.ingest_streams, 8, 6  # warm-up reaches the 49-token steady state (workspace sized)
            strings = [sp_ref.synth_token_strings(70000 + i, warm_f + timed_ticks) for i in range(ns)]
            kk = np.arange(ns, dtype=np.uint64)

            def dec_arrays(tok, ntok_or_none):
                return eng.decode_windows(tok, ntok=ntok_or_none, noise="philox", seed=3, keys=kk[: tok.shape[0]])

            def dec_lists(windows):
                tok = np.asarray(windows, dtype=np.int64).astype(np.int32)
                pcm_l, st_l = dec_arrays(tok, None)
                return [pcm_l[i].tobytes() if st_l[i] == _lib.WIN_OK else (b"" if st_l[i] == _lib.WIN_EMPTY else None)
                        for i in range(len(windows))]

            def null_arrays(tok, ntok):
                return np.zeros((tok.shape[0], 2048), dtype=np.int16), np.where(ntok >= 14, _lib.WIN_OK, _lib.WIN_EMPTY).astype(np.int32)

            def null_lists(windows):
                return [b"\0" * 4096 if len(w) >= 14 else b"" for w in windows]

            def drive(sched):
                for i in range(ns):
                    sched.add_stream(i)
                    sched.push_many(i, strings[i][: 7 * warm_f])
                sched.drain()
                for i in range(ns):
                    sched.pop_audio(i)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                got = 0
                for t in range(timed_ticks):
                    lo = 7 * (warm_f + t)
                    for i in range(ns):
                        sched.push_many(i, strings[i][lo: lo + 7])
                    sched.tick()
                    for i in range(ns):
                        got += len(sched.pop_audio(i))
                if got < ns * timed_ticks:  # pipelined scheduler: the last tick is still in flight
                    sched.drain()
                    for i in range(ns):
                        got += len(sched.pop_audio(i))
                dt = (time.perf_counter() - t0) / timed_ticks
                assert got == ns * timed_ticks, (got, ns, timed_ticks)
                return dt

            res = {}
            for name, mk in (("python_scheduler", lambda: TickScheduler(dec_lists, max_windows_per_tick=ns)),
                             ("native_ingest", lambda: NativeTickScheduler(max_streams=ns, max_windows_per_tick=ns,
                                                                           decode_arrays=lambda a, b: decode_arrays_with(dec_arrays, a, b))),
                             ("native_ingest_pipelined", lambda: NativeTickScheduler(max_streams=ns, max_windows_per_tick=ns, engine=eng,
                                                                                     noise="philox", seed=3)),
                             ("python_scheduler_host_only", lambda: TickScheduler(null_lists, max_windows_per_tick=ns)),
                             ("native_ingest_host_only", lambda: NativeTickScheduler(max_streams=ns, max_windows_per_tick=ns,
                                                                                     decode_arrays=null_arrays))):
                dt = drive(mk())
                res[name] = {"ms_per_tick": 1e3 * dt, "token_strings_per_s": 7 * ns / dt,
                             "audio_s_per_s": ns * AUDIO_S_PER_WINDOW / dt}
            res["streams"] = ns
            res["note"] = ("token strings in -> PCM chunks out for every stream, 7 new strings per stream per tick; "
                           "*_host_only replaces the decode by a no-op to isolate parsing + window planning + egress")
            extra["n2_token_ingress"] = res

        # Next row N3: PCM egress - the reference's overlap-add stitcher (numpy, one call per chunk) against the native one,
        # one 4096-byte chunk per stream per tick, 10 ms crossfade (the server's default overlap of 0 is a pass-through).
        if args.ingest_streams > 0:
            from oracle import egress_ref
            ns_e, ticks_e = args.ingest_streams, 8
            rng_e = np.random.default_rng(7)
            chunks_e = [rng_e.integers(-20000, 20000, size=2048).astype("<i2").tobytes() for _ in range(16)]
            from project_morpheus_b200.egress import StitcherBank
            bank_e = StitcherBank(ns_e, 24000, 10.0)
            mats = [np.stack([np.frombuffer(chunks_e[(i + t) % 16], dtype="<i2") for i in range(ns_e)]) for t in range(ticks_e)]
            slots_e = np.arange(ns_e, dtype=np.int32)
            t0 = time.perf_counter()
            nbytes = 0
            for t in range(ticks_e):
                _, ol, _ = bank_e.push_tick(slots_e, mats[t])
                nbytes += 2 * int(ol[ol > 0].sum())
            dt_nat = (time.perf_counter() - t0) / ticks_e
            bank_e.close()
            sts = []

            def one_stream(i):
                return sum(len(p_) for p_, _ in egress_ref.stitch(((chunks_e[(i + t) % 16], False) for t in range(ticks_e)), 24000, 10.0))

            sample = max(1, ns_e // 8)
            t0 = time.perf_counter()
            for i in range(sample):
                one_stream(i)
            dt_np = (time.perf_counter() - t0) / ticks_e * (ns_e / sample)
            for s_ in sts:
                s_.close()
            extra["n3_pcm_egress"] = {"streams": ns_e, "overlap_ms": 10.0, "native_ms_per_tick": 1e3 * dt_nat,
                                      "numpy_reference_ms_per_tick": 1e3 * dt_np, "bytes_per_tick": nbytes // ticks_e,
                                      "note": "overlap-add stitcher, one 2048-sample chunk per stream per tick, native = one StitcherBank.push_tick call per tick; numpy = the "
                                              "reference's algorithm (oracle restatement), scaled from a 1/8 sample of the streams"}
            # the same join on the GPU: ring kernel right after the decoder tail, PCM written into pinned per-stream rings
            try:
                from project_morpheus_b200.egress import GpuPcmRing
                gr = GpuPcmRing(ns_e, 8192, overlap_ms=10.0, device=local)
                d_mats = [torch.from_numpy(m).to(dev) for m in mats]
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                stream_ptr = torch.cuda.current_stream(dev).cuda_stream
                gr.push_device(slots_e, d_mats[0].data_ptr(), 2048, 2048, stream=stream_ptr)
                gr.sync(stream_ptr)
                for s_i in range(ns_e):
                    gr.read(s_i, 1 << 20)
                t_host, k_ms = 0.0, 0.0
                for t in range(1, ticks_e):
                    t0 = time.perf_counter()
                    ev0.record()
                    gr.push_device(slots_e, d_mats[t].data_ptr(), 2048, 2048, stream=stream_ptr)
                    ev1.record()
                    gr.sync(stream_ptr)
                    t_host += time.perf_counter() - t0
                    k_ms += ev0.elapsed_time(ev1)
                    for s_i in range(ns_e):
                        gr.read(s_i, 1 << 20)
                gr.close()
                extra["n3_pcm_egress"].update({"gpu_ring_ms_per_tick": 1e3 * t_host / (ticks_e - 1), "gpu_ring_kernel_ms": k_ms / (ticks_e - 1),
                                               "gpu_ring_note": "k_stitch_ring over the device PCM matrix of a tick -> pinned rings (wall time of push + sync; "
                                                                "kernel time by CUDA events); replaces the D2H matrix copy + the host join"})
            except Exception as exc:  # noqa: BLE001
                extra["n3_pcm_egress"]["gpu_ring_error"] = repr(exc)[:200]

        # precision safety net: the same tick through the split-operand recipe (SNACB_PREC_FP16X3, <= 2 LSB vs fp32)
        if args.precision == "fp16" and not args.no_cpu_baseline:
            try:
                eng3 = SnacEngine(weights.random_state_dict(0, "w1"), device=local, precision="fp16x3", trim=not args.no_trim)
                for i in range(2):
                    eng3.decode_windows_device(tok_dev, noise="philox", seed=i, keys=keys, pcm=pcm_dev, status=st_dev)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for i in range(3):
                    eng3.decode_windows_device(tok_dev, noise="philox", seed=10 + i, keys=keys, pcm=pcm_dev, status=st_dev)
                torch.cuda.synchronize(dev)
                dt3 = (time.perf_counter() - t0) / 3
                extra["precision_fp16x3"] = {"ms_per_tick": 1e3 * dt3, "audio_s_per_s": S * AUDIO_S_PER_WINDOW / dt3,
                                             "note": "two-term fp16 splits of both GEMM operands, 3 tcgen05 products per k-block, fp32 "
                                                     "elementwise kernels; <= 2 LSB vs the fp32 oracle (tests/test_gpu_parity.py)"}
                eng3.close()
            except Exception as exc:  # noqa: BLE001
                extra["precision_fp16x3"] = {"error": repr(exc)[:300]}

        # Next row N4: the encode direction (voice prompts / data preparation): audio -> 3 code levels, fp32 CUDA-core kernels
        if not args.no_cpu_baseline:
            try:
                sd_full = dict(weights.random_state_dict(0, "w1"))
                sd_full.update(weights.random_encoder_state_dict(0, "w1"))
                eng_e = SnacEngine(sd_full, device=local, precision="fp32")
                secs, Be = 10.0, 8
                aud = (0.3 * torch.randn((Be, 1, int(secs * SAMPLE_RATE)), generator=torch.Generator().manual_seed(3))).to(dev)
                eng_e.encode(aud)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                for _ in range(3):
                    codes_e = eng_e.encode(aud)
                torch.cuda.synchronize(dev)
                dt_e = (time.perf_counter() - t0) / 3
                extra["n4_encoder"] = {"batch": Be, "seconds_per_item": secs, "ms": 1e3 * dt_e, "audio_s_per_s": Be * secs / dt_e,
                                       "frames": int(codes_e[0].shape[1]),
                                       "note": "SNAC.encode (encoder + residual VQ) through snacb_encode, exact fp32 kernels"}
                eng_e.close()
            except Exception as exc:  # noqa: BLE001
                extra["n4_encoder"] = {"error": repr(exc)[:300]}

        # BASELINE config 3 (long_read): one-shot decode of 720-frame utterances, time-tiled; reduced batch by default
        if args.long_read_batch > 0:
            Fl, Bl = 720, args.long_read_batch
            lr_tok = synth_tokens(50000, Bl, Fl).reshape(Bl, Fl, 7)
            c0 = torch.from_numpy(np.ascontiguousarray(lr_tok[:, :, 0])).to(dev)
            c1 = torch.from_numpy(np.ascontiguousarray(lr_tok[:, :, [1, 4]].reshape(Bl, 2 * Fl))).to(dev)
            c2 = torch.from_numpy(np.ascontiguousarray(lr_tok[:, :, [2, 3, 5, 6]].reshape(Bl, 4 * Fl))).to(dev)
            eng.decode_codes([c0, c1, c2], noise="philox", seed=1)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            reps = 3
            for i in range(reps):
                eng.decode_codes([c0, c1, c2], noise="philox", seed=2 + i)
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / reps
            extra["cfg3_long_read"] = {"batch": Bl, "frames": Fl, "ms": 1e3 * dt,
                                       "audio_s_per_s": Bl * Fl * AUDIO_S_PER_WINDOW / dt,
                                       "note": "one-shot decode, 8-frame time tiles with halo recompute, no slice"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    windows_total = world * S * K
    value = windows_total * AUDIO_S_PER_WINDOW / (tot_dev_ms * 1e-3)
    e2e_value = windows_total * AUDIO_S_PER_WINDOW / (tot_e2e_ms * 1e-3)
    pipe_value = windows_total * AUDIO_S_PER_WINDOW / (tot_pipe_ms * 1e-3)
    wps = windows_total / (tot_dev_ms * 1e-3)

    # dominant kernel class and its roofline.  Which roof applies follows the roofline model: arithmetic
    # intensity (executed FLOPs / algorithmic bytes of the class) against the ridge point of the measured peaks.
    roof = None
    if stats:
        tot_ms = sum(v["ms"] for v in stats.values()) or 1e-9
        name, top = max(stats.items(), key=lambda kv: kv[1]["ms"])
        ridge = peaks["tensor_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
        ai = top["flops"] / max(top["bytes"], 1.0)
        tflops = top["flops"] / (top["ms"] * 1e-3) / 1e12
        gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        if ai >= ridge:
            roof = {"bound": "tensor", "achieved": tflops, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                    "frac": tflops / peaks["tensor_tflops"]}
        else:
            roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"]}
        launches_per_step = top["launches"] / 2
        traffic = None
        try:  # DRAM bytes per window of this kernel class from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if name in tj.get("dram_bytes_per_window_per_launch", {}):
                traffic = tj["dram_bytes_per_window_per_launch"][name] * S
        except Exception:  # noqa: BLE001
            traffic = None
        roof.update({
            "traffic": traffic, "kernel": name, "launches_per_step": launches_per_step,
            "avg_launch_ms": top["ms"] / max(1, top["launches"]), "share_of_step": top["ms"] / tot_ms,
            "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
            "tensor_tflops": tflops, "tensor_frac": tflops / peaks["tensor_tflops"], "hbm_gbs": gbs,
            "hbm_frac": gbs / peaks["hbm_gbs"], "bytes_per_launch": top["bytes"] / max(1, top["launches"]),
            "peak_source": peaks["source"] + " (bf16 dense sustained / copy)",
            "flops_counted": "executed (2*MAC of the launches, dependency-cone trimmed)",
            "classes": {k: {"ms_per_step": v["ms"] / 2, "launches_per_step": v["launches"] / 2,
                            "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0,
                            "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else 0.0}
                        for k, v in stats.items() if v["launches"]},
        })

    gpu_torch = None
    if world == 1 and not args.no_cpu_baseline:
        # second baseline: the same restated reference algorithm run by stock PyTorch/cuDNN on this B200,
        # the way the reference drives it (one B=1 decode per window) and, for information, as one B=64 batch
        try:
            from oracle import snac_ref
            torch.set_grad_enabled(False)
            ref_model = snac_ref.SNAC.from_state_dict(weights.random_state_dict(0, "w1")).eval().to(dev)
            ref_model.set_noise("randn")
            tk = torch.from_numpy(tok_host[:64].astype(np.int64)).to(dev).reshape(64, F, 7)
            cs = [tk[:, :, 0], tk[:, :, [1, 4]].reshape(64, 2 * F), tk[:, :, [2, 3, 5, 6]].reshape(64, 4 * F)]
            def one(b0, b1):
                y = ref_model.decode([c[b0:b1] for c in cs])[:, :, 2048:4096]
                return (y * 32767).to(torch.int16).cpu()
            for _ in range(3):
                one(0, 1)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            nrep = 32
            for i in range(nrep):
                one(i, i + 1)
            dt1 = (time.perf_counter() - t0) / nrep
            one(0, 64)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(3):
                one(0, 64)
            dt64 = (time.perf_counter() - t0) / 3
            gpu_torch = {"kind": "oracle port on stock PyTorch CUDA (cuDNN/ATen), not this repo's kernels",
                         "b1_ms_per_window": 1e3 * dt1, "b1_audio_s_per_s": AUDIO_S_PER_WINDOW / dt1,
                         "b64_ms_per_tick": 1e3 * dt64, "b64_audio_s_per_s": 64 * AUDIO_S_PER_WINDOW / dt64}
            try:  # the whole 1024-window tick as ONE stock-PyTorch batch (the reference never batches; upper bar for cuDNN)
                tk_all = torch.from_numpy(tok_host.astype(np.int64)).to(dev).reshape(S, F, 7)
                cs = [tk_all[:, :, 0], tk_all[:, :, [1, 4]].reshape(S, 2 * F), tk_all[:, :, [2, 3, 5, 6]].reshape(S, 4 * F)]
                one(0, S)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                one(0, S)
                dtS = time.perf_counter() - t0
                gpu_torch.update({f"b{S}_ms_per_tick": 1e3 * dtS, f"b{S}_audio_s_per_s": S * AUDIO_S_PER_WINDOW / dtS})
            except Exception as exc:  # noqa: BLE001
                gpu_torch[f"b{S}_error"] = repr(exc)[:200]
            del ref_model
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            gpu_torch = {"error": repr(exc)}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cw, ctimes, cores = cpu_reference_run(args.ref_windows, F, steps=50, warmup=1, budget_s=args.cpu_budget)
        cpu = {"value": cw * AUDIO_S_PER_WINDOW, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.ref_windows} windows per pass x {len(ctimes)} passes of the same tick, B=1 decode per "
                         f"window on torch CPU threads (reference semantics), {sum(ctimes):.1f} s",
               "windows_per_s": cw}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": tot_dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "streams_per_gpu": S, "frames_per_window": F,
                   "tokens_per_window": 7 * F, "weights": "random-init seed 0 variant w1", "noise": "in-kernel philox",
                   "precision": args.precision, "trim": not args.no_trim, "partition": "stream s -> rank s mod G (project_morpheus_b200.partition.partition_of)",
                   "l2": "256 MiB memset between steps (untimed); per-tick activations exceed L2"},
        "windows_per_s": wps, "realtime_factor_per_gpu": value / world,
        "tflops_reference_equivalent": wps * FLOP_PER_WINDOW_REFERENCE / 1e12,
        "tflops_cone": wps * FLOP_PER_WINDOW_CONE / 1e12,
        "e2e": {"value": pipe_value, "unit": UNIT, "h2d_bytes_per_step": int(tok_host.nbytes),
                "d2h_bytes_per_step": int(S * 2048 * 2 + S * 4), "ms_per_step": tot_pipe_ms / K,
                "api": "snacb_decode_windows_host_submit/_wait via SnacEngine.submit_windows/wait_windows, two ticks in "
                       "flight (host tokens -> H2D -> kernels -> D2H -> PCM + statuses on the host, every step)",
                "sync_value": e2e_value, "sync_ms_per_step": tot_e2e_ms / K,
                "sync_api": "snacb_decode_windows_host via SnacEngine.decode_windows, one blocking call per tick"},
        "gpu_launches": int(launches), "wall_s_device_region": wall_dev, "clocks": clocks,
        "sustained": sustained, "cfg4_strong_scaling": cfg4, "cfg5_partitioned": cfg5p,
        "roofline": roof, "cpu_baseline": cpu, "gpu_torch_baseline": gpu_torch, "latency": extra, "checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--streams", type=int, default=1024, help="concurrent streams (= windows per tick) per GPU")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--precision", choices=["fp16", "fp32"], default="fp16")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--persistent-ru", action="store_true")
    ap.add_argument("--no-trim", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--ref-windows", type=int, default=32, help="windows per step of the CPU reference sample")
    ap.add_argument("--latency-reps", type=int, default=200)
    ap.add_argument("--ingest-streams", type=int, default=1024, help="streams of the token-ingress measurement (0 = skip)")
    ap.add_argument("--ragged-streams", type=int, default=512, help="streams of the config-5 ragged-tick side measurement (0 = skip)")
    ap.add_argument("--long-read-batch", type=int, default=32, help="streams of the config-3 long_read side measurement (0 = skip)")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="seconds of back-to-back ticks for the sustained figure")
    ap.add_argument("--cfg4-streams", type=int, default=1024, help="total streams of the config-4 strong-scaling leg (0 = skip)")
    ap.add_argument("--cfg5-streams-per-gpu", type=int, default=512, help="config-5 whole-box leg, N > 1 only (0 = skip)")
    ap.add_argument("--pull-streams", type=str, default="64,1024", help="concurrent adapters of the ticker / pull() leg ('' = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
