#!/usr/bin/env python3
"""Benchmark of the HiFi-GAN synthesis hot path (BASELINE.json metric: audio-seconds synthesized
per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp16]

Workload (BASELINE.json configs[1]): hifigan_bn_tdnnf_wav2vec2_vq_48_v1-shaped generator
(Cin 504, 512->16 channels, x320), random-init weights (seed 0), one batch of 64 synthetic
utterances of 10-15 s, padded to the longest item exactly as the reference pipeline pads
(/root/reference/satools/satools/bin/pipeline.py:43-66).  One step = one generator forward over
the batch.  `value` counts the TRUE audio seconds of the 64 items (not the padding).

  value     forward with x resident in HBM, CUDA events on the launching stream
  e2e       the same batch through the host-buffer C ABI as the corpus driver calls it (HostPipeline: two
            slots over sa_hifigan_synthesize_host_async): pinned host x -> H2D -> forward -> D2H of the
            fp32 waveform, every step; the blocking one-call-per-step form is reported beside it
  roofline  tensor-pipe roofline of the whole conv chain + per-stage breakdown from a per-launch
            CUDA-event profile (sa_hifigan_get_profile)
  cpu_baseline  the torch-CPU port of the reference (oracle/hifigan_torch_cpu.py) on the same
            weights, all host threads, on a bounded sample of the same batch

Multi-GPU: launched under torchrun, one rank per GPU; every rank synthesizes its own batch of 64
(weak scaling, utterances sharded, no collective on the data path); a barrier and a max over
ranks bracket the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "sa-toolkit_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

FRAMES_PER_SEC = 50
GFLOP_PER_AUDIO_S = 16.172              # 161,717,248 MAC/frame * 2 * 50
# per-section algorithmic GFLOP per audio second: conv_pre, stages 0-4, tail
SECTION_GFLOP = [0.1806, 4.2729, 4.2598, 4.2598, 2.1299, 1.0650, 0.0036]
SECTION_NAMES = ["conv_pre", "stage0_c256", "stage1_c128", "stage2_c64", "stage3_c32", "stage4_c16", "tail"]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                "tflops_burst": float(p.get("bf16_tflops", 1590.0)),
                "hbm_gbs": float(p.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """DRAM bytes of one forward of this workload from the newest committed ncu launch list (dram__bytes_read.sum +
    dram__bytes_write.sum summed over the launches of one step); None when no summary is present."""
    import glob
    import re
    best = None
    for path in glob.glob(os.path.join(ROOT, "profiles", "r*_launches_fp16_v*_summary.txt")):
        m = re.search(r"r(\d+)_launches_fp16_v(\d+)_summary", path)
        if m and (best is None or (int(m.group(1)), int(m.group(2))) > best[0]):
            best = ((int(m.group(1)), int(m.group(2))), path)
    try:
        m = re.search(r"DRAM traffic ([0-9.]+) GB", open(best[1]).read())
        return {"traffic": float(m.group(1)) * 1e9, "traffic_unit": "bytes per step (ncu, all launches of one forward)",
                "traffic_source": os.path.relpath(best[1], ROOT)}
    except Exception:
        return {"traffic": None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        hi = sorted(sm)[len(sm) // 2:]          # samples under load = upper half
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


def make_workload(rank: int, batch: int):
    from satools_b200 import conditioning
    rng = np.random.default_rng(1234 + 1)
    frames = rng.integers(10 * FRAMES_PER_SEC, 15 * FRAMES_PER_SEC + 1, size=batch).tolist()
    x = conditioning.batch(1234 + 1 + 1000 * rank, frames, pad_to=15 * FRAMES_PER_SEC)
    return frames, x


def cpu_port_runner(state):
    from oracle import hifigan_torch_cpu as otc
    p = otc.fold(state)
    return lambda x: otc.generator_forward(p, x)


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path (torch-CPU port of
    archi.py:77-91; the reference is Python and /root/reference does not travel to the GPU box),
    all host threads, each step a bounded sample (2 utterances) of the same batch."""
    if rank != 0:
        return
    from satools_b200 import CoreHifiGan
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    gen = CoreHifiGan(imput_dim=504)
    fwd = cpu_port_runner(gen.state_dict())
    frames, x = make_workload(0, args.batch)
    per_step = 2
    t_audio = 0.0
    for s in range(args.warmup):
        fwd(torch.from_numpy(x[:1, :, :100]))
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = (s * per_step) % args.batch
        idx = [(lo + i) % args.batch for i in range(per_step)]
        T = max(frames[i] for i in idx)
        fwd(torch.from_numpy(np.ascontiguousarray(x[idx, :, :T])))
        t_audio += sum(frames[i] for i in idx) / FRAMES_PER_SEC
    dt = time.perf_counter() - t0
    v = t_audio / dt
    sample = f"{per_step} utterances of the 64-item batch per step, {args.steps} steps, fp32, torch {torch.__version__} CPU"
    emit({
        "impl": "reference", "metric": "hifigan_audio_seconds_per_second", "value": v, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "cpu"),
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def workload_config(args, where):
    return {"workload": "configs[1]: hifigan_bn_tdnnf_wav2vec2_vq_48_v1-shaped generator, batch of 64 synthetic 10-15 s utterances",
            "batch": args.batch, "frames_padded": 15 * FRAMES_PER_SEC, "input": "[64,504,750] fp32",
            "output": "[64,1,240001] fp32", "weights": "reference random init, seed 0",
            "padding": "items padded to 750 frames as the reference pipeline does; value counts true audio only",
            "l2": "inputs + activations (GBs) exceed the 126 MB L2; no flush needed",
            "parallelism": f"utterance sharding x{args.gpus}, no collective", "where": where}


def main():
    # Only the final JSON line may reach stdout: NCCL / torch print banners from C code, so fd 1 is
    # pointed at stderr until the result is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SATOOLS_B200_PRECISION", "fp16"))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch.distributed as dist
    from satools_b200 import CoreHifiGan
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    gen = CoreHifiGan(imput_dim=504, precision=args.precision)
    state_cpu = {k: v.clone() for k, v in gen.state_dict().items()}
    gen = gen.to(dev)
    frames, x_np = make_workload(rank, args.batch)
    audio_s = sum(frames) / FRAMES_PER_SEC
    padded_audio_s = args.batch * x_np.shape[2] / FRAMES_PER_SEC
    x_host = torch.from_numpy(x_np).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------
    for _ in range(args.warmup):
        gen(x_dev)
    launches_per_step = gen.last_launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            gen(x_dev)
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps

    # ---- the same forward told the true lengths (frames_per_item): tiles past an item's kept samples are skipped ----
    for _ in range(2):
        gen(x_dev, frames_per_item=frames)
    barrier()
    e0.record()
    for _ in range(args.steps):
        gen(x_dev, frames_per_item=frames)
    e1.record()
    barrier()
    rms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(rms, op=dist.ReduceOp.MAX)
    ragged_ms_per_step = float(rms.item()) / args.steps

    # ---- end to end through the host-buffer C-ABI entry --------------------------------------
    y_host = torch.empty((args.batch, 1, gen.output_length(x_np.shape[2])), dtype=torch.float32, pin_memory=True)
    gen.synthesize_host(x_host, out=y_host, device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gen.synthesize_host(x_host, out=y_host, device=dev)       # returns after the D2H completed
    torch.cuda.synchronize()
    sync_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(sync_s, op=dist.ReduceOp.MAX)
    sync_ms_per_step = 1e3 * float(sync_s.item()) / args.steps

    # The corpus driver's call pattern (satools_b200/synth.py): batches go through HostPipeline, two slots over
    # sa_hifigan_synthesize_host_async, so the H2D / D2H copies of one batch run under the kernels of its neighbour.
    # Every step still copies its own input from pinned host memory and reads its own waveform back on the host.
    from satools_b200 import HostPipeline
    pipe = HostPipeline(gen, depth=2, device=dev)
    xs = [x_host, x_host.clone().pin_memory()]
    ys = [y_host, torch.empty_like(y_host).pin_memory()]
    checksum = 0.0

    def run_pipelined(n, fpi):
        nonlocal checksum
        prev = None
        for k in range(n):
            t = pipe.submit(xs[k & 1], out=ys[k & 1], frames_per_item=fpi)
            if prev is not None:
                checksum += float(pipe.result(prev)[0, 0, 1000])   # host read of the previous step's result
            prev = t
        checksum += float(pipe.result(prev)[0, 0, 1000])

    def timed_pipeline(fpi):
        run_pipelined(max(2, args.warmup), fpi)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps, fpi)
        torch.cuda.synchronize()
        sec = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return 1e3 * float(sec.item()) / args.steps

    e2e_padded_ms_per_step = timed_pipeline(None)          # padded semantics: every item computed to 750 frames
    e2e_ms_per_step = timed_pipeline(frames)               # what synth.synthesize_corpus does: true lengths passed along

    # ---- per-launch profile -> per-section roofline --------------------------------------
    peaks = load_peaks()
    prof = gen.profile(x_dev, repeats=2)
    n_sections = len(SECTION_NAMES)
    sec_ms = [0.0] * n_sections
    for tag, t in prof:
        sec_ms[min(max(tag, 0) // 16, n_sections - 1)] += t
    stages = []
    for name, gf, t in zip(SECTION_NAMES, SECTION_GFLOP, sec_ms):
        ach = gf * padded_audio_s / t if t > 0 else 0.0          # GFLOP / ms = TFLOP/s
        stages.append({"section": name, "ms": round(t, 4), "achieved_tflops": round(ach, 2),
                       "frac": round(ach / peaks["tflops"], 4)})
    conv_ms = sum(sec_ms)
    flops_step = GFLOP_PER_AUDIO_S * padded_audio_s               # GFLOP of one (padded) step
    achieved = flops_step / ms_per_step                           # TFLOP/s over the timed region
    dominant = max(range(n_sections), key=lambda i: sec_ms[i])
    roofline = {"bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["tflops"], 4), **ncu_traffic(),
                "peak_source": peaks["source"] + ", sustained bf16 (kernels timed inside a long step)",
                "kernel": "whole conv chain of one forward (conv_pre + 5 x (upsampler + 18 ResBlock convs) + tail)",
                "algorithmic_gflop_per_step": round(flops_step, 1), "profiled_ms_per_step": round(conv_ms, 3),
                "dominant_section": SECTION_NAMES[dominant], "sections": stages}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        fwd = cpu_port_runner(state_cpu)
        sample = torch.from_numpy(np.ascontiguousarray(x_np[:2, :, :max(frames[:2])]))
        fwd(sample[:, :, :100])
        t0 = time.perf_counter()
        n = 0
        while True:
            fwd(sample)
            n += 1
            if time.perf_counter() - t0 > 10.0 or n >= 8:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": n * sum(frames[:2]) / FRAMES_PER_SEC / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"first 2 utterances of the batch x {n} runs ({dt:.1f} s), torch-CPU fp32 port of the reference"}

    if rank == 0:
        value = world * audio_s / (ms_per_step / 1e3)
        out = {
            "metric": "hifigan_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": dict(workload_config(args, "B200"), precision=args.precision,
                           accumulate="fp32", audio_s_per_step_per_gpu=audio_s, padded_audio_s_per_step_per_gpu=padded_audio_s),
            "clocks": clk.summary(),
            "e2e": {"value": world * audio_s / (e2e_ms_per_step / 1e3), "unit": "audio-s/s",
                    "ms_per_step": e2e_ms_per_step, "h2d_bytes_per_step": int(x_host.numel() * 4),
                    "d2h_bytes_per_step": int(y_host.numel() * 4), "api": "HostPipeline over sa_hifigan_synthesize_host_async (pinned host buffers, two slots), frames_per_item "
                           "passed as satools_b200.synth does: tiles past an item's true length + 24 frames are skipped, the "
                           "kept samples are bit-identical to the padded run (test_ragged_batch_*)",
                    "padded_value": world * audio_s / (e2e_padded_ms_per_step / 1e3),
                    "padded_ms_per_step": e2e_padded_ms_per_step,
                    "single_call_value": world * audio_s / (sync_ms_per_step / 1e3),
                    "single_call_ms_per_step": sync_ms_per_step,
                    "single_call_api": "sa_hifigan_synthesize_host (one blocking call per step)"},
            "ragged": {"value": world * audio_s / (ragged_ms_per_step / 1e3), "unit": "audio-s/s",
                       "ms_per_step": ragged_ms_per_step,
                       "note": "device-resident forward with frames_per_item (same batch, same kept samples); "
                               "`value` above is the padded forward"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
