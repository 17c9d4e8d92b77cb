#!/usr/bin/env python3
"""Benchmark of the HiFi-GAN synthesis hot path (BASELINE.json metric: audio-seconds synthesized
per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp16]

Workload (BASELINE.json configs[1]): hifigan_bn_tdnnf_wav2vec2_vq_48_v1-shaped generator
(Cin 504, 512->16 channels, x320), random-init weights (seed 0), one batch of 64 synthetic
utterances of 10-15 s, padded to the longest item exactly as the reference pipeline pads
(/root/reference/satools/satools/bin/pipeline.py:43-66).  One step = one generator forward over
the batch.  `value` counts the TRUE audio seconds of the 64 items (not the padding).

  value     forward with x resident in HBM (padded semantics: every item computed to 750 frames, as the reference does),
            CUDA events on the launching stream, exactly K steps; `sustained` = the same loop run for >= 5 s (the chip
            is power capped: the longer the run, the lower the clock)
  e2e       the same padded batch through the host-buffer C ABI (HostPipeline: two slots over
            sa_hifigan_synthesize_host_async): pinned host x -> H2D -> forward -> D2H of the fp32 waveform, every step.
            Same semantics as `value`.  Beside it: the ragged form (frames_per_item), the corpus driver's default form
            (ragged + trimmed PCM16 output) and the compact-conditioning form (VQ index input, 5 bytes per frame)
  roofline  tensor-pipe roofline of the whole conv chain + per-stage breakdown from a per-launch
            CUDA-event profile (sa_hifigan_get_profile)
  cpu_baseline  the torch-CPU port of the reference (oracle/hifigan_torch_cpu.py) on the same weights on a bounded
            sample of the same batch: all host threads, and 1 thread ("as shipped": importing satools sets
            torch.set_num_threads(1), /root/reference/satools/satools/hifigan/yaapt.py:27)
  gpu_eager_reference  the reference's op sequence on THIS GPU the way the reference runs it there: torch eager
            (cuDNN) under torch.autocast('cuda') (egs/vc/libritts/local/tuning/hifigan.py:99), same weights and batch
  extra     the other BASELINE.json configs: [2] bf16 + quant_16_awgn_2 batch 64, [3] one 60 s utterance (chunked,
            halo 20), [0] one 5 s utterance (latency: direct, CUDA graph, host entry)
  corpus    configs[4]: a synthetic LibriSpeech-length corpus through satools_b200.synth.synthesize_corpus (LPT
            sharding over the ranks, length buckets, pinned-slab staging threads, two-slot pipeline, trimmed PCM16):
            STRONG scaling (the corpus is fixed, ranks split it), time = max over ranks

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


def corpus_lengths(hours: float, seed: int = 20240):
    """LibriSpeech-like utterance lengths: log-normal, mean 12.3 s, clipped to 1-35 s; in frames."""
    rng = np.random.default_rng(seed)
    n = max(8, int(round(hours * 3600.0 / 12.3)))
    sec = np.clip(rng.lognormal(np.log(12.3) - 0.18, 0.6, n), 1.0, 35.0)
    return np.maximum(1, np.round(sec * FRAMES_PER_SEC)).astype(int).tolist()


def run_corpus(gen, dev, rank, world, hours, compact):
    """One pass of the corpus driver over this rank's shard.  The corpus references a pool of 192 distinct synthetic
    utterances (cut to length), so host memory stays small while every utterance is staged, copied, synthesized, copied
    back and handed to a sink like a real one."""
    from satools_b200 import conditioning, synth
    lengths = corpus_lengths(hours)
    rng = np.random.default_rng(99)
    cb = conditioning.codebook()
    pool = []
    for k in range(192):
        idx, f0, spk = conditioning.utterance_parts(rng, 35 * FRAMES_PER_SEC)
        pool.append((idx, f0, spk, None if compact else conditioning.assemble(idx, f0, spk, cb=cb)))
    feats = {}
    for i, n in enumerate(lengths):
        idx, f0, spk, x = pool[i % len(pool)]
        feats[f"utt{i:06d}"] = synth.VQFeatures(idx[:n], f0[:n], spk) if compact else x[:, :n]
    if compact:
        gen.set_codebook(torch.from_numpy(cb))
    acc = [0, 0]

    def sink(u, w):
        acc[0] += int(w[w.shape[0] // 2])
        acc[1] += w.shape[0]

    warm = {u: feats[u] for u in sorted(feats, key=lambda u: -synth._frames(feats[u]))[:96 * world]}   # the longest: sizes every slab
    synth.synthesize_corpus(gen, warm, rank=rank, world_size=world, sink=sink, device=dev)
    torch.cuda.synchronize()
    stats = {}
    t0 = time.perf_counter()
    synth.synthesize_corpus(gen, feats, rank=rank, world_size=world, sink=sink, stats=stats, device=dev)
    torch.cuda.synchronize()
    stats["seconds"] = time.perf_counter() - t0
    stats["samples"] = acc[1] // 2
    stats["audio_s_total"] = sum(lengths) / FRAMES_PER_SEC
    stats["n_utts_total"] = len(lengths)
    return stats


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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra configs, the eager-GPU leg and the corpus leg")
    ap.add_argument("--sustain-seconds", type=float, default=5.0)
    ap.add_argument("--corpus-hours", type=float, default=300.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch.distributed as dist
    from satools_b200 import CoreHifiGan, HostPipeline, conditioning
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

    def max_over_ranks(v):
        t = torch.tensor([float(v)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_forward(n, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            gen(x_dev, **kw)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n

    # ---- device-resident throughput: exactly K steps, then the same loop for >= sustain seconds ----------
    for _ in range(args.warmup):
        gen(x_dev)
    launches_per_step = gen.last_launch_count
    with ClockSampler(local) as clk:
        ms_per_step = timed_forward(args.steps)
    n_sustain = max(args.steps, int(np.ceil(args.sustain_seconds * 1e3 / ms_per_step)))
    with ClockSampler(local) as clk_sustain:
        sustain_ms = timed_forward(n_sustain)
    ragged_ms_per_step = timed_forward(args.steps, frames_per_item=frames)

    # ---- end to end through the host-buffer C-ABI entries -----------------------------------------------
    y_host = torch.empty((args.batch, 1, gen.output_length(x_np.shape[2])), dtype=torch.float32, pin_memory=True)
    gen.synthesize_host(x_host, out=y_host, device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gen.synthesize_host(x_host, out=y_host, device=dev)       # returns after the D2H completed
    torch.cuda.synchronize()
    sync_ms_per_step = 1e3 * max_over_ranks(time.perf_counter() - t0) / args.steps

    pipe = HostPipeline(gen, depth=2, device=dev)
    xs = [x_host, x_host.clone().pin_memory()]
    ys = [y_host, torch.empty_like(y_host).pin_memory()]
    n_trim = sum(gen.output_length(f) for f in frames)
    ys_pcm = [torch.empty(n_trim, dtype=torch.int16).pin_memory() for _ in range(2)]
    # compact conditioning of the same 64 utterances (regenerated with the same seeds: identical content)
    rngw = np.random.default_rng(1234 + 1 + 1000 * rank)
    T_pad = x_np.shape[2]
    idx_np = np.full((args.batch, T_pad), 255, dtype=np.uint8)
    f0_np = np.zeros((args.batch, T_pad), dtype=np.float32)
    spk_np = np.zeros(args.batch, dtype=np.int32)
    for b, n in enumerate(frames):
        i_, f_, s_ = conditioning.utterance_parts(rngw, n)
        idx_np[b, :n], f0_np[b, :n], spk_np[b] = i_, f_, s_
    gen.set_codebook(torch.from_numpy(conditioning.codebook()))
    vq_in = [tuple(torch.from_numpy(a.copy()).pin_memory() for a in (idx_np, f0_np, spk_np)) for _ in range(2)]
    checksum = 0.0

    def run_pipelined(n, mode):
        nonlocal checksum
        prev = None
        for k in range(n):
            if mode == "padded":
                t = pipe.submit(xs[k & 1], out=ys[k & 1])
            elif mode == "ragged":
                t = pipe.submit(xs[k & 1], out=ys[k & 1], frames_per_item=frames)
            elif mode == "trimmed_pcm16":
                t = pipe.submit(xs[k & 1], out=ys_pcm[k & 1], frames_per_item=frames, trimmed=True)
            else:
                t = pipe.submit_vq(*vq_in[k & 1], frames, out=ys_pcm[k & 1])
            if prev is not None:
                checksum += float(pipe.result(prev).view(-1)[1000])   # host read of the previous step's result
            prev = t
        checksum += float(pipe.result(prev).view(-1)[1000])

    def timed_pipeline(mode):
        run_pipelined(max(2, args.warmup), mode)
        barrier()
        h0, d0 = pipe.h2d_bytes, pipe.d2h_bytes
        t0 = time.perf_counter()
        run_pipelined(args.steps, mode)
        torch.cuda.synchronize()
        ms = 1e3 * max_over_ranks(time.perf_counter() - t0) / args.steps
        return ms, (pipe.h2d_bytes - h0) // args.steps, (pipe.d2h_bytes - d0) // args.steps

    e2e = {m: timed_pipeline(m) for m in ("padded", "ragged", "trimmed_pcm16", "vq_trimmed_pcm16")}

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
                "sustained_frac": round(flops_step / sustain_ms / peaks["tflops"], 4),
                "dominant_section": SECTION_NAMES[dominant], "sections": stages}

    extras, eager, cpu, corpus = None, None, None, None
    single = rank == 0 and world == 1

    # ---- the reference's own GPU path on this GPU: torch eager + cuDNN under autocast(fp16) ----------------
    if single and not args.no_extras:
        from oracle import hifigan_torch_cpu as otc
        p_cuda = {k: (w.to(dev), b.to(dev)) for k, (w, b) in otc.fold(state_cpu, torch.float32).items()}
        with torch.autocast("cuda", dtype=torch.float16):
            otc.generator_forward(p_cuda, x_dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                otc.generator_forward(p_cuda, x_dev)
            e1.record()
            torch.cuda.synchronize()
        eg_ms = e0.elapsed_time(e1) / 3
        eager = {"value": audio_s / (eg_ms / 1e3), "unit": "audio-s/s", "ms_per_step": eg_ms,
                 "what": "reference op sequence (torch eager, cuDNN convs, weight-norm folded once) under torch.autocast('cuda', fp16) "
                         "as hifigan.py:99 runs it; same weights, same padded batch, device-resident", "steps": 3}
        del p_cuda
        torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs ------------------------------------------------------------------
    if single and not args.no_extras:
        from satools_b200 import synth
        extras = {}
        # configs[2]: bf16 operands, fp32 accumulate, quant_16_awgn_2 conditioning, batch 64
        x3 = torch.from_numpy(conditioning.batch(4343, frames, pad_to=x_np.shape[2], f0_transformation="quant_16_awgn_2")).to(dev)
        gen.precision = "bf16"
        for _ in range(3):
            gen(x3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            gen(x3)
        e1.record()
        torch.cuda.synchronize()
        ms3 = e0.elapsed_time(e1) / args.steps
        extras["config2_bf16_quant_awgn_b64"] = {"value": audio_s / (ms3 / 1e3), "unit": "audio-s/s", "ms_per_step": ms3,
                                                 "precision": "bf16 operands, fp32 accumulate", "steps": args.steps}
        gen.precision = args.precision
        del x3
        # configs[0]: one 5 s utterance (latency path)
        x5 = torch.from_numpy(conditioning.batch(21, [250])).to(dev)
        x5h = torch.from_numpy(conditioning.batch(21, [250])).pin_memory()
        g5 = gen.graphed(1, 250)

        def lat(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            return 1e3 * float(np.median(ts))
        extras["config0_5s_latency"] = {"direct_ms": lat(lambda: gen(x5)), "cuda_graph_ms": lat(lambda: g5(x5)),
                                        "host_entry_ms": lat(lambda: gen.synthesize_host(x5h, device=dev)),
                                        "launches": g5.launches, "unit": "ms per 5 s utterance (median of 50)"}
        # configs[3]: one 60 s utterance, chunked synthesis with halo overlap vs one unchunked forward
        x60 = conditioning.batch(31, [3000])[0]
        x60d = torch.from_numpy(x60[None]).to(dev)
        chunked = lat(lambda: synth.synthesize_corpus(gen, {"long": x60}, chunk_frames=512, device=dev), n=7)
        extras["config3_60s_latency"] = {"chunked_host_ms": chunked, "unchunked_device_ms": lat(lambda: gen(x60d), n=10),
                                         "chunks": "6 windows of 512 + 2 x 20 halo frames in one batch, host features in, PCM16 out",
                                         "unit": "ms per 60 s utterance (median)"}
        del x60d
        # N2 (first step): the YAAPT front end (band-pass biquads + NLFER energy + voiced flags, yaapt.py:42-52,148-176) of the
        # same 64 utterances, waveforms of frames * 320 samples, options of bin/pipeline.py (frame_length 35, frame_space 20)
        from satools_b200 import yaapt_frontend as yf
        yopts = dict(frame_length=35.0, frame_space=20.0)
        lens = [f * 320 for f in frames]
        wav_h = torch.zeros(len(frames), max(lens)).pin_memory()
        for b, n_s in enumerate(lens):
            wav_h[b, :n_s] = torch.from_numpy(conditioning.waveform(500 + b, n_s / 16000.0)[:n_s])
        wav_d = wav_h.to(dev)

        def timed(fn, n):
            for _ in range(3):
                fn()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(n):
                fn()
            t1.record()
            torch.cuda.synchronize()
            return t0.elapsed_time(t1) / n

        def from_host():
            r = yf.nlfer(wav_h.to(dev, non_blocking=True), lengths=lens, **yopts)
            return r.energy.cpu(), r.vuv.cpu()
        ms_dev = timed(lambda: yf.nlfer(wav_d, lengths=lens, **yopts), args.steps)
        front = yf.nlfer(wav_d, lengths=lens, **yopts)
        ms_shc = timed(lambda: yf.spec_shc(front, lengths=lens, candidates=True, **yopts), args.steps)
        ms_track = timed(lambda: yf.spec_track(front, lengths=lens, **yopts), args.steps)
        full_opts = dict(yopts, nccf_thresh1=0.25, tda_frame_length=25.0)                    # bin/pipeline.py's _yaapt_opts
        ms_full = timed(lambda: yf.yaapt(wav_d, lengths=lens, **full_opts), args.steps)

        def full_from_host():
            return yf.yaapt(wav_h.to(dev, non_blocking=True), lengths=lens, **full_opts).cpu()
        ms_full_host = timed(full_from_host, args.steps)
        ms_host = timed(from_host, args.steps)
        from oracle import yaapt_nlfer_numpy as onp
        t0 = time.perf_counter()
        n_cpu = 0
        tpar = onp.track_params(**full_opts)
        while time.perf_counter() - t0 < 8.0 and n_cpu < len(lens):
            w_cpu = wav_h[n_cpu, :lens[n_cpu]].numpy()
            o_cpu = onp.nlfer(w_cpu, onp.params(**yopts))
            sp_cpu, sd_cpu = onp.spec_track_finish(*onp.spec_candidates(onp.shc(o_cpu["filtered_nl"], o_cpu["vuv"], onp.params(**yopts)),
                                                                        o_cpu["vuv"], onp.params(**yopts)), onp.params(**yopts))
            t1_cpu = onp.time_track(o_cpu["filtered"], sp_cpu, sd_cpu, tpar)
            t2_cpu = onp.time_track(o_cpu["filtered_nl"], sp_cpu, sd_cpu, tpar)
            r_cpu = onp.refine(t1_cpu[0], t1_cpu[1], t2_cpu[0], t2_cpu[1], sp_cpu, o_cpu["energy"], o_cpu["vuv"], tpar)
            onp.dynamic(r_cpu[0], r_cpu[1], o_cpu["energy"].astype(np.float32), tpar)
            n_cpu += 1
        dt = time.perf_counter() - t0
        cpu_sec = sum(lens[i] for i in range(n_cpu)) / 16000.0
        extras["yaapt_frontend_b64"] = {
            "value": audio_s / (ms_dev / 1e3), "unit": "audio-s/s", "ms_per_step": ms_dev,
            "host_buffers_value": audio_s / (ms_host / 1e3), "host_buffers_ms_per_step": ms_host,
            "spec_shc_ms_per_step": ms_shc, "spec_track_ms_per_step": ms_track,
            "with_spec_track_value": audio_s / ((ms_dev + ms_track) / 1e3),
            "yaapt_value": audio_s / (ms_full / 1e3), "yaapt_ms_per_step": ms_full,
            "yaapt_host_buffers_value": audio_s / (ms_full_host / 1e3), "yaapt_host_buffers_ms_per_step": ms_full_host,
            "voiced_frames": int(front.vuv.sum()), "frames": int(sum(front.nframes)),
            "h2d_bytes_per_step": int(wav_h.numel() * 4), "d2h_bytes_per_step": int(len(lens) * yf.num_frames(max(lens), **yopts) * 5),
            "what": "SignalObj.filtered of the signal and the squared signal + PitchObj.energy / vuv / mean_energy for the batch "
                    "(the part of _yaapt before spec_track); spec_shc = the SHC vector of every voiced frame and the candidates peaks() picks from it "
                    "(spec_track's per-frame loop); spec_track = that + its per-utterance DP / smoothing / re-sampling = "
                    "spec_pitch, pitch_std as the reference's spec_track returns them; yaapt = the whole extractor "
                    "(front end + spec_track + time_track x 2 + refine + dynamic): waveforms in, final pitch per frame out",
            "cpu_port": {"value": cpu_sec / dt, "unit": "audio-s/s", "cores": 1, "kind": "port",
                         "sample": f"{n_cpu} utterances through oracle/yaapt_nlfer_numpy.py (the whole extractor; float64 filters, scipy recursion, "
                                   f"numpy rfft) in {dt:.1f} s; compare with yaapt_value; the reference's own TorchScript yaapt runs at ~20 audio-s/s per core (BASELINE.md)"}}
        # N3 (last step): nearest-codeword assignment of the batch's 64 x 750 bottleneck rows (chain/nn.py:402-477)
        rng_v = np.random.default_rng(9)
        cb_v = conditioning.codebook()
        bn_d = torch.from_numpy((cb_v[rng_v.integers(0, cb_v.shape[0], size=(len(frames), max(frames)))]
                                 + 0.5 * rng_v.standard_normal((len(frames), max(frames), cb_v.shape[1]))).astype(np.float32)).to(dev)
        gen.set_codebook(torch.from_numpy(cb_v))
        ms_vq = timed(lambda: gen.vq_assign(bn_d), max(args.steps, 20))
        vq_bytes = bn_d.numel() * 4 + bn_d.numel() // cb_v.shape[1]
        vq_rows = int(bn_d.numel() // cb_v.shape[1])
        vq_flop = 2.0 * vq_rows * cb_v.shape[1] * cb_v.shape[0]
        extras["vq_assign_b64"] = {"ms_per_step": ms_vq, "rows": vq_rows, "algorithmic_bytes": int(vq_bytes),
                                   "achieved_gbs": vq_bytes / (ms_vq * 1e-3) / 1e9, "hbm_peak_gbs": load_peaks()["hbm_gbs"],
                                   "hbm_frac": vq_bytes / (ms_vq * 1e-3) / 1e9 / load_peaks()["hbm_gbs"],
                                   "achieved_fp32_tflops": vq_flop / (ms_vq * 1e-3) / 1e12,
                                   "what": "sa_hifigan_vq_assign: fp32 [64 x 750, 256] rows in, uint8 code index out (48 codes), "
                                           "through CoreHifiGan.vq_assign; 24.6 kFLOP per 1 KB row in exact fp32 puts the step on the "
                                           "fp32 FMA pipe (ncu: profiles/r2_vq_assign_summary.txt), not on HBM"}
        del bn_d
        del wav_d

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------
    if single and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        fwd = cpu_port_runner(state_cpu)
        sample = torch.from_numpy(np.ascontiguousarray(x_np[:2, :, :max(frames[:2])]))

        def cpu_rate(threads, budget_s, max_runs):
            torch.set_num_threads(threads)
            x_s = sample if threads > 1 else sample[:1, :, :frames[0]]
            sec = (sum(frames[:2]) if threads > 1 else frames[0]) / FRAMES_PER_SEC
            fwd(x_s[:, :, :100])
            t0 = time.perf_counter()
            n = 0
            while True:
                fwd(x_s)
                n += 1
                if time.perf_counter() - t0 > budget_s or n >= max_runs:
                    break
            dt = time.perf_counter() - t0
            return n * sec / dt, n, dt
        v, n, dt = cpu_rate(cores, 10.0, 8)
        v1, n1, dt1 = cpu_rate(1, 8.0, 2)
        torch.set_num_threads(cores)
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"first 2 utterances of the batch x {n} runs ({dt:.1f} s), torch-CPU fp32 port of the reference",
               "as_shipped_1thread": {"value": v1, "cores": 1,
                                      "sample": f"first utterance of the batch x {n1} runs ({dt1:.1f} s); importing satools sets "
                                                "torch.set_num_threads(1) (hifigan/yaapt.py:27)"}}

    # ---- configs[4]: the corpus driver, strong scaling over the ranks ---------------------------------------
    if not args.no_extras:
        out_c = {}
        for name, compact in (("dense_x", False), ("vq_index", True)):
            st = run_corpus(gen, dev, rank, world, args.corpus_hours, compact)
            sec = max_over_ranks(st["seconds"])
            per_rank = [st["seconds"]]
            loads = [st["frames"]]
            if world > 1:
                g = [None] * world
                dist.all_gather_object(g, (st["seconds"], st["frames"]))
                per_rank, loads = [a for a, _ in g], [b for _, b in g]
            out_c[name] = {"value": st["audio_s_total"] / sec, "unit": "audio-s/s", "seconds": sec, "scaling": "strong",
                           "corpus_hours": st["audio_s_total"] / 3600.0, "utterances": st["n_utts_total"],
                           "per_rank_seconds": [round(v, 3) for v in per_rank],
                           "load_imbalance": round(max(loads) / (sum(loads) / len(loads)) - 1.0, 5),
                           "rank0": {"batches": st["batches"], "padding_waste": round(1.0 - st["frames"] / st["padded_frames"], 4),
                                     "host_stage_ms_per_batch": round(1e3 * st["stage_s"] / st["batches"], 3),
                                     "exposed_stage_wait_ms_per_batch": round(1e3 * st["stage_wait_s"] / st["batches"], 3),
                                     "collect_ms_per_batch": round(1e3 * st["collect_s"] / st["batches"], 3),
                                     "gpu_wait_ms_per_batch": round(1e3 * st["gpu_wait_s"] / st["batches"], 3),
                                     "h2d_bytes_per_audio_s": round(st["h2d_bytes"] * FRAMES_PER_SEC / st["frames"], 1),
                                     "d2h_bytes_per_audio_s": round(st["d2h_bytes"] * FRAMES_PER_SEC / st["frames"], 1)}}
        corpus = dict(out_c, api="satools_b200.synth.synthesize_corpus: scheduler.shard (LPT) -> scheduler.batches -> pinned-slab "
                                 "staging threads -> HostPipeline (trimmed PCM16 D2H) -> sink; host features in, host waveforms out",
                      length_law="log-normal, mean 12.3 s, clipped to 1-35 s")

    if rank == 0:
        value = world * audio_s / (ms_per_step / 1e3)
        pad_ms, pad_h2d, pad_d2h = e2e["padded"]
        out = {
            "metric": "hifigan_audio_seconds_per_second", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": dict(workload_config(args, "B200"), precision=args.precision,
                           accumulate="fp32", audio_s_per_step_per_gpu=audio_s, padded_audio_s_per_step_per_gpu=padded_audio_s),
            "clocks": clk.summary(),
            "sustained": {"value": world * audio_s / (sustain_ms / 1e3), "unit": "audio-s/s", "steps": n_sustain,
                          "ms_per_step": sustain_ms, "seconds": n_sustain * sustain_ms / 1e3, "clocks": clk_sustain.summary(),
                          "note": "the same padded forward loop run for >= 5 s: the power-capped steady state"},
            "e2e": {"value": world * audio_s / (pad_ms / 1e3), "unit": "audio-s/s", "ms_per_step": pad_ms,
                    "h2d_bytes_per_step": int(pad_h2d), "d2h_bytes_per_step": int(pad_d2h),
                    "api": "HostPipeline over sa_hifigan_synthesize_host_async (pinned host buffers, two slots): padded semantics, "
                           "fp32 [64,504,750] in, fp32 [64,1,240001] out -- the same work as `value`",
                    "ragged_value": world * audio_s / (e2e["ragged"][0] / 1e3), "ragged_ms_per_step": e2e["ragged"][0],
                    "trimmed_pcm16_value": world * audio_s / (e2e["trimmed_pcm16"][0] / 1e3),
                    "trimmed_pcm16_ms_per_step": e2e["trimmed_pcm16"][0],
                    "trimmed_pcm16_d2h_bytes_per_step": int(e2e["trimmed_pcm16"][2]),
                    "vq_trimmed_pcm16_value": world * audio_s / (e2e["vq_trimmed_pcm16"][0] / 1e3),
                    "vq_trimmed_pcm16_ms_per_step": e2e["vq_trimmed_pcm16"][0],
                    "vq_trimmed_pcm16_h2d_bytes_per_step": int(e2e["vq_trimmed_pcm16"][1]),
                    "vq_trimmed_pcm16_d2h_bytes_per_step": int(e2e["vq_trimmed_pcm16"][2]),
                    "variants": "ragged: frames_per_item passed (tiles past an item's true length + 24 frames are skipped; kept samples "
                                "bit-identical, test_ragged_batch_*); trimmed_pcm16: + only the kept samples come back, as int16 "
                                "(the corpus driver's default, pipeline.py:156-160); vq: + VQ index / F0 / speaker id in (N1)",
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
            "gpu_eager_reference": eager,
            "extra": extras,
            "corpus": corpus,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
