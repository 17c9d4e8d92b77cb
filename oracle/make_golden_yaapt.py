#!/usr/bin/env python3
"""Mint tests/golden/yaapt_nlfer.npz by RUNNING THE REFERENCE's YAAPT front end (build container only).

    SA_JIT_TWEAK=true PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_yaapt.py

Imports /root/reference/satools (read-only) and executes, on synthetic waveforms, exactly what `_yaapt` does before its
spectral tracker (satools/satools/hifigan/yaapt.py:873-899): padding, SignalObj / squared SignalObj, `filtered_version`
(torchaudio biquads), PitchObj, `nlfer`.  Recorded per case: the waveform, the two filtered signals, the normalised NLFER
energy, the voiced flags and the mean energy.  The cases use the options the reference pipeline passes
(`bin/pipeline.py`: frame_length 35, frame_space 20) and the defaults (frame_space 10).
"""
import os
import sys
from math import floor

os.environ.setdefault("SA_JIT_TWEAK", "true")
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/satools")
sys.path.insert(0, os.path.join(ROOT, "sa-toolkit_b200"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import satools  # noqa: E402,F401  (the reference)
from satools.hifigan import yaapt as ref  # noqa: E402

from satools_b200 import conditioning  # noqa: E402
from oracle import yaapt_nlfer_numpy as onp  # noqa: E402

PIPELINE = dict(frame_length=35.0, frame_space=20.0, nccf_thresh1=0.25, tda_frame_length=25.0)   # bin/pipeline.py's _yaapt_opts
CASES = [  # (seed, seconds, options)
    (0, 1.5, PIPELINE),
    (1, 2.3, PIPELINE),
    (2, 0.9, dict()),
    (3, 0.05, PIPELINE),       # shorter than two frames
]


def run_reference(wav: np.ndarray, opts: dict):
    p = onp.params(**{k: v for k, v in opts.items() if k in onp.DEFAULTS})
    x = torch.from_numpy(wav).reshape(1, -1)
    to_pad = int(p["frame_length"] / 1000 * int(p["sr"])) // 2
    x = torch.nn.functional.pad(x.squeeze(), (to_pad, to_pad))
    signal = ref.SignalObj(x, p["sr"])
    nonlinear = ref.SignalObj(signal.data ** 2, p["sr"])
    signal.filtered_version(p)
    nonlinear.filtered_version(p)
    frame_size = floor(torch.tensor(p["frame_length"] * signal.fs / 1000))
    frame_jump = floor(torch.tensor(p["frame_space"] * signal.fs / 1000))
    pitch = ref.PitchObj(int(frame_size), int(frame_jump), int(p["fft_length"]))
    ref.nlfer(signal, pitch, p)
    # spec_track's SHC vectors: run the reference's own spec_track and record what it hands to `peaks` per voiced frame
    full = dict(p)
    full.update(dict(nlfer_thresh2=0.1, shc_maxpeaks=4.0, shc_thresh1=5.0, shc_thresh2=1.25, f0_double=150.0, f0_half=150.0,
                     dp5_k1=11.0, nccf_thresh1=0.3, nccf_thresh2=0.9, nccf_maxcands=3.0, nccf_pwidth=5.0, merit_boost=0.2,
                     merit_pivot=0.99, merit_extra=0.4, median_value=7.0, dp_w1=0.15, dp_w2=0.5, dp_w3=0.1, dp_w4=0.9,
                     spec_pitch_min_std=0.05, tda_frame_length=35.0))               # the defaults of _yaapt (yaapt.py:818-866)
    full.update(opts)
    captured = []
    original = ref.peaks

    returned = []

    def spy(shc, delta, maxpeaks, parameters):
        captured.append(shc.clone().numpy())
        out = original(shc, delta, maxpeaks, parameters)
        returned.append((out[0].clone().numpy(), out[1].clone().numpy()))
        return out
    ref.peaks = spy
    spec_pitch, pitch_std = None, None
    try:
        spec_pitch, pitch_std = ref.spec_track(nonlinear, pitch, full)
    except IndexError:
        # fewer than four frames: the reference's spec_track raises at `spec_pitch[1] = spec_pitch[3]` (yaapt.py:311), AFTER
        # the per-frame loop -- the SHC vectors it computed up to there are recorded all the same
        pass
    finally:
        ref.peaks = original
    shc = np.zeros((int(pitch.nframes), len(captured[0]) if captured else 0), dtype=np.float32)
    shc[np.nonzero(pitch.vuv.numpy())[0]] = np.stack(captured) if captured else 0
    voiced = np.nonzero(pitch.vuv.numpy())[0]
    cand_pitch = np.zeros((4, int(pitch.nframes)), dtype=np.float32)
    cand_merit = np.ones((4, int(pitch.nframes)), dtype=np.float32)
    for f, (cp, cm) in zip(voiced, returned):
        cand_pitch[:, f], cand_merit[:, f] = cp, cm
    extra = {}
    if spec_pitch is not None:
        tp1, tm1 = ref.time_track(signal, spec_pitch.clone(), pitch_std, pitch, full)
        tp2, tm2 = ref.time_track(nonlinear, spec_pitch.clone(), pitch_std, pitch, full)
        extra = dict(time_pitch1=tp1.numpy(), time_merit1=tm1.numpy(), time_pitch2=tp2.numpy(), time_merit2=tm2.numpy())
        final = ref.yaapt(torch.from_numpy(wav).reshape(1, -1), {k: float(v) for k, v in opts.items()})
        extra["final_pitch"] = final[0].numpy().astype(np.float32)
    spec_pitch = np.zeros(0, dtype=np.float32) if spec_pitch is None else spec_pitch.numpy().astype(np.float32)
    pitch_std = np.float32(np.nan) if pitch_std is None else np.float32(float(pitch_std))
    return dict(**extra, spec_pitch=spec_pitch, pitch_std=pitch_std, cand_pitch=cand_pitch, cand_merit=cand_merit, shc=shc, filtered=signal.filtered.numpy(), filtered_nl=nonlinear.filtered.numpy(), energy=pitch.energy.numpy(),
                vuv=pitch.vuv.numpy(), mean_energy=np.float32(pitch.mean_energy.item()), nframes=np.int64(pitch.nframes))


def main():
    out = {}
    for i, (seed, seconds, opts) in enumerate(CASES):
        wav = conditioning.waveform(seed, seconds)
        r = run_reference(wav, opts)
        fopts = {k: v for k, v in opts.items() if k in onp.DEFAULTS}
        o = onp.nlfer(wav, onp.params(**fopts))
        rel = np.abs(o["energy"] - r["energy"]).max() / max(1e-12, np.abs(r["energy"]).max())
        flips = int((o["vuv"] != r["vuv"]).sum())
        so = onp.shc(r["filtered_nl"], r["vuv"], onp.params(**fopts))
        cp, cm = onp.spec_candidates(so, r["vuv"], onp.params(**fopts))
        same = (cp == r["cand_pitch"]).all(0)
        print(f"        peaks on the oracle's SHC: candidate pitches identical in {int(same.sum())} of {len(same)} frames, "
              f"merit max err {np.abs(cm - r['cand_merit'])[:, same].max():.2e}")
        if len(r["spec_pitch"]):
            sp, sd = onp.spec_track_finish(r["cand_pitch"], r["cand_merit"], onp.params(**fopts))
            print(f"        spec_track finish on the reference's candidates: spec_pitch max err {np.abs(sp - r['spec_pitch']).max():.2e} Hz, "
                  f"pitch_std {float(sd):.6f} vs {float(r['pitch_std']):.6f}")
        if "time_pitch1" in r:
            tp = onp.track_params(**opts)
            for tag, sig in (("1", r["filtered"]), ("2", r["filtered_nl"])):
                op, om = onp.time_track(sig, r["spec_pitch"], r["pitch_std"], tp)
                same = (op == r["time_pitch" + tag]).all(0)
                print(f"        time_track{tag}: {op.shape} pitches identical in {int(same.sum())} of {len(same)} frames, merit max err "
                      f"{np.abs(om - r['time_merit' + tag])[:, same].max():.2e}; final_pitch {r['final_pitch'].shape}")
        if "time_pitch1" in r:
            rp, rm = onp.refine(r["time_pitch1"], r["time_merit1"], r["time_pitch2"], r["time_merit2"], r["spec_pitch"], r["energy"], r["vuv"], tp)
            fin = onp.dynamic(rp, rm, r["energy"], tp)
            print(f"        refine + dynamic on the reference's tracks: final pitch max err {np.abs(fin - r['final_pitch']).max():.2e} Hz "
                  f"({int((r['final_pitch'] > 0).sum())} voiced of {len(fin)})")
        print(f"        SHC peak {r['shc'].max():.3e}, oracle-vs-reference max err / peak {np.abs(so - r['shc']).max() / r['shc'].max():.2e}")
        print(f"case {i}: n={len(wav)} frames={int(r['nframes'])} voiced={int(r['vuv'].sum())} oracle-vs-reference energy rel-err {rel:.2e}, "
              f"vuv flips {flips}, filtered max-abs {np.abs(r['filtered']).max():.3e} err {np.abs(o['filtered'] - r['filtered']).max():.2e}")
        out[f"c{i}_seed"] = np.int64(seed)
        out[f"c{i}_seconds"] = np.float64(seconds)
        out[f"c{i}_opts"] = np.array([opts.get("frame_length", 35.0), opts.get("frame_space", 10.0)])
        out[f"c{i}_track_opts"] = np.array([opts.get("tda_frame_length", 35.0), opts.get("nccf_thresh1", 0.3)])
        for k, v in r.items():
            out[f"c{i}_{k}"] = v
    out["n_cases"] = np.int64(len(CASES))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "yaapt_nlfer.npz"), **out)


if __name__ == "__main__":
    main()
