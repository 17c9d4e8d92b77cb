"""YAAPT front end (first step): band-pass biquads, NLFER energy, voiced flags.

CPU tests: the numpy oracle against the outputs of the reference itself (tests/golden/yaapt_nlfer.npz, minted by
oracle/make_golden_yaapt.py from satools/hifigan/yaapt.py), the C ABI symbol table, the frame geometry.
GPU tests: the CUDA path (csrc/yaapt_frontend.cu through include/sa_yaapt.h) against the same fixtures, against the oracle at
batch sizes the fixtures do not cover, and through size-independent properties.

Tolerances.  The reference evaluates the two biquads in float32; the low-pass at 50 Hz has its poles at radius 0.986 and the
high-pass at 1500 Hz then removes all but ~1e-3 of its output, so the reference's own rounding noise on `filtered` is ~5e-5
of its peak (measured: float64 restatement vs reference, make_golden_yaapt.py).  Both the oracle and the CUDA path run the
recursion in double; they are held to the reference within 5e-4 of the peak for `filtered`, 1e-4 for the normalised NLFER
energy (measured 4e-6), identical voiced flags except for frames whose energy is within 1e-3 of the threshold.
"""
import os
import re

import numpy as np
import pytest
import torch

import helpers
from oracle import yaapt_nlfer_numpy as onp
from satools_b200 import _lib, conditioning

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden", "yaapt_nlfer.npz")
HEADER = os.path.join(helpers.ROOT, "include", "sa_yaapt.h")


def track_opts(i):
    """tda_frame_length / nccf_thresh1 of golden case i (bin/pipeline.py passes 25 ms / 0.25; the defaults are 35 ms / 0.3)."""
    t = np.load(GOLDEN)[f"c{i}_track_opts"]
    return dict(tda_frame_length=float(t[0]), nccf_thresh1=float(t[1]))


def cases():
    z = np.load(GOLDEN)
    for i in range(int(z["n_cases"])):
        fl, fs = [float(v) for v in z[f"c{i}_opts"]]
        wav = conditioning.waveform(int(z[f"c{i}_seed"]), float(z[f"c{i}_seconds"]))
        yield i, wav, dict(frame_length=fl, frame_space=fs), {k: z[f"c{i}_{k}"] for k in ("filtered", "filtered_nl", "energy", "vuv",
                                                                                          "mean_energy", "nframes", "shc", "cand_pitch", "cand_merit",
                                                                                          "spec_pitch", "pitch_std", "time_pitch1", "time_merit1",
                                                                                          "time_pitch2", "time_merit2", "final_pitch")
                                                              if f"c{i}_{k}" in z}


def compare(got, ref, what, thr=0.75):
    """got / ref: dicts with filtered, filtered_nl, energy, vuv, mean_energy for ONE utterance."""
    n, f = len(ref["filtered"]), len(ref["energy"])
    for k in ("filtered", "filtered_nl"):
        peak = np.abs(ref[k]).max()
        err = np.abs(np.asarray(got[k][:n], dtype=np.float64) - ref[k]).max()
        assert err <= 5e-4 * peak, f"{what}: {k} differs by {err:.3e} (peak {peak:.3e})"
    e_ref = np.asarray(ref["energy"], dtype=np.float64)
    e_got = np.asarray(got["energy"][:f], dtype=np.float64)
    assert np.abs(e_got - e_ref).max() <= 1e-4 * max(1.0, e_ref.max()), f"{what}: energy differs by {np.abs(e_got - e_ref).max():.3e}"
    decided = np.abs(e_ref - thr) > 1e-3
    assert np.array_equal(np.asarray(got["vuv"][:f]).astype(bool)[decided], np.asarray(ref["vuv"]).astype(bool)[decided]), f"{what}: vuv"
    assert abs(float(got["mean_energy"]) - float(ref["mean_energy"])) <= 1e-3 * float(ref["mean_energy"]), f"{what}: mean energy"


def test_oracle_matches_the_reference_outputs():
    n = 0
    for i, wav, opts, ref in cases():
        o = onp.nlfer(wav, onp.params(**opts))
        assert o["nframes"] == int(ref["nframes"])
        compare(o, ref, f"oracle case {i}")
        n += 1
    assert n >= 4


def compare_shc(got, ref, vuv, what, tol=2e-3):
    """SHC is a sum of products of four magnitudes: four times the relative noise of the filtered signal.  Per voiced frame
    within `tol` of that frame's peak (the float64 oracle is 5e-7 from the reference on the reference's own filtered signal)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: {got.shape} vs {ref.shape}"
    for f in range(ref.shape[0]):
        if not vuv[f]:
            assert not got[f].any(), f"{what}: unvoiced frame {f} must be zero"
            continue
        peak = ref[f].max()
        assert peak > 0 and np.abs(got[f] - ref[f]).max() <= tol * peak, f"{what}: frame {f} differs by {np.abs(got[f] - ref[f]).max() / peak:.2e} of its peak"


def test_oracle_shc_matches_the_vectors_the_reference_hands_to_peaks():
    for i, wav, opts, ref in cases():
        p = onp.params(**opts)
        compare_shc(onp.shc(ref["filtered_nl"], ref["vuv"], p), ref["shc"], ref["vuv"], f"oracle SHC case {i} (reference's filtered)", 1e-5)
        o = onp.nlfer(wav, p)
        compare_shc(onp.shc(o["filtered_nl"], ref["vuv"], p), ref["shc"], ref["vuv"], f"oracle SHC case {i} (own filtered)")


def test_oracle_peaks_matches_what_the_reference_returns():
    """`peaks` restated (oracle) on the reference's own SHC vectors: the same candidates, merits to float32 rounding."""
    total = 0
    for i, wav, opts, ref in cases():
        cp, cm = onp.spec_candidates(ref["shc"], ref["vuv"], onp.params(**opts))
        np.testing.assert_array_equal(cp, ref["cand_pitch"])
        np.testing.assert_allclose(cm, ref["cand_merit"], rtol=0, atol=2e-7)
        total += int(ref["vuv"].sum())
    assert total >= 150


def test_oracle_spec_track_finish_matches_the_reference():
    """The per-utterance part of spec_track restated (median smoothing, dynamic5 / path1, re-sampling) on the reference's own
    candidates against what the reference's spec_track returned."""
    n = 0
    for i, wav, opts, ref in cases():
        if not len(ref["spec_pitch"]):
            continue                                            # fewer than four frames: the reference raises (yaapt.py:311)
        sp, sd = onp.spec_track_finish(ref["cand_pitch"], ref["cand_merit"], onp.params(**opts))
        assert np.abs(sp - ref["spec_pitch"]).max() <= 1e-3 and abs(float(sd) - float(ref["pitch_std"])) <= 1e-4 * float(ref["pitch_std"])
        n += 1
    assert n >= 3


def test_oracle_trackers_match_the_reference():
    """time_track (NCCF + cmp_rate, with the in-place mean removal of crs_corr), refine and dynamic restated: on the
    reference's own intermediate results the oracle returns the reference's tracks (pitches identical but for a borderline
    frame, merits to float32 rounding) and EXACTLY the final pitch yaapt() returns."""
    n = 0
    for i, wav, opts, ref in cases():
        if "final_pitch" not in ref:
            continue
        tp = onp.track_params(**opts, **track_opts(i))
        for tag, sig in (("1", ref["filtered"]), ("2", ref["filtered_nl"])):
            op, om = onp.time_track(sig, ref["spec_pitch"], ref["pitch_std"], tp)
            same = (op == ref["time_pitch" + tag]).all(0)
            assert same.mean() >= 0.98 and np.abs(om - ref["time_merit" + tag])[:, same].max() <= 1e-5
        rp, rm = onp.refine(ref["time_pitch1"], ref["time_merit1"], ref["time_pitch2"], ref["time_merit2"], ref["spec_pitch"],
                            ref["energy"], ref["vuv"], tp)
        np.testing.assert_array_equal(onp.dynamic(rp, rm, ref["energy"], tp), ref["final_pitch"])
        n += 1
    assert n >= 3


def check_candidates(cp, cm, shc_rows, vuv, opts, what):
    """The GPU's candidates against `peaks` (oracle) run on the GPU's OWN SHC rows: same input, so the same decisions --
    identical pitches, merits to rounding (the mean over the lag range is summed in another order)."""
    ocp, ocm = onp.spec_candidates(shc_rows, vuv, onp.params(**opts))
    same = (cp == ocp).all(0)
    assert same.mean() >= 0.99, f"{what}: candidate pitches differ in {int((~same).sum())} of {len(same)} frames"
    assert np.abs(cm - ocm)[:, same].max() <= 1e-5, f"{what}: merits differ by {np.abs(cm - ocm)[:, same].max():.2e}"


def test_oracle_loop_and_compiled_recursions_agree():
    """The documented sample loop of the oracle and the compiled recursion it normally uses are the same filter."""
    wav = conditioning.waveform(5, 0.2).astype(np.float64)
    b, a = onp.normalized_coeffs("low", 16000.0, 50.0)
    saved = onp._scipy_lfilter
    try:
        onp._scipy_lfilter = None
        slow = onp.lfilter_clamped(wav, b, a)
    finally:
        onp._scipy_lfilter = saved
    assert np.abs(slow - onp.lfilter_clamped(wav, b, a)).max() < 1e-15


def test_library_exports_every_symbol_of_the_yaapt_header():
    lib = _lib.load()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(sa_yaapt_[a-z_0-9]+)\s*\(", src)))
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sa_yaapt.h but not exported"
    assert sorted(_lib.YAAPT_SYMBOLS) == declared


def test_frame_geometry_matches_the_reference():
    """No GPU needed: pitch.nframes / signal.size as the reference computes them (yaapt.py:164-166, 875-876)."""
    from satools_b200 import yaapt_frontend as yf
    lib = _lib.load()
    for i, wav, opts, ref in cases():
        assert yf.num_frames(len(wav), **opts) == int(ref["nframes"])
        assert lib.sa_yaapt_padded_length(yf.params(**opts), len(wav)) == len(ref["filtered"])
    assert lib.sa_yaapt_num_frames(None, 10) < 0 and b"bad" in lib.sa_yaapt_last_error()
    with pytest.raises(KeyError):
        yf.params(frame_len=35.0)                     # not an option of _yaapt
    p = yf.params()
    assert (p.sr, p.frame_length, p.frame_space, p.fft_length, p.bp_low, p.bp_high) == (16000.0, 35.0, 10.0, 8192.0, 50.0, 1500.0)


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _item(out, b):
    return dict(filtered=out.filtered[b].cpu().numpy(), filtered_nl=out.filtered_nl[b].cpu().numpy(), energy=out.energy[b].cpu().numpy(),
                vuv=out.vuv[b].cpu().numpy(), mean_energy=float(out.mean_energy[b]))


@pytest.mark.gpu
def test_cuda_front_end_matches_the_reference_outputs():
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    for i, wav, opts, ref in cases():
        out = yf.nlfer(torch.from_numpy(wav).to("cuda:0"), **opts)
        assert out.nframes == [int(ref["nframes"])] and out.padded_lengths == [len(ref["filtered"])]
        compare(_item(out, 0), ref, f"cuda case {i}")
        shc, cp, cm = yf.spec_shc(out, candidates=True, **opts)
        assert tuple(shc.shape) == (1,) + ref["shc"].shape and tuple(cp.shape) == (1,) + ref["cand_pitch"].shape
        check_candidates(cp[0].cpu().numpy(), cm[0].cpu().numpy(), shc[0].cpu().numpy(), out.vuv[0].cpu().numpy(), opts, f"cuda peaks case {i}")
        if np.array_equal(out.vuv[0].cpu().numpy(), ref["vuv"]):        # against the reference's own candidates: SHC noise may move a
            same = (cp[0].cpu().numpy() == ref["cand_pitch"]).all(0)     # borderline peak, so a fraction, and merits where they agree
            assert same.mean() >= 0.95, f"case {i}: {int((~same).sum())} of {len(same)} frames differ from the reference's candidates"
            assert np.abs(cm[0].cpu().numpy() - ref["cand_merit"])[:, same].max() <= 5e-3
        if np.array_equal(out.vuv[0].cpu().numpy(), ref["vuv"]):
            compare_shc(shc[0].cpu().numpy(), ref["shc"], ref["vuv"], f"cuda SHC case {i}")
        else:                                               # a frame at the threshold flipped: compare the common voiced frames
            both = out.vuv[0].cpu().numpy() & ref["vuv"].astype(bool)
            compare_shc(shc[0].cpu().numpy() * both[:, None], ref["shc"] * both[:, None], both, f"cuda SHC case {i}")


@pytest.mark.gpu
def test_cuda_front_end_ragged_batch_against_the_oracle_and_alone():
    """A padded batch with true lengths (chunks that end inside an item, items shorter than one chunk, one of 20 s: 300+
    chunks with a warm-up start) against the float64 oracle; every item also bit-identical to its solo run."""
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    opts = dict(frame_length=35.0, frame_space=20.0)
    secs = [20.0, 3.1, 0.064, 7.77, 0.5, 12.3]
    wavs = [conditioning.waveform(100 + i, s) for i, s in enumerate(secs)]
    n = max(len(w) for w in wavs)
    x = np.zeros((len(wavs), n), dtype=np.float32)
    rng = np.random.default_rng(0)
    for b, w in enumerate(wavs):
        x[b, :len(w)] = w
        x[b, len(w):] = rng.standard_normal(n - len(w)) * 0.1      # garbage beyond the true length must not matter
    out = yf.nlfer(torch.from_numpy(x).to("cuda:0"), lengths=[len(w) for w in wavs], **opts)
    shc, cp, cm = yf.spec_shc(out, lengths=[len(w) for w in wavs], candidates=True, **opts)
    for b, w in enumerate(wavs):
        o = onp.nlfer(w, onp.params(**opts))
        assert out.nframes[b] == o["nframes"]
        compare(_item(out, b), {k: o[k] for k in ("filtered", "filtered_nl", "energy", "vuv", "mean_energy")}, f"batch item {b}")
        f, npad = out.nframes[b], out.padded_lengths[b]
        assert float(out.energy[b, f:].abs().sum()) == 0.0 and float(out.filtered[b, npad:].abs().sum()) == 0.0
        vb = out.vuv[b, :f].cpu().numpy()
        compare_shc(shc[b, :f].cpu().numpy(), onp.shc(o["filtered_nl"], vb, onp.params(**opts)), vb, f"batch item {b} SHC")
        assert float(shc[b, f:].abs().sum()) == 0.0
        check_candidates(cp[b, :, :f].cpu().numpy(), cm[b, :, :f].cpu().numpy(), shc[b, :f].cpu().numpy(), vb, opts, f"batch item {b} peaks")
        assert float(cp[b, :, f:].abs().sum()) == 0.0 and bool((cm[b, :, f:] == 1).all())
        solo = yf.nlfer(torch.from_numpy(w).to("cuda:0"), **opts)
        s_shc, s_cp, s_cm = yf.spec_shc(solo, candidates=True, **opts)
        assert torch.equal(s_shc[0], shc[b, :f]) and torch.equal(s_cp[0], cp[b, :, :f]) and torch.equal(s_cm[0], cm[b, :, :f])
        assert torch.equal(solo.energy[0], out.energy[b, :f]) and torch.equal(solo.vuv[0], out.vuv[b, :f])
        assert torch.equal(solo.filtered[0], out.filtered[b, :npad]) and torch.equal(solo.filtered_nl[0], out.filtered_nl[b, :npad])


@pytest.mark.gpu
def test_cuda_front_end_properties_at_full_size():
    """BASELINE configs[1] size (64 utterances of 10-15 s).  Scaling the input by a power of two scales `filtered` by the same
    factor exactly (every rounding commutes with it, and nothing reaches the clamp), `filtered_nl` by its square, and leaves
    the normalised energy and the voiced flags bit-identical; items 0 / 31 / 63 against the oracle."""
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    opts = dict(frame_length=35.0, frame_space=20.0)
    rng = np.random.default_rng(3)
    lens = [int(v) for v in rng.integers(160000, 240001, size=64)]
    x = np.zeros((64, 240000), dtype=np.float32)
    for b, n in enumerate(lens):
        x[b, :n] = conditioning.waveform(200 + b, n / 16000.0)[:n]
    xd = torch.from_numpy(x).to("cuda:0")
    a = yf.nlfer(xd, lengths=lens, **opts)
    h = yf.nlfer(xd * 0.5, lengths=lens, **opts)
    assert torch.equal(h.filtered, a.filtered * 0.5) and torch.equal(h.filtered_nl, a.filtered_nl * 0.25)
    assert torch.equal(h.energy, a.energy) and torch.equal(h.vuv, a.vuv)
    (sa, cpa, cma), sh = yf.spec_shc(a, lengths=lens, candidates=True, **opts), yf.spec_shc(h, lengths=lens, **opts)
    assert torch.equal(sh, sa * (0.25 ** 4))                 # four magnitudes of the squared signal per product
    voiced = a.vuv.unsqueeze(1).expand_as(cpa)
    assert bool(((cpa[voiced] == 0) | ((cpa[voiced] >= 30.0) & (cpa[voiced] <= 860.0))).all())     # n * delta, halved or doubled
    assert bool((cma[voiced] >= 0).all()) and bool((cma[voiced] <= 1.0).all())
    assert 0.2 < float(a.vuv.float().mean()) < 0.9
    for b in (0, 31, 63):
        o = onp.nlfer(x[b, :lens[b]], onp.params(**opts))
        compare(_item(a, b), {k: o[k] for k in ("filtered", "filtered_nl", "energy", "vuv", "mean_energy")}, f"full-size item {b}")
        f = a.nframes[b]
        vb = a.vuv[b, :f].cpu().numpy()
        compare_shc(sa[b, :f].cpu().numpy(), onp.shc(o["filtered_nl"], vb, onp.params(**opts)), vb, f"full-size item {b} SHC")


@pytest.mark.gpu
def test_cuda_spec_track_against_the_reference():
    """sa_yaapt_spec_track on the reference's own candidate matrices reproduces the reference's spec_pitch / pitch_std (same
    decisions, float32 rounding); the whole GPU chain (waveform -> spec_pitch) stays close to it: its candidates differ from
    the reference's in a few borderline frames, which the dynamic programming and the median filters mostly absorb."""
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    for i, wav, opts, ref in cases():
        front = yf.nlfer(torch.from_numpy(wav).to("cuda:0"), **opts)
        if not len(ref["spec_pitch"]):
            with pytest.raises(IndexError):
                yf.spec_track(front, **opts)
            continue
        cp = torch.from_numpy(ref["cand_pitch"][None]).to("cuda:0")
        cm = torch.from_numpy(ref["cand_merit"][None]).to("cuda:0")
        sp, sd = yf.spec_track_from_candidates(cp, cm, len(wav), **opts)
        assert np.abs(sp[0].cpu().numpy() - ref["spec_pitch"]).max() <= 1e-3, f"case {i}"
        assert abs(float(sd[0]) - float(ref["pitch_std"])) <= 1e-4 * float(ref["pitch_std"])
        sp2, sd2 = yf.spec_track(front, **opts)
        dev = np.abs(sp2[0].cpu().numpy() - ref["spec_pitch"])
        print(f"case {i}: whole chain vs reference spec_pitch: median |d| {np.median(dev):.3f} Hz, 90 % {np.quantile(dev, 0.9):.3f} Hz, "
              f"max {dev.max():.2f} Hz; pitch_std {float(sd2[0]):.3f} vs {float(ref['pitch_std']):.3f}")
        assert np.median(dev) <= 0.5 and np.quantile(dev, 0.9) <= 5.0
        assert abs(float(sd2[0]) - float(ref["pitch_std"])) <= 0.1 * float(ref["pitch_std"])


@pytest.mark.gpu
def test_cuda_spec_track_batch_against_the_oracle():
    """Ragged batch: every item's spec_pitch / pitch_std against the oracle's restatement run on the GPU's own candidates
    (same input -> same decisions), zero beyond the item's frames."""
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    opts = dict(frame_length=35.0, frame_space=20.0)
    secs = [6.0, 2.2, 9.4, 0.7]
    wavs = [conditioning.waveform(300 + i, s) for i, s in enumerate(secs)]
    n = max(len(w) for w in wavs)
    x = np.zeros((len(wavs), n), dtype=np.float32)
    for b, w in enumerate(wavs):
        x[b, :len(w)] = w
    lens = [len(w) for w in wavs]
    front = yf.nlfer(torch.from_numpy(x).to("cuda:0"), lengths=lens, **opts)
    _, cp, cm = yf.spec_shc(front, lengths=lens, candidates=True, **opts)
    sp, sd = yf.spec_track(front, lengths=lens, **opts)
    for b in range(len(wavs)):
        f = front.nframes[b]
        osp, osd = onp.spec_track_finish(cp[b, :, :f].cpu().numpy(), cm[b, :, :f].cpu().numpy(), onp.params(**opts))
        assert np.abs(sp[b, :f].cpu().numpy() - osp).max() <= 1e-3, f"item {b}"
        assert abs(float(sd[b]) - float(osd)) <= 1e-4 * float(osd)
        assert float(sp[b, f:].abs().sum()) == 0.0


@pytest.mark.gpu
def test_cuda_yaapt_final_pitch_against_the_reference():
    """The whole extractor on the GPU, waveform -> final pitch per frame, against what the reference's yaapt() returned for
    the same waveform and options: voiced / unvoiced decisions and pitch values."""
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    seen = 0
    for i, wav, opts, ref in cases():
        if "final_pitch" not in ref:
            continue
        got = yf.yaapt(torch.from_numpy(wav).to("cuda:0"), **opts, **track_opts(i))[0].cpu().numpy()
        want = ref["final_pitch"]
        assert got.shape == want.shape
        vuv_same = ((got > 0) == (want > 0)).mean()
        close = (np.abs(got - want) <= 1e-2).mean()
        print(f"case {i}: voiced/unvoiced decisions equal in {100 * vuv_same:.1f} % of {len(want)} frames, pitch within 0.01 Hz in "
              f"{100 * close:.1f} %, max |d| {np.abs(got - want).max():.2f} Hz")
        assert vuv_same >= 0.97 and close >= 0.95
        seen += 1
    assert seen >= 3


@pytest.mark.gpu
def test_cuda_yaapt_batch_equals_solo_and_oracle_chain():
    """A ragged batch through yaapt(): every item bit-identical to its solo run, and equal to the oracle's trackers run on the
    GPU's own front-end / spec_track results (same inputs -> the same decisions in all but borderline frames)."""
    _need_gpu()
    from satools_b200 import yaapt_frontend as yf
    opts = dict(frame_length=35.0, frame_space=20.0, nccf_thresh1=0.25, tda_frame_length=25.0)
    fopts = dict(frame_length=35.0, frame_space=20.0)
    secs = [4.0, 1.3, 6.5]
    wavs = [conditioning.waveform(400 + i, s) for i, s in enumerate(secs)]
    n = max(len(w) for w in wavs)
    x = np.zeros((len(wavs), n), dtype=np.float32)
    for b, w in enumerate(wavs):
        x[b, :len(w)] = w
    lens = [len(w) for w in wavs]
    xd = torch.from_numpy(x).to("cuda:0")
    final = yf.yaapt(xd, lengths=lens, **opts)
    front = yf.nlfer(xd, lengths=lens, **fopts)
    spec, std = yf.spec_track(front, lengths=lens, **fopts)
    tp = onp.track_params(**opts)
    for b, w in enumerate(wavs):
        f, npad = front.nframes[b], front.padded_lengths[b]
        solo = yf.yaapt(torch.from_numpy(w).to("cuda:0"), **opts)
        assert torch.equal(solo[0], final[b, :f]) and float(final[b, f:].abs().sum()) == 0.0
        sp, sd = spec[b, :f].cpu().numpy(), float(std[b])
        t1 = onp.time_track(front.filtered[b, :npad].cpu().numpy(), sp, sd, tp)
        t2 = onp.time_track(front.filtered_nl[b, :npad].cpu().numpy(), sp, sd, tp)
        rp, rm = onp.refine(t1[0], t1[1], t2[0], t2[1], sp, front.energy[b, :f].cpu().numpy(), front.vuv[b, :f].cpu().numpy(), tp)
        want = onp.dynamic(rp, rm, front.energy[b, :f].cpu().numpy(), tp)
        got = final[b, :f].cpu().numpy()
        close = (np.abs(got - want) <= 1e-2).mean()
        assert close >= 0.97, f"item {b}: {100 * close:.1f} % of the frames agree with the oracle chain"


@pytest.mark.gpu
def test_drop_in_yaapt_call_signature():
    """satools_b200.install.yaapt(_in, kwargs) -- the reference's call (`hifigan.yaapt.yaapt(wav, self.f0_yaapt_opts)`): CPU
    tensor in, option dict with the reference's names (unknown names ignored), [B, n_frames] on the input's device out."""
    _need_gpu()
    import importlib
    from satools_b200 import yaapt_frontend as yf
    inst = importlib.import_module("satools_b200.install")      # the module (satools_b200.install is the function)
    wav = torch.from_numpy(np.stack([conditioning.waveform(7, 1.0), conditioning.waveform(8, 1.0)]))
    opts = {"frame_length": 35.0, "frame_space": 20.0, "nccf_thresh1": 0.25, "tda_frame_length": 25.0, "not_an_option": 1.0}
    f0 = inst.yaapt(wav, opts)
    assert f0.device == wav.device and tuple(f0.shape) == (2, yf.num_frames(16000, frame_length=35.0, frame_space=20.0))
    ref = yf.yaapt(wav.to("cuda:0"), frame_length=35.0, frame_space=20.0, nccf_thresh1=0.25, tda_frame_length=25.0)
    assert torch.equal(f0, ref.cpu()) and bool((f0 >= 0).all()) and float(f0.max()) > 60.0
