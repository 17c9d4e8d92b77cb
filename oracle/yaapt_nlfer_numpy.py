"""CPU restatement of the FRONT END of the reference's YAAPT F0 extractor (row N2 of SURVEY.md section 8f).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu legs and __graft_entry__.smoke(); the product path
(satools_b200.yaapt_frontend -> libsatools_hifigan.so) never imports it.

What it restates (/root/reference/satools/satools/hifigan/yaapt.py):
  * `_yaapt` lines 873-880: zero padding of the waveform by frame_length / 2 on both sides, the squared ("nonlinear") signal;
  * `SignalObj.filtered_version` lines 42-52: torchaudio `lowpass_biquad(x, fs, bp_low)` then `highpass_biquad(., fs, bp_high)`
    (Q = 0.707, each through `lfilter(clamp=True)`: FIR part b / a0, IIR recursion with a / a0, output clamped to [-1, 1]).
    NB the reference passes bp_low = 50 Hz to the LOW-pass and bp_high = 1500 Hz to the HIGH-pass: what is left is the small
    residue both filters leak.  A drop-in has to reproduce that, not fix it;
  * `nlfer` lines 148-176: frames of frame_size samples every frame_jump samples, Hann window `hann_window(n + 2)[1:-1]`,
    magnitude of the nfft-point DFT summed over bins [N_f0_min - 1, N_f0_max), N_f0_min = round(2 f0_min / fs * nfft),
    N_f0_max = round(f0_max / fs * nfft);
  * `PitchObj.set_energy` lines 124-127: energy / mean(energy), voiced = energy > nlfer_thresh1;
  * `spec_track` lines 184-231 up to the call of `peaks`: for every voiced frame the spectral harmonics correlation SHC of the
    squared signal: 2 frame_size samples x Kaiser(beta 0.5, periodic) window, mean removed, |rfft_nfft|, then
    SHC[k] = sum_c prod_{h=1..numharms+1} |X|[h k + c - half_window] over a window of `window_length` bins, k in [min_SHC, max_SHC];
  * `peaks` lines 383-497: the (up to shc_maxpeaks) pitch candidates and merits of one SHC vector.

Pinned by tests/golden/yaapt_nlfer.npz (outputs of the reference itself, oracle/make_golden_yaapt.py).  The filters are
evaluated in float64 here: the reference runs them in float32, and its own rounding noise is what sets the tolerance of the
comparison (see tests/test_yaapt_frontend.py).
"""
import math

import numpy as np

try:
    from scipy.signal import lfilter as _scipy_lfilter
except Exception:  # pragma: no cover
    _scipy_lfilter = None

DEFAULTS = dict(sr=16000.0, frame_length=35.0, frame_space=10.0, f0_min=60.0, f0_max=400.0, fft_length=8192.0, bp_low=50.0,
                bp_high=1500.0, nlfer_thresh1=0.75, shc_numharms=3.0, shc_window=40.0, shc_pwidth=50.0, shc_maxpeaks=4.0,
                shc_thresh1=5.0, shc_thresh2=1.25, f0_double=150.0, f0_half=150.0, merit_extra=0.4)


def params(**kw):
    p = dict(DEFAULTS)
    p.update(kw)
    return p


def biquad_coeffs(kind, fs, cutoff, q=0.707):
    """torchaudio lowpass_biquad / highpass_biquad: (b0, b1, b2, a0, a1, a2), evaluated in float32 as torchaudio does for a
    float32 waveform (`torch.as_tensor(cutoff, dtype=waveform.dtype)`)."""
    f32 = np.float32
    w0 = f32(2 * math.pi) * f32(cutoff) / f32(int(fs))
    alpha = np.sin(w0, dtype=f32) / f32(2) / f32(q)
    c = np.cos(w0, dtype=f32)
    if kind == "low":
        b0 = (f32(1) - c) / f32(2)
        b1 = f32(1) - c
    else:
        b0 = (f32(1) + c) / f32(2)
        b1 = f32(-1) - c
    return tuple(float(v) for v in (b0, b1, b0, f32(1) + alpha, f32(-2) * c, f32(1) - alpha))


def normalized_coeffs(kind, fs, cutoff):
    """(b / a0, a / a0) as float32 arrays: what `_lfilter` hands to its FIR and IIR parts."""
    b0, b1, b2, a0, a1, a2 = biquad_coeffs(kind, fs, cutoff)
    f32 = np.float32
    b = np.array([b0, b1, b2], dtype=f32) / f32(a0)
    a = np.array([a0, a1, a2], dtype=f32) / f32(a0)
    return b, a


def lfilter_clamped(x, b, a, dtype=np.float64):
    """y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2], zero initial state, clamped to [-1, 1]."""
    x = np.asarray(x, dtype=dtype)
    b = b.astype(dtype)
    a = a.astype(dtype)
    if _scipy_lfilter is not None and dtype == np.float64:       # same recursion, compiled
        return np.clip(_scipy_lfilter(b, a, x), -1.0, 1.0)
    xp = np.concatenate([np.zeros(2, dtype=dtype), x])
    fir = b[0] * xp[2:] + b[1] * xp[1:-1] + b[2] * xp[:-2]
    y = np.zeros(len(x) + 2, dtype=dtype)
    a1, a2 = a[1], a[2]
    for n in range(len(x)):
        y[n + 2] = fir[n] - a2 * y[n] - a1 * y[n + 1]
    return np.clip(y[2:], -1.0, 1.0)


def filtered(x, p, dtype=np.float64):
    bl, al = normalized_coeffs("low", p["sr"], p["bp_low"])
    bh, ah = normalized_coeffs("high", p["sr"], p["bp_high"])
    return lfilter_clamped(lfilter_clamped(x, bl, al, dtype), bh, ah, dtype)


def geometry(n_samples, p):
    """(pad, frame_size, frame_jump, n_frames, bin_lo, bin_hi) for a waveform of n_samples (before padding)."""
    pad = int(p["frame_length"] / 1000 * int(p["sr"])) // 2
    frame_size = int(math.floor(p["frame_length"] * p["sr"] / 1000))
    frame_jump = int(math.floor(p["frame_space"] * p["sr"] / 1000))
    size = n_samples + 2 * pad
    half = frame_size // 2
    n_frames = len(range(half, size - half, frame_jump))
    nfft = int(p["fft_length"])
    # torch.round: half to even, on float32 values
    lo = int(np.round(np.float32(p["f0_min"] * 2 / float(p["sr"])) * np.float32(nfft)))
    hi = int(np.round(np.float32(p["f0_max"] / float(p["sr"])) * np.float32(nfft)))
    return pad, frame_size, frame_jump, n_frames, lo - 1, hi


def nlfer(wav, p=None):
    """wav: [n] float -> dict(energy [n_frames] normalised, vuv [n_frames] bool, mean_energy, frame_energy, filtered)."""
    p = p or params()
    wav = np.asarray(wav, dtype=np.float64).reshape(-1)
    pad, frame_size, frame_jump, n_frames, lo, hi = geometry(len(wav), p)
    x = np.concatenate([np.zeros(pad), wav, np.zeros(pad)])
    filt = filtered(x, p)
    n = np.arange(frame_size + 2, dtype=np.float64)
    window = (0.5 - 0.5 * np.cos(2 * np.pi * n / (frame_size + 2)))[1:-1]        # periodic hann_window(frame_size + 2)[1:-1]
    nfft = int(p["fft_length"])
    k = np.arange(lo, hi, dtype=np.float64)
    ang = -2j * np.pi * np.outer(np.arange(frame_size, dtype=np.float64), k) / nfft
    basis = np.exp(ang)                                                           # [frame_size, bins]
    frames = np.stack([filt[i * frame_jump:i * frame_jump + frame_size] for i in range(n_frames)]) * window
    frame_energy = np.abs(frames @ basis).sum(1)
    mean_energy = frame_energy.mean() if n_frames else 0.0
    energy = frame_energy / mean_energy
    return dict(energy=energy, vuv=energy > p["nlfer_thresh1"], mean_energy=mean_energy, frame_energy=frame_energy, filtered=filt,
                filtered_nl=filtered(x * x, p), nframes=n_frames)


def shc_geometry(p):
    """(nframe_size, window_length, half_window_length, min_SHC, max_SHC, n_harm) of spec_track (yaapt.py:189-201)."""
    frame_size = int(math.floor(p["frame_length"] * p["sr"] / 1000))
    delta = p["sr"] / int(p["fft_length"])
    window_length = int(math.floor(p["shc_window"] / delta))
    half = int(math.floor(float(window_length) / 2))
    if not (window_length % 2):
        window_length += 1
    max_shc = int(math.floor((p["f0_max"] + p["shc_pwidth"] * 2) / delta))
    min_shc = int(math.ceil(p["f0_min"] / delta))
    return 2 * frame_size, window_length, half, min_shc, max_shc, int(p["shc_numharms"]) + 1


def shc(filtered_nl, vuv, p=None):
    """filtered_nl: SignalObj.filtered of the squared signal [size]; vuv [n_frames] -> [n_frames, max_SHC] (zero rows for
    unvoiced frames; entries outside [min_SHC - 1, max_SHC) are zero, as in the reference's SHC buffer)."""
    p = p or params()
    nframe, wl, half, min_shc, max_shc, n_harm = shc_geometry(p)
    frame_jump = int(math.floor(p["frame_space"] * p["sr"] / 1000))
    nfft = int(p["fft_length"])
    x = np.asarray(filtered_nl, dtype=np.float64)
    n_frames = len(vuv)
    need = nframe + (n_frames - 1) * frame_jump
    if need > len(x):
        x = np.concatenate([x, np.zeros(need - len(x))])
    window = np.kaiser(nframe + 1, 0.5)[:-1]                     # torch.kaiser_window(nframe, periodic=True, beta=0.5)
    out = np.zeros((n_frames, max_shc))
    rows = max_shc - min_shc + 1
    for f in np.nonzero(np.asarray(vuv))[0]:
        sl = x[f * frame_jump:f * frame_jump + nframe] * window
        sl = sl - sl.mean()
        mag = np.concatenate([np.zeros(half), np.abs(np.fft.rfft(sl, nfft))])
        acc = np.ones((rows, wl))
        for h in range(1, n_harm + 1):
            idx = min_shc * h + np.arange(rows)[:, None] * h + np.arange(wl)[None, :]
            acc = acc * mag[idx]
        out[f, min_shc - 1:max_shc] = acc.sum(1)
    return out


def peaks(data, p=None):
    """`peaks(data, delta, maxpeaks, parameters)` of the reference (yaapt.py:383-497) on one SHC vector: (pitch [maxpeaks],
    merit [maxpeaks]) float32.  float32 arithmetic where the reference has float32 tensors."""
    p = p or params()
    f32 = np.float32
    data = np.asarray(data, dtype=f32)
    delta = p["sr"] / int(p["fft_length"])
    maxpeaks = int(p["shc_maxpeaks"])
    t1, t2 = p["shc_thresh1"], p["shc_thresh2"]
    width = int(math.floor(p["shc_pwidth"] / delta))
    if not (float(width) % 2):
        width += 1
    center = int(math.ceil(width / 2))
    min_lag = int(math.floor(p["f0_min"] / delta - center))
    max_lag = int(math.floor(p["f0_max"] / delta + center))
    min_lag = max(min_lag, 1)
    max_lag = min(max_lag, len(data) - width)
    unvoiced = (np.zeros(maxpeaks, dtype=f32), np.ones(maxpeaks, dtype=f32))
    max_data = data[min_lag:max_lag + 1].max()
    if max_data > 1e-14:
        data = data / max_data
    avg = data[min_lag:max_lag + 1].mean(dtype=f32)
    if avg > 1 / t1:
        return unvoiced
    lo, hi = min_lag + center + 1, max_lag - center + 1
    mid = data[lo:hi]
    is_peak = (mid > data[lo - 1:hi - 1]) & (mid > data[lo + 1:hi + 1]) & (mid > t2 * avg)
    pitch, merit = [], []
    for n in (np.nonzero(is_peak)[0] + lo):
        if int(np.argmax(data[n - center:n + center + 1])) == center:
            pitch.append(f32(float(n) * delta))
            merit.append(data[n])
    numpeaks = len(pitch)
    pitch = np.array(pitch + [0.0] * max(0, maxpeaks - numpeaks), dtype=f32)
    merit = np.array(merit + [0.0] * max(0, maxpeaks - numpeaks), dtype=f32)
    if merit.max() / avg < t1:
        return unvoiced
    idx = np.argsort(-merit, kind="stable")
    merit, pitch = merit[idx], pitch[idx]
    numpeaks = min(numpeaks, maxpeaks)
    pitch = np.concatenate([pitch[:numpeaks], np.zeros(maxpeaks - numpeaks, dtype=f32)])
    merit = np.concatenate([merit[:numpeaks], np.zeros(maxpeaks - numpeaks, dtype=f32)])
    if numpeaks == 0:
        return unvoiced
    if pitch[0] > p["f0_double"]:
        numpeaks = min(numpeaks + 1, maxpeaks)
        pitch[numpeaks - 1] = pitch[0] / f32(2.0)
        merit[numpeaks - 1] = p["merit_extra"]
    if pitch[0] < p["f0_half"]:
        numpeaks = min(numpeaks + 1, maxpeaks)
        pitch[numpeaks - 1] = pitch[0] * f32(2.0)
        merit[numpeaks - 1] = p["merit_extra"]
    if numpeaks < maxpeaks:
        pitch[numpeaks:] = pitch[0]
        merit[numpeaks:] = merit[0]
    return pitch, merit


def spec_candidates(shc_rows, vuv, p=None):
    """cand_pitch / cand_merit [maxpeaks, n_frames] as spec_track fills them (yaapt.py:204-205, 231): zeros / ones for unvoiced frames."""
    p = p or params()
    m = int(p["shc_maxpeaks"])
    cp, cm = np.zeros((m, len(vuv)), dtype=np.float32), np.ones((m, len(vuv)), dtype=np.float32)
    for f in np.nonzero(np.asarray(vuv))[0]:
        cp[:, f], cm[:, f] = peaks(shc_rows[f], p)
    return cp, cm


# ---- the rest of spec_track (yaapt.py:233-316): from the per-frame candidates to the spectral pitch track -------------------
def _medfilt(x, k):
    """`medfilt` (yaapt.py:54-70): zero padding, median of k (odd) samples."""
    pad = k // 2
    xp = np.concatenate([np.zeros(pad, dtype=x.dtype), x, np.zeros(pad, dtype=x.dtype)])
    return np.array([np.sort(xp[i:i + k])[(k - 1) // 2] for i in range(len(x))], dtype=x.dtype)


def _path1(local, trans):
    """`path1` (yaapt.py:530-569): lowest-cost path; ties go to the LAST minimum (the flip / argmin idiom)."""
    n_lin, n_col = local.shape
    pred = np.zeros((n_lin, n_col), dtype=np.int64)
    p_small = np.zeros(n_col, dtype=np.int64)
    pcost = local[:, 0].copy()
    for i in range(1, n_col):
        aux = pcost[None, :] + trans[:, :, i]                       # aux[a, b] = PCOST[b] + trans[a, b, i]
        k = n_lin - np.argmin(aux[:, ::-1], axis=1) - 1
        pred[:, i] = k
        ccost = pcost[k] + trans[k, np.arange(n_lin), i] + local[:, i]
        pcost = ccost.astype(np.float32)
        p_small[i] = n_lin - np.argmin(ccost[::-1]) - 1
    path = np.ones(n_col, dtype=np.int64)
    path[-1] = p_small[-1]
    for i in range(n_col - 2, -1, -1):
        path[i] = pred[path[i + 1], i + 1]
    return path


def _dynamic5(pitch, merit, k1, f0_min):
    f32 = np.float32
    n_cand, n_frames = pitch.shape
    local = (f32(1) - merit).astype(f32)
    trans = np.zeros((n_cand, n_cand, n_frames), dtype=f32)
    d = np.abs(pitch[None, :, 1:] - pitch[:, None, :-1]) / f32(f0_min)      # trans[a, b, t] = |pitch[b, t] - pitch[a, t - 1]| / f0_min
    trans[:, :, 1:] = f32(0.05) * d + d * d
    trans = (f32(k1) * trans).astype(f32)
    path = _path1(local, trans)
    return pitch[path, np.arange(n_frames)]


def _interp_linear(x, size):
    """torch.nn.functional.interpolate(mode='linear', align_corners=False) of a 1-D float32 signal."""
    f32 = np.float32
    n = len(x)
    scale = f32(n) / f32(size)
    out = np.empty(size, dtype=f32)
    for i in range(size):
        src = max(f32(0), scale * (f32(i) + f32(0.5)) - f32(0.5))
        i0 = min(int(src), n - 1)
        i1 = min(i0 + 1, n - 1)
        lam = f32(src - f32(i0))
        out[i] = (f32(1) - lam) * x[i0] + lam * x[i1]
    return out


def spec_track_finish(cand_pitch, cand_merit, p=None, median_value=7.0, dp5_k1=11.0, spec_pitch_min_std=0.05):
    """(spec_pitch [n_frames], pitch_std) from spec_track's candidate matrices, float32 like the reference's tensors."""
    p = p or params()
    f32 = np.float32
    cand_pitch = np.array(cand_pitch, dtype=f32)
    cand_merit = np.array(cand_merit, dtype=f32)
    spec_pitch = cand_pitch[0].copy()
    mask = cand_pitch[0] > 0
    vcp, vcm = cand_pitch[:, mask].copy(), cand_merit[:, mask].copy()
    num = vcp.shape[1]
    k = max(1, int(median_value) - 2)
    if num > 2:
        avg_v = vcp[0].mean(dtype=f32)
        std_v = vcp[0].std(ddof=1, dtype=f32)
        delta1 = np.abs(vcp - f32(0.8) * avg_v) * (f32(3) - vcm)
        index = delta1.argmin(0)
        cols = np.arange(num)
        vcp[index, cols] = _medfilt(vcp[index, cols], k)
        weight = f32(dp5_k1) * std_v / avg_v
        voiced = _medfilt(_dynamic5(vcp, vcm, weight, p["f0_min"]), k)
    elif num > 0:
        voiced = np.full(num, 150.0, dtype=f32)
    else:
        voiced = np.array([150.0], dtype=f32)
    pitch_avg = voiced.mean(dtype=f32)
    std = voiced.std(ddof=1, dtype=f32) if len(voiced) > 1 else f32(np.nan)
    pitch_std = np.maximum(std, pitch_avg * f32(spec_pitch_min_std))
    if num > 0:
        spec_pitch[mask] = voiced
    if spec_pitch[0] < pitch_avg / 2:
        spec_pitch[0] = pitch_avg
    if spec_pitch[-1] < pitch_avg / 2:
        spec_pitch[-1] = pitch_avg
    nz = spec_pitch[spec_pitch != 0]
    out = _interp_linear(nz, len(spec_pitch))
    out[0] = out[2]
    out[1] = out[3]
    return out, f32(pitch_std)


# ---- time_track / crs_corr / cmp_rate (yaapt.py:577-731) ------------------------------------------------------------
TRACK_DEFAULTS = dict(tda_frame_length=35.0, nccf_thresh1=0.3, nccf_thresh2=0.9, nccf_maxcands=3.0, nccf_pwidth=5.0, merit_boost=0.2,
                      nlfer_thresh2=0.1, merit_pivot=0.99, median_value=7.0, dp_w1=0.15, dp_w2=0.5, dp_w3=0.1, dp_w4=0.9)


def track_params(**kw):
    p = params(**{k: v for k, v in kw.items() if k in DEFAULTS})
    p.update(TRACK_DEFAULTS)
    p.update({k: v for k, v in kw.items() if k in TRACK_DEFAULTS})
    return p


def time_track(filtered_sig, spec_pitch, pitch_std, p):
    """`time_track` with `crs_corr` and `cmp_rate` inlined.  Quirks kept: `crs_corr` removes the frame mean IN PLACE from a
    view of the signal buffer, so later (overlapping) frames see samples already shifted (yaapt.py:588, 709-711); `cmp_rate`
    looks at the FIRST local maximum only (`nonzero()[0]`, yaapt.py:636) and so returns at most one candidate per frame."""
    f32 = np.float32
    fs = p["sr"]
    jump = int(math.floor(p["frame_space"] * fs / 1000))
    tda_len = int(p["tda_frame_length"] * fs / 1000)
    data = np.array(filtered_sig, dtype=f32)
    n_frames = int((len(data) - (tda_len - jump)) / jump)
    spec_pitch = np.asarray(spec_pitch, dtype=f32)
    if n_frames < len(spec_pitch):
        spec_pitch = spec_pitch[:n_frames]
    elif n_frames > len(spec_pitch):
        n_frames = len(spec_pitch)
    maxcands = int(p["nccf_maxcands"])
    pitch_std = f32(pitch_std)
    freq_thresh = f32(5.0) * pitch_std
    lo = np.maximum(spec_pitch - f32(2.0) * pitch_std, f32(p["f0_min"]))
    hi = np.minimum(spec_pitch + f32(2.0) * pitch_std, f32(p["f0_max"]))
    center = int(math.floor(p["nccf_pwidth"] / 2.0))
    t1, t2 = p["nccf_thresh1"], p["nccf_thresh2"]
    time_pitch = np.zeros((maxcands, n_frames), dtype=f32)
    time_merit = np.zeros((maxcands, n_frames), dtype=f32)
    for f in range(n_frames):
        qa, qb = f32(fs) / hi[f], f32(fs) / lo[f]
        if np.isnan(qa) or np.isnan(qb):
            continue
        lag_min = int(math.floor(qa)) - center
        lag_max = int(math.floor(qb)) + center
        x = data[f * jump:f * jump + tda_len]                       # a VIEW: the mean removal below stays in `data`
        n = tda_len - lag_max
        assert n > 0
        x -= x.mean(dtype=f32)
        xj = x[:n]
        pw = f32(np.dot(xj, xj))
        phi = np.zeros(tda_len, dtype=f32)
        with np.errstate(divide="ignore", invalid="ignore"):       # silent frames: 0 / 0 as in the reference (eps1 = 0, yaapt.py:579)
            for lag in range(lag_min, lag_max):
                xr = x[lag:lag + n]
                phi[lag] = f32(np.dot(xr, xj)) / np.sqrt(f32(np.dot(xr, xr)) * pw)
        seg = phi[lag_min + center:lag_max - center + 1]
        pk = (seg > phi[lag_min + center - 1:lag_max - center]) & (seg > phi[lag_min + center + 1:lag_max - center + 2]) & (seg > t1)
        nz = np.nonzero(pk)[0]
        if len(nz):
            n0 = int(nz[0]) + lag_min + center
            if phi.max() > t2 or int(np.argmax(phi[n0 - center:n0 + center + 1])) == center:
                time_pitch[0, f] = f32(fs / float(n0 + 1))
                time_merit[0, f] = phi[n0]
        if time_merit[:, f].max() > 1.0:
            time_merit[:, f] = time_merit[:, f] / time_merit[:, f].max()
    diff = np.abs(time_pitch - spec_pitch[None, :])
    match = (f32(1) - diff / freq_thresh) * (diff < freq_thresh)
    time_merit = (f32(1 + p["merit_boost"]) * time_merit) * match
    return time_pitch, time_merit.astype(f32)


# ---- refine / dynamic (yaapt.py:732-787, 321-372) and the assembly in _yaapt (lines 921-944) ---------------------------
def refine(tp1, tm1, tp2, tm2, spec_pitch, energy, vuv, p):
    f32 = np.float32
    n = len(spec_pitch)

    def padded(a):                                                  # _yaapt lines 921-931
        a = np.asarray(a, dtype=f32)
        return a if a.shape[1] >= n else np.concatenate([a, np.zeros((a.shape[0], n - a.shape[1]), dtype=f32)], 1)
    time_pitch = np.concatenate([padded(tp1), padded(tp2)], 0)
    time_merit = np.concatenate([padded(tm1), padded(tm2)], 0)
    c = time_pitch.shape[0]
    idx = np.argsort(-time_merit, axis=0, kind="stable")
    cols = np.arange(n)
    time_merit = time_merit[idx, cols]
    time_pitch = time_pitch[idx, cols]
    energy = np.asarray(energy, dtype=f32)
    best = _medfilt(time_pitch[0].copy(), int(p["median_value"])) * np.asarray(vuv, dtype=f32)
    t2 = f32(p["nlfer_thresh2"])
    idx1 = energy <= t2
    idx2 = (energy > t2) & (time_pitch[0] > 0)
    idx3 = (energy > t2) & (time_pitch[0] <= 0)
    merit_mat = np.zeros((c, n), dtype=bool)
    merit_mat[1:c - 1] = (time_pitch[1:c - 1] == 0) & idx2
    time_pitch[:, idx1] = 0
    time_merit[:, idx1] = f32(p["merit_pivot"])
    time_pitch[c - 1, idx2] = 0
    time_merit[c - 1, idx2] = f32(1) - time_merit[0, idx2]
    time_merit[merit_mat] = 0
    time_pitch[0, idx3] = np.asarray(spec_pitch, dtype=f32)[idx3]
    time_merit[0, idx3] = np.minimum(f32(1), energy[idx3] / f32(2))
    time_pitch[1:, idx3] = 0
    time_merit[1:, idx3] = f32(1) - time_merit[0, idx3]
    time_pitch[c - 2] = best
    nzf = best > 0
    time_merit[c - 2, nzf] = time_merit[0, nzf]
    time_merit[c - 2, ~nzf] = f32(1) - np.minimum(f32(1), energy[~nzf] / f32(2))
    time_pitch[c - 3] = spec_pitch
    time_merit[c - 3] = energy / f32(5)
    return time_pitch, time_merit


def dynamic(ref_pitch, ref_merit, energy, p):
    f32 = np.float32
    c, n = ref_pitch.shape
    best = ref_pitch[c - 2]
    mean_pitch = best[best > 0].mean(dtype=f32)
    local = (f32(1) - ref_merit).astype(f32)
    energy = np.asarray(energy, dtype=f32)
    trans = np.ones((c, c, n), dtype=f32)
    cur = ref_pitch[None, :, 1:]                                    # [a, b, t] = pitch[b, t]
    prev = ref_pitch[:, None, :-1]                                  # [a, b, t] = pitch[a, t - 1]
    cur, prev = np.broadcast_to(cur, (c, c, n - 1)), np.broadcast_to(prev, (c, c, n - 1))
    benefit = np.minimum(f32(1), np.abs(energy[:-1] - energy[1:]))
    t = trans[:, :, 1:]
    both = (cur > 0) & (prev > 0)
    one = ((cur == 0) & (prev > 0)) | ((cur > 0) & (prev == 0))
    none = (cur == 0) & (prev == 0)
    t[both] = (f32(p["dp_w1"]) * (np.abs(cur - prev) / mean_pitch))[both]
    t[one] = np.broadcast_to(f32(p["dp_w2"]) * (f32(1) - benefit), (c, c, n - 1))[one]
    t[none] = f32(p["dp_w3"])
    trans = (trans / f32(p["dp_w4"])).astype(f32)
    path = _path1(local, trans)
    return ref_pitch[path, np.arange(n)]
