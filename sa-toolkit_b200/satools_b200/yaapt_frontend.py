"""Front end of the reference's YAAPT F0 extractor on the GPU, for a whole batch at once (first step).

Mirrors what `_yaapt` (/root/reference/satools/satools/hifigan/yaapt.py:873-899) computes before its trackers run, with the
reference's names: `SignalObj.filtered` of the signal and of the squared signal (`filtered_version`, lines 42-52), and
`PitchObj.energy / vuv / mean_energy / nframes` as `nlfer` (lines 148-176) and `set_energy` (124-127) leave them.  Options are
the `_yaapt` keyword options (`frame_length`, `frame_space`, `f0_min`, `f0_max`, `fft_length`, `bp_low`, `bp_high`,
`nlfer_thresh1`; `bin/pipeline.py` passes frame_length 35 / frame_space 20).

The compute runs in libsatools_hifigan.so (csrc/yaapt_frontend.cu through the C ABI of include/sa_yaapt.h); there is no CPU
path here.  The spectral / temporal trackers and the dynamic programming of YAAPT are not on the GPU yet.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

OPTION_NAMES = ("sr", "frame_length", "frame_space", "f0_min", "f0_max", "fft_length", "bp_low", "bp_high", "nlfer_thresh1",
                "shc_numharms", "shc_window", "shc_pwidth", "shc_maxpeaks", "shc_thresh1", "shc_thresh2", "f0_double", "f0_half",
                "merit_extra", "median_value", "dp5_k1", "spec_pitch_min_std", "tda_frame_length", "nccf_thresh1", "nccf_thresh2",
                "nccf_maxcands", "nccf_pwidth", "merit_boost", "nlfer_thresh2", "merit_pivot", "dp_w1", "dp_w2", "dp_w3", "dp_w4")


def params(**kwargs) -> "_lib.YaaptParams":
    lib = _lib.load()
    p = _lib.YaaptParams()
    if lib.sa_yaapt_default_params(p) != 0:
        raise _lib.SaHifiganError(lib.sa_yaapt_last_error().decode())
    for k, v in kwargs.items():
        if k not in OPTION_NAMES:
            raise KeyError(f"unknown YAAPT option {k!r}")
        setattr(p, k, float(v))
    return p


def num_frames(n_samples: int, **kwargs) -> int:
    """`pitch.nframes` for an utterance of n_samples (yaapt.py:164-166, 131)."""
    return int(_lib.load().sa_yaapt_num_frames(params(**kwargs), int(n_samples)))


@dataclass
class FrontEnd:
    filtered: torch.Tensor        # [B, n_max + 2 pad] float32: SignalObj.filtered (zero beyond an item's padded length)
    filtered_nl: torch.Tensor     # [B, n_max + 2 pad] float32: the squared signal's
    energy: torch.Tensor          # [B, F_max] float32: PitchObj.energy (0 beyond an item's frames)
    vuv: torch.Tensor             # [B, F_max] bool: PitchObj.vuv
    mean_energy: torch.Tensor     # [B] float32: PitchObj.mean_energy
    nframes: list                 # per item: PitchObj.nframes
    padded_lengths: list          # per item: SignalObj.size


def nlfer(wav: torch.Tensor, lengths: Optional[Sequence[int]] = None, **kwargs) -> FrontEnd:
    """wav: [B, n] (or [n]) float32 on a CUDA device, item b valid in [0, lengths[b])."""
    if wav.dim() == 1:
        wav = wav.unsqueeze(0)
    if not wav.is_cuda:
        raise ValueError("yaapt_frontend.nlfer needs a CUDA tensor: this path has no CPU implementation")
    lib = _lib.load()
    p = params(**kwargs)
    wav = wav.contiguous().float()
    B, n = wav.shape
    n_pad = int(lib.sa_yaapt_padded_length(p, n))
    f_max = int(lib.sa_yaapt_num_frames(p, n))
    if n_pad < 0 or f_max < 0:
        raise _lib.SaHifiganError(lib.sa_yaapt_last_error().decode())
    dev = wav.device
    with torch.cuda.device(dev):
        out = FrontEnd(torch.empty(B, n_pad, device=dev), torch.empty(B, n_pad, device=dev), torch.empty(B, f_max, device=dev),
                       torch.empty(B, f_max, dtype=torch.uint8, device=dev), torch.empty(B, device=dev), [], [])
        ws = torch.empty(int(lib.sa_yaapt_frontend_workspace_bytes(p, B, n)), dtype=torch.uint8, device=dev)
        lens = None
        if lengths is not None:
            if len(lengths) != B:
                raise ValueError("lengths must have one entry per item")
            lens = (C.c_int32 * B)(*[int(v) for v in lengths])
        rc = lib.sa_yaapt_frontend(p, wav.data_ptr(), B, n, lens, out.filtered.data_ptr(), out.filtered_nl.data_ptr(),
                                   out.energy.data_ptr(), out.vuv.data_ptr(), out.mean_energy.data_ptr(), ws.data_ptr(), ws.numel(),
                                   torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise _lib.SaHifiganError(f"sa_yaapt_frontend: {lib.sa_yaapt_last_error().decode()}")
        torch.cuda.current_stream(dev).synchronize()          # `ws` and `lens` are released when this returns
    out.vuv = out.vuv.bool()
    for b in range(B):
        nb = n if lengths is None else int(lengths[b])
        out.nframes.append(int(lib.sa_yaapt_num_frames(p, nb)))
        out.padded_lengths.append(int(lib.sa_yaapt_padded_length(p, nb)))
    return out


def spec_shc(front: FrontEnd, lengths: Optional[Sequence[int]] = None, candidates: bool = False, **kwargs):
    """The SHC vectors `spec_track` (yaapt.py:184-231) hands to `peaks`, for every voiced frame of the batch:
    [B, F_max, max_SHC] float32 (zero rows for unvoiced frames).  `front` is the result of `nlfer` with the same
    lengths and options.  candidates=True: returns (shc, cand_pitch, cand_merit), the last two [B, maxpeaks, F_max] as
    `spec_track` fills them from `peaks` (yaapt.py:204-205, 231, 383-497)."""
    lib = _lib.load()
    p = params(**kwargs)
    x = front.filtered_nl
    B, n_pad = x.shape
    n = n_pad - (int(lib.sa_yaapt_padded_length(p, 0)))
    f_max = front.vuv.shape[1]
    k = int(lib.sa_yaapt_shc_length(p))
    if k < 0 or int(lib.sa_yaapt_num_frames(p, n)) != f_max:
        raise _lib.SaHifiganError(f"spec_shc: options do not match the front end ({lib.sa_yaapt_last_error().decode()})")
    dev = x.device
    with torch.cuda.device(dev):
        out = torch.empty(B, f_max, k, device=dev)
        m = int(p.shc_maxpeaks)
        cp = torch.empty(B, m, f_max, device=dev) if candidates else None
        cm = torch.empty(B, m, f_max, device=dev) if candidates else None
        vuv = front.vuv.to(torch.uint8).contiguous()
        ws = torch.empty(int(lib.sa_yaapt_shc_workspace_bytes(p, B, n)), dtype=torch.uint8, device=dev)
        lens = None
        if lengths is not None:
            lens = (C.c_int32 * B)(*[int(v) for v in lengths])
        rc = lib.sa_yaapt_shc(p, x.data_ptr(), B, n, lens, vuv.data_ptr(), out.data_ptr(), cp.data_ptr() if candidates else None,
                              cm.data_ptr() if candidates else None, ws.data_ptr(), ws.numel(),
                              torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise _lib.SaHifiganError(f"sa_yaapt_shc: {lib.sa_yaapt_last_error().decode()}")
        torch.cuda.current_stream(dev).synchronize()
    return (out, cp, cm) if candidates else out


def spec_track_from_candidates(cand_pitch: torch.Tensor, cand_merit: torch.Tensor, n_samples: int,
                               lengths: Optional[Sequence[int]] = None, **kwargs):
    """The per-utterance part of `spec_track` (yaapt.py:233-316) on candidate matrices [B, maxpeaks, F_max] (CUDA tensors):
    (spec_pitch [B, F_max], pitch_std [B]).  n_samples: samples per row of the waveform batch the candidates came from."""
    lib = _lib.load()
    p = params(**kwargs)
    cp, cm = cand_pitch.contiguous().float(), cand_merit.contiguous().float()
    B, m, f_max = cp.shape
    if m != int(p.shc_maxpeaks) or f_max != int(lib.sa_yaapt_num_frames(p, int(n_samples))):
        raise ValueError("candidate matrices do not match shc_maxpeaks / the frame count of n_samples")
    nfr = [int(lib.sa_yaapt_num_frames(p, int(v))) for v in (lengths if lengths is not None else [n_samples] * B)]
    if min(nfr) < 4:
        raise IndexError("spec_track: an utterance has fewer than four frames (the reference fails the same way, yaapt.py:311)")
    dev = cp.device
    with torch.cuda.device(dev):
        spec = torch.empty(B, f_max, device=dev)
        std = torch.empty(B, device=dev)
        ws = torch.empty(int(lib.sa_yaapt_spec_track_workspace_bytes(p, B, int(n_samples))), dtype=torch.uint8, device=dev)
        lens = (C.c_int32 * B)(*[int(v) for v in lengths]) if lengths is not None else None
        rc = lib.sa_yaapt_spec_track(p, cp.data_ptr(), cm.data_ptr(), B, int(n_samples), lens, spec.data_ptr(), std.data_ptr(),
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise _lib.SaHifiganError(f"sa_yaapt_spec_track: {lib.sa_yaapt_last_error().decode()}")
        torch.cuda.current_stream(dev).synchronize()
    return spec, std


def spec_track(front: FrontEnd, lengths: Optional[Sequence[int]] = None, **kwargs):
    """`spec_track(nonlinear_sign, pitch, parameters)` of the reference (yaapt.py:184-316) for the whole batch:
    (spec_pitch [B, F_max] float32, pitch_std [B] float32).  Like the reference it raises IndexError for an utterance of
    fewer than four frames (yaapt.py:311)."""
    lib = _lib.load()
    p = params(**kwargs)
    if min(front.nframes) < 4:
        raise IndexError("spec_track: an utterance has fewer than four frames (the reference fails the same way, yaapt.py:311)")
    _, cp, cm = spec_shc(front, lengths=lengths, candidates=True, **kwargs)
    n = front.filtered_nl.shape[1] - int(lib.sa_yaapt_padded_length(p, 0))
    return spec_track_from_candidates(cp, cm, n, lengths=lengths, **kwargs)


def yaapt(wav: torch.Tensor, lengths: Optional[Sequence[int]] = None, **kwargs) -> torch.Tensor:
    """`yaapt(_in, kwargs)` of the reference (yaapt.py:947-952) for a batch on the GPU: the final pitch track per frame,
    [B, F_max] float32 (0 = unvoiced; zero beyond an item's frames), i.e. `pitch.samp_values` of every utterance.
    `frame_lengtht` is accepted for `tda_frame_length` as in `_yaapt` (lines 803-810)."""
    if "frame_lengtht" in kwargs:
        v = kwargs.pop("frame_lengtht")
        kwargs.setdefault("tda_frame_length", v)
    lib = _lib.load()
    p = params(**kwargs)
    front = nlfer(wav, lengths=lengths, **kwargs)
    spec, std = spec_track(front, lengths=lengths, **kwargs)
    B, f_max = spec.shape
    n = front.filtered.shape[1] - int(lib.sa_yaapt_padded_length(p, 0))
    dev = spec.device
    with torch.cuda.device(dev):
        final = torch.empty(B, f_max, device=dev)
        vuv = front.vuv.to(torch.uint8).contiguous()
        ws = torch.empty(int(lib.sa_yaapt_track_workspace_bytes(p, B, n)), dtype=torch.uint8, device=dev)
        lens = (C.c_int32 * B)(*[int(v) for v in lengths]) if lengths is not None else None
        rc = lib.sa_yaapt_track(p, front.filtered.data_ptr(), front.filtered_nl.data_ptr(), front.energy.data_ptr(), vuv.data_ptr(),
                                spec.data_ptr(), std.data_ptr(), B, n, lens, final.data_ptr(), ws.data_ptr(), ws.numel(),
                                torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise _lib.SaHifiganError(f"sa_yaapt_track: {lib.sa_yaapt_last_error().decode()}")
        torch.cuda.current_stream(dev).synchronize()
    return final
