"""Drop-in replacement for ``satools.hifigan.archi.CoreHifiGan`` running on the B200 library.

Mirrors the reference interface (/root/reference/satools/satools/hifigan/archi.py:21-116):
same constructor keywords (incl. the ``imput_dim`` spelling), same sub-module names and hence
the same 291 state-dict keys ``{conv_pre, ups.N, resblocks.M.convs{1,2}.K, conv_post}.
{weight_g, weight_v, bias}`` (strict ``load_state_dict`` of a reference checkpoint works,
infer_helper.py:57-58), ``forward(x) -> (wav, torch.empty(1))`` and ``remove_weight_norm()``.

The parameters live in ordinary torch modules, but no torch op ever computes with them:
``forward`` hands raw pointers to ``libsatools_hifigan.so`` (include/sa_hifigan.h), which folds
weight-norm once, packs the weights for the sm_100a kernels and runs the whole generator.
There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

warnings.filterwarnings("ignore", message=r".*weight_norm is deprecated.*")
from torch.nn.utils import remove_weight_norm, weight_norm  # noqa: E402  (old-style, as the reference: archi.py:4)

_TORCH_DTYPE = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16,
                torch.bfloat16: _lib.DTYPE_BF16, torch.float64: _lib.DTYPE_F64}
_OUT_DTYPE = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.int16: _lib.DTYPE_PCM16}


def _redraw(conv: nn.Module) -> None:
    # The reference calls init_weights (normal(0, 0.01) on `.weight`, nn.py:11-14) after
    # weight_norm has wrapped the conv.  That write is overwritten by the weight_norm hook on
    # the next forward, i.e. it changes no effective weight -- but it advances the global RNG.
    # Drawing the same amount here keeps "same seed -> same random weights as the reference".
    conv.weight.data.normal_(0.0, 0.01)


class ResBlock1(nn.Module):
    """Parameter container with the layout of satools.hifigan.nn.ResBlock1 (nn.py:93-166)."""

    def __init__(self, channels: int, kernel_size: int, dilation: Sequence[int]):
        super().__init__()
        self.kernel_size = kernel_size
        self.dilation = tuple(dilation)
        self.convs1 = nn.ModuleList(
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                  padding=(kernel_size * d - d) // 2)) for d in dilation)
        for c in self.convs1:
            _redraw(c)
        self.convs2 = nn.ModuleList(
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1,
                                  padding=(kernel_size - 1) // 2)) for _ in dilation)
        for c in self.convs2:
            _redraw(c)

    def remove_weight_norm(self):
        for c in list(self.convs1) + list(self.convs2):
            remove_weight_norm(c)


class CoreHifiGan(nn.Module):
    """B200-native HiFi-GAN generator; see module docstring."""

    def __init__(
        self,
        upsample_rates=[5, 4, 4, 2, 2],
        upsample_kernel_sizes=[11, 8, 8, 4, 4],
        imput_dim=256 + 1,
        upsample_initial_channel=512,
        resblock_kernel_sizes=[3, 7, 11],
        resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        iSTFTNetout=False,
        iSTFTNet_n_fft=16,
        precision: Optional[str] = None,
    ):
        super().__init__()
        if iSTFTNetout:
            raise NotImplementedError("iSTFTNetout=True is never enabled by any SA-toolkit model file; not supported")
        self.iSTFTNetout = False
        self.resblock_kernel_sizes = list(resblock_kernel_sizes)
        self.resblock_dilation_sizes = [list(d) for d in resblock_dilation_sizes]
        self.num_kernels = len(resblock_kernel_sizes)
        self.upsample_rates = list(upsample_rates)
        self.upsample_kernel_sizes = list(upsample_kernel_sizes)
        self.imput_dim = imput_dim
        self.upsample_initial_channel = upsample_initial_channel
        self.precision = precision or os.environ.get("SATOOLS_B200_PRECISION", "fp16")

        # Same construction order as the reference so the RNG stream is consumed identically.
        self.conv_pre = weight_norm(nn.Conv1d(imput_dim, upsample_initial_channel, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
            self.ups.append(weight_norm(nn.ConvTranspose1d(
                upsample_initial_channel // (2 ** i), upsample_initial_channel // (2 ** (i + 1)),
                k, u, padding=(k - u) // 2)))
        self.resblocks = nn.ModuleList()
        ch = upsample_initial_channel
        for i in range(len(self.ups)):
            ch = upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(ResBlock1(ch, k, d))
        self.conv_post = weight_norm(nn.Conv1d(ch, 1, 7, 1, padding=3))
        for up in self.ups:
            _redraw(up)
        _redraw(self.conv_post)
        # weight_norm() leaves a non-leaf `.weight` attribute behind, which breaks deepcopy /
        # pickling (the reference works around it with fix_weight_norm_deepcopy, nn.py:177-181).
        # Nothing here ever reads it, so keep a detached copy.
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)) and "weight" in m.__dict__:
                m.weight = m.weight.detach()

        self._handle: Optional[int] = None
        self._handle_pid = -1
        self._handle_device = -1
        self._weights_sig = None
        self._finalized_precision = None
        self._workspace: Optional[torch.Tensor] = None
        self._host_scratch: Optional[torch.Tensor] = None
        self._debug_buf: Optional[torch.Tensor] = None
        self.last_launch_count = 0

    # ---- reference API -----------------------------------------------------------------
    def remove_weight_norm(self):
        """archi.py:109-115."""
        for up in self.ups:
            remove_weight_norm(up)
        for rb in self.resblocks:
            rb.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)

    def output_length(self, frames: int) -> int:
        r = 1
        for u in self.upsample_rates:
            r *= u
        return r * frames + 1

    @torch.no_grad()
    def forward(self, x: torch.Tensor, frames_per_item: Optional[Sequence[int]] = None,
                out_dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
        """x [B, imput_dim, T] on a CUDA device -> (wav [B, 1, 320*T+1], torch.empty(1)) as
        archi.py:93-107.  The autocast context the caller holds (hifigan.py:99) is ignored."""
        if not x.is_cuda:
            raise RuntimeError("satools_b200.CoreHifiGan runs on CUDA (sm_100a) only; got a CPU tensor. "
                               "There is no CPU fallback.")
        if x.dim() != 3 or x.shape[1] != self.imput_dim:
            raise ValueError(f"expected x [B, {self.imput_dim}, T], got {tuple(x.shape)}")
        lib = _lib.load()
        x = x.detach().to(torch.float32).contiguous()
        B, _, T = x.shape
        with torch.cuda.device(x.device):
            self._ensure_ready(x.device)
            need = lib.sa_hifigan_workspace_bytes(self._handle, B, T)
            ws = self._get_workspace(need, x.device)
            y = torch.empty((B, 1, self.output_length(T)), dtype=out_dtype, device=x.device)
            fpi = None
            if frames_per_item is not None:
                fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.sa_hifigan_forward(self._handle, x.data_ptr(), B, T, fpi, y.data_ptr(),
                                              _OUT_DTYPE[out_dtype], ws.data_ptr(), ws.numel(), stream))
            self.last_launch_count = int(lib.sa_hifigan_last_launch_count(self._handle))
        return (y, torch.empty((1)))

    # ---- conditioning parts instead of the concatenated tensor (Net._forward, hifigan.py:83-97) ----
    @torch.no_grad()
    def forward_parts(self, bn: torch.Tensor, f0: torch.Tensor, spk_id: torch.Tensor,
                      frames_per_item: Optional[Sequence[int]] = None,
                      out_dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
        """bn [B, n_bn, T], f0 [B, 1, T] or [B, T] (normalised, transformed, at T frames), spk_id [B, n_spk]
        (one-hot).  Equals forward(torch.cat([bn, f0, spk_id[:, :, None].expand(-1, -1, T)], 1)) bit for bit
        without building that tensor."""
        if not (bn.is_cuda and f0.is_cuda and spk_id.is_cuda):
            raise RuntimeError("satools_b200.CoreHifiGan runs on CUDA (sm_100a) only; there is no CPU fallback.")
        lib = _lib.load()
        bn = bn.detach().to(torch.float32).contiguous()
        B, n_bn, T = bn.shape
        f0 = f0.detach().to(torch.float32).reshape(B, 1, -1).contiguous()
        spk = spk_id.detach().to(torch.float32).reshape(B, -1).contiguous()
        n_spk = spk.shape[1]
        if f0.shape[2] != T or n_bn + 1 + n_spk != self.imput_dim:
            raise ValueError(f"parts do not add up: bn {tuple(bn.shape)}, f0 {tuple(f0.shape)}, spk {tuple(spk.shape)}, "
                             f"imput_dim {self.imput_dim}")
        with torch.cuda.device(bn.device):
            self._ensure_ready(bn.device)
            ws = self._get_workspace(lib.sa_hifigan_workspace_bytes(self._handle, B, T), bn.device)
            y = torch.empty((B, 1, self.output_length(T)), dtype=out_dtype, device=bn.device)
            fpi = None
            if frames_per_item is not None:
                fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            stream = torch.cuda.current_stream(bn.device).cuda_stream
            _lib.check(lib.sa_hifigan_forward_parts(self._handle, bn.data_ptr(), n_bn, f0.data_ptr(), spk.data_ptr(), n_spk,
                                                    B, T, fpi, y.data_ptr(), _OUT_DTYPE[out_dtype], ws.data_ptr(),
                                                    ws.numel(), stream))
            self.last_launch_count = int(lib.sa_hifigan_last_launch_count(self._handle))
        return (y, torch.empty((1)))

    # ---- compact conditioning: VQ code index + F0 per frame, speaker id per item ----
    def set_codebook(self, codebook: torch.Tensor) -> None:
        """codebook [n_codes, n_bn] (the ASR-BN extractor's VectorQuantizerEMA embedding, chain/nn.py:427-459).  Kept on the
        module (not a parameter: the generator's state dict stays the reference's) and uploaded with the weights."""
        cb = codebook.detach().to("cpu", torch.float32).contiguous()
        if cb.dim() != 2 or cb.shape[0] > 255 or cb.shape[1] + 1 > self.imput_dim:
            raise ValueError(f"codebook must be [n_codes <= 255, n_bn < imput_dim], got {tuple(cb.shape)}")
        self._codebook = cb
        self._codebook_handle = None                   # re-upload on the next call

    def _ensure_codebook(self) -> None:
        cb = getattr(self, "_codebook", None)
        if cb is None:
            raise RuntimeError("call set_codebook(codebook) before the compact (VQ index) entries")
        if getattr(self, "_codebook_handle", None) != self._handle:
            _lib.check(_lib.load().sa_hifigan_set_codebook(self._handle, cb.data_ptr(), cb.shape[0], cb.shape[1]))
            self._codebook_handle = self._handle

    @torch.no_grad()
    def forward_vq(self, vq_idx: torch.Tensor, f0: torch.Tensor, spk_ids: torch.Tensor,
                   frames_per_item: Optional[Sequence[int]] = None,
                   out_dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
        """vq_idx uint8 [B, T], f0 fp32 [B, T] (or [B, 1, T]), spk_ids int [B], on a CUDA device.  Equals forward(x) on
        x = cat(codebook[vq_idx], f0, one_hot(spk_ids)) bit for bit (index >= n_codes: zero BN vector)."""
        if not (vq_idx.is_cuda and f0.is_cuda and spk_ids.is_cuda):
            raise RuntimeError("satools_b200.CoreHifiGan runs on CUDA (sm_100a) only; there is no CPU fallback.")
        lib = _lib.load()
        idx = vq_idx.detach().to(torch.uint8).contiguous()
        B, T = idx.shape
        f0 = f0.detach().to(torch.float32).reshape(B, -1).contiguous()
        spk = spk_ids.detach().to(torch.int32).reshape(B).contiguous()
        if f0.shape[1] != T:
            raise ValueError(f"f0 has {f0.shape[1]} frames, vq_idx {T}")
        with torch.cuda.device(idx.device):
            self._ensure_ready(idx.device)
            self._ensure_codebook()
            ws = self._get_workspace(lib.sa_hifigan_workspace_bytes(self._handle, B, T), idx.device)
            y = torch.empty((B, 1, self.output_length(T)), dtype=out_dtype, device=idx.device)
            fpi = None
            if frames_per_item is not None:
                fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            stream = torch.cuda.current_stream(idx.device).cuda_stream
            _lib.check(lib.sa_hifigan_forward_vq(self._handle, idx.data_ptr(), f0.data_ptr(), spk.data_ptr(), B, T, fpi,
                                                 y.data_ptr(), _OUT_DTYPE[out_dtype], ws.data_ptr(), ws.numel(), stream))
            self.last_launch_count = int(lib.sa_hifigan_last_launch_count(self._handle))
        return (y, torch.empty((1)))

    @torch.no_grad()
    def vq_assign(self, bn: torch.Tensor, return_quantized: bool = False):
        """bn fp32 [..., n_bn] on a CUDA device (the bottleneck rows VectorQuantizerEMA.forward receives, chain/nn.py:402):
        returns encoding_indices uint8 [...] -- the vq_idx of forward_vq -- and, on request, the module's `quantized`
        output (inputs + (codeword - inputs), chain/nn.py:448-456)."""
        if not bn.is_cuda:
            raise RuntimeError("satools_b200.CoreHifiGan runs on CUDA (sm_100a) only; there is no CPU fallback.")
        lib = _lib.load()
        x = bn.detach().to(torch.float32).contiguous()
        with torch.cuda.device(x.device):
            self._ensure_handle(x.device)                 # the codebook only: no weight signature walk on this path
            self._ensure_codebook()
            if x.shape[-1] != self._codebook.shape[1]:
                raise ValueError(f"rows of {x.shape[-1]} values against a codebook of dimension {self._codebook.shape[1]}")
            idx = torch.empty(x.shape[:-1], dtype=torch.uint8, device=x.device)
            q = torch.empty_like(x) if return_quantized else None
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.sa_hifigan_vq_assign(self._handle, x.data_ptr(), idx.numel(), idx.data_ptr(),
                                                q.data_ptr() if q is not None else None, stream))
        return (idx, q) if return_quantized else idx

    # ---- latency path: one CUDA graph launch instead of ~50 kernel launches ---------------------
    def graphed(self, B: int, T: int, device=None, out_dtype: torch.dtype = torch.float32) -> "GraphedForward":
        """Capture forward() for one fixed input shape [B, imput_dim, T] into a CUDA graph (single utterances,
        `convert()` one file at a time, hubconf usage).  The returned callable copies x into a static buffer, replays
        the graph and returns the static output tensor (overwritten by the next call)."""
        return GraphedForward(self, B, T, device, out_dtype)

    # ---- host-buffer entry (anonymize pipeline: H2D, convert, D2H; pipeline.py:104-149) ----
    @torch.no_grad()
    def synthesize_host(self, x_host: torch.Tensor, out: Optional[torch.Tensor] = None,
                        out_dtype: torch.dtype = torch.float32, device=None,
                        frames_per_item: Optional[Sequence[int]] = None) -> torch.Tensor:
        """x_host: CPU fp32 [B, imput_dim, T] (pinned for full speed).  Returns a CPU tensor
        [B, 1, 320*T+1]; host<->device copies happen inside the C-ABI call."""
        if x_host.is_cuda:
            raise ValueError("synthesize_host takes a CPU tensor")
        lib = _lib.load()
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        x_host = x_host.to(torch.float32).contiguous()
        B, _, T = x_host.shape
        if out is None:
            out = torch.empty((B, 1, self.output_length(T)), dtype=out_dtype, pin_memory=True)
        with torch.cuda.device(device):
            self._ensure_ready(device)
            need = lib.sa_hifigan_host_scratch_bytes(self._handle, B, T, _OUT_DTYPE[out.dtype])
            if self._host_scratch is None or self._host_scratch.numel() < need or self._host_scratch.device != device:
                self._host_scratch = None
                self._host_scratch = torch.empty(need, dtype=torch.uint8, device=device)
            fpi = None
            if frames_per_item is not None:
                fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.sa_hifigan_synthesize_host(self._handle, x_host.data_ptr(), B, T, fpi, out.data_ptr(),
                                                      _OUT_DTYPE[out.dtype], self._host_scratch.data_ptr(),
                                                      self._host_scratch.numel(), stream))
            self.last_launch_count = int(lib.sa_hifigan_last_launch_count(self._handle))
        return out

    # ---- test hook: stage activations ------------------------------------------------------
    @torch.no_grad()
    def forward_with_tap(self, x: torch.Tensor, tap: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Run forward and also return activation `tap` (0 = conv_pre, 1+i = stage i) as fp32 [B,C,L]."""
        lib = _lib.load()
        B, _, T = x.shape
        with torch.cuda.device(x.device):
            self._ensure_ready(x.device)
            if tap == 0:
                c, length = self.upsample_initial_channel, T
            else:
                c, length = self.upsample_initial_channel >> tap, T
                for u in self.upsample_rates[:tap]:
                    length *= u
            buf = torch.zeros((B, c, length), dtype=torch.float32, device=x.device)
            _lib.check(lib.sa_hifigan_set_debug_tap(self._handle, tap, buf.data_ptr()))
            try:
                y, _ = self.forward(x)
                torch.cuda.synchronize(x.device)
            finally:
                lib.sa_hifigan_set_debug_tap(self._handle, 0, None)
        return y, buf

    def check(self, device=None) -> None:
        """Synchronize the current stream and raise if any enqueued generator work failed."""
        if self._handle is None:
            return
        dev = torch.device("cuda", self._handle_device)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().sa_hifigan_check(self._handle, torch.cuda.current_stream(dev).cuda_stream))

    @torch.no_grad()
    def profile(self, x: torch.Tensor, repeats: int = 1):
        """Per-launch device times of forward(x) (CUDA events on the launching stream).
        Returns a list of (tag, milliseconds) averaged over `repeats` runs; tag = 16*section + kind
        (see include/sa_hifigan.h)."""
        lib = _lib.load()
        with torch.cuda.device(x.device):
            self._ensure_ready(x.device)
            _lib.check(lib.sa_hifigan_set_profiling(self._handle, 1))
            try:
                acc = None
                for _ in range(repeats):
                    self.forward(x)
                    n = lib.sa_hifigan_get_profile(self._handle, None, None, 0)
                    ms = (C.c_float * n)()
                    tags = (C.c_int32 * n)()
                    lib.sa_hifigan_get_profile(self._handle, ms, tags, n)
                    cur = [float(v) for v in ms]
                    acc = cur if acc is None else [a + b for a, b in zip(acc, cur)]
                return [(int(t), v / repeats) for t, v in zip(tags, acc)]
            finally:
                lib.sa_hifigan_set_profiling(self._handle, 0)

    # ---- internals ---------------------------------------------------------------------
    def _cfg(self, device_index: int) -> "_lib.Cfg":
        cfg = _lib.Cfg()
        cfg.input_dim = self.imput_dim
        cfg.initial_channels = self.upsample_initial_channel
        cfg.n_stages = len(self.upsample_rates)
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernels[i] = k
        cfg.n_resblocks = len(self.resblock_kernel_sizes)
        cfg.n_dilations = len(self.resblock_dilation_sizes[0])
        for j, k in enumerate(self.resblock_kernel_sizes):
            cfg.resblock_kernels[j] = k
            if len(self.resblock_dilation_sizes[j]) != cfg.n_dilations:
                raise ValueError("all ResBlocks must have the same number of dilations")
            for m, d in enumerate(self.resblock_dilation_sizes[j]):
                cfg.resblock_dilations[j][m] = d
        cfg.device = device_index
        return cfg

    def _signature(self):
        """(storage, version) of every parameter: changes on in-place updates, .to(), load_state_dict, remove_weight_norm.
        Walking the module tree costs ~0.3 ms per call -- a third of a 5 s utterance's latency -- so the (owner dict, name,
        parameter) triples are cached and only checked for identity; a replaced or deleted parameter triggers a new walk."""
        cache = self.__dict__.get("_sig_cache")
        if cache is not None:
            for d, n, q in cache:
                if d.get(n) is not q:
                    cache = None
                    break
        if cache is None:
            cache = [(m._parameters, n, q) for m in self.modules() for n, q in m._parameters.items() if q is not None]
            self.__dict__["_sig_cache"] = cache
        return tuple((q.data_ptr(), q._version) for _, _, q in cache)

    def _ensure_handle(self, device: torch.device) -> None:
        """The native handle on `device` (no weight upload: enough for the entries that only use the codebook)."""
        lib = _lib.load()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and (self._handle_pid != os.getpid() or self._handle_device != idx):
            if self._handle_pid == os.getpid():
                lib.sa_hifigan_destroy(self._handle)
            self._handle = None                      # a forked child never touches the parent's handle
        if self._handle is None:
            out = C.c_void_p()
            cfg = self._cfg(idx)
            _lib.check(lib.sa_hifigan_create(C.byref(cfg), C.byref(out)))
            self._handle, self._handle_pid, self._handle_device = out.value, os.getpid(), idx
            self._weights_sig = None
            self._codebook_handle = None

    def _ensure_ready(self, device: torch.device) -> None:
        lib = _lib.load()
        self._ensure_handle(device)
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {self.precision!r}")
        sig = self._signature()
        if sig != self._weights_sig or self._finalized_precision != self.precision:
            if sig != self._weights_sig:
                for name, p in self.named_parameters():
                    t = p.detach()
                    if t.dtype not in _TORCH_DTYPE:
                        t = t.float()
                    t = t.contiguous()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(lib.sa_hifigan_set_weight(self._handle, name.encode(), t.data_ptr(), shape,
                                                         t.dim(), _TORCH_DTYPE[t.dtype]))
            _lib.check(lib.sa_hifigan_finalize(self._handle, _lib.PRECISIONS[self.precision]))
            self._weights_sig = sig
            self._finalized_precision = self.precision

    def _get_workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        ws = self._workspace
        if ws is None or ws.numel() < nbytes or ws.device != device:
            self._workspace = None                   # release before growing
            self._workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        return self._workspace

    def release(self) -> None:
        """Free the native handle and cached device buffers."""
        if self._handle is not None and self._handle_pid == os.getpid():
            _lib.load().sa_hifigan_destroy(self._handle)
        self._handle = None
        self._workspace = None
        self._host_scratch = None
        self._weights_sig = None
        self._codebook_handle = None

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle_pid == os.getpid():
                _lib.load().sa_hifigan_destroy(self._handle)
        except Exception:
            pass

    def __getstate__(self):
        # Pickling / deepcopy (DataLoader workers capture the model, pipeline.py:175): the native
        # handle and device scratch stay behind.
        state = self.__dict__.copy()
        for k in ("_handle", "_workspace", "_host_scratch", "_weights_sig", "_finalized_precision", "_debug_buf",
                  "_codebook_handle"):
            state[k] = None
        state["_handle_pid"] = -1
        return state


class GraphedForward:
    """CoreHifiGan.forward for one fixed shape, captured once and replayed (see CoreHifiGan.graphed)."""

    def __init__(self, gen: CoreHifiGan, B: int, T: int, device=None, out_dtype: torch.dtype = torch.float32):
        self.gen = gen
        self.device = torch.device(device) if device is not None else next(gen.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device (no CPU fallback)")
        self.x = torch.zeros((B, gen.imput_dim, T), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up: weight fold, kernel attributes, workspace
                gen.forward(self.x, out_dtype=out_dtype)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._sig = gen._weights_sig
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.y, _ = gen.forward(self.x, out_dtype=out_dtype)
        self.launches = gen.last_launch_count
        self._ws = gen._workspace                         # the captured kernels point into this allocation

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if tuple(x.shape) != tuple(self.x.shape):
            raise ValueError(f"graph was captured for {tuple(self.x.shape)}, got {tuple(x.shape)}")
        if self.gen._signature() != self._sig:
            raise RuntimeError("the generator's weights changed after the capture (folded weights are baked in); "
                               "call CoreHifiGan.graphed() again")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y, torch.empty((1))
