"""Two-deep host pipeline over the stream-ordered host entries of the C ABI.

The reference keeps the GPU fed through a DataLoader that prefetches the next batch while the current
one is converted, then does `.to(device)`, convert and `.cpu()` back to back
(/root/reference/satools/satools/bin/pipeline.py:91-101,104-149).  Here each batch is one stream-ordered
C-ABI call (H2D copy -> generator -> D2H copy); two slots with their own stream, device scratch and
pinned output alternate, so the copies of one batch run under the kernels of the other.

    pipe = HostPipeline(gen)
    t0 = pipe.submit(x0_pinned)
    t1 = pipe.submit(x1_pinned)          # H2D of batch 1 overlaps the kernels of batch 0
    y0 = pipe.result(t0)                 # waits for slot 0 only

Three input forms: submit (x [B, Cin, T]), submit_parts (bn, f0, speaker one-hot), submit_vq (VQ code index, f0,
speaker id: 5 bytes per frame).  With trimmed=True (submit, submit_vq) only the kept samples of every item come back,
packed one item after the other -- `offsets(frames)` gives the item boundaries (bin/pipeline.py:156 trims on the host).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from .archi import _OUT_DTYPE


class _Slot:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.scratch: Optional[torch.Tensor] = None
        self.ticket = -1                   # ticket in flight (or finished, not yet collected) in this slot
        self.keep = None                   # host objects the enqueued call still reads / writes
        self.out: Optional[torch.Tensor] = None


def trimmed_offsets(gen, frames_per_item: Sequence[int]) -> List[int]:
    """Sample offsets of the items in a trimmed result: item b is out[off[b]:off[b + 1]] (320 * frames + 1 samples)."""
    off = [0]
    for f in frames_per_item:
        off.append(off[-1] + gen.output_length(int(f)))
    return off


class HostPipeline:
    def __init__(self, gen, depth: int = 2, device=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.gen = gen
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device (no CPU fallback)")
        self._slots: List[_Slot] = [_Slot(self.device) for _ in range(depth)]
        self._next = 0
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ---- shared plumbing ----------------------------------------------------------------------------------
    def _begin(self):
        ticket = self._next
        slot = self._slots[ticket % len(self._slots)]
        if slot.ticket >= 0:
            slot.stream.synchronize()                       # the slot's previous batch (result() may be skipped)
        return ticket, slot

    def _out(self, out, B, T, out_dtype, frames_per_item, trimmed):
        if trimmed:
            if frames_per_item is None:
                raise ValueError("trimmed output needs frames_per_item")
            n = sum(self.gen.output_length(int(f)) for f in frames_per_item)
            if out is None:
                out = torch.empty((n,), dtype=out_dtype, pin_memory=True)
            elif out.numel() < n:
                raise ValueError(f"out holds {out.numel()} samples, the trimmed batch needs {n}")
            return out, n
        if out is None:
            out = torch.empty((B, 1, self.gen.output_length(T)), dtype=out_dtype, pin_memory=True)
        return out, out.numel()

    def _scratch(self, slot, lib, h, B, T, out):
        need = lib.sa_hifigan_host_scratch_bytes(h, B, T, _OUT_DTYPE[out.dtype])
        if slot.scratch is None or slot.scratch.numel() < need:
            slot.scratch = None
            slot.scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return slot.scratch

    def _finish(self, ticket, slot, keep, out, lib, h, h2d, d2h):
        self.launches += int(lib.sa_hifigan_last_launch_count(h))
        self.h2d_bytes += int(h2d)
        self.d2h_bytes += int(d2h)
        slot.ticket, slot.keep, slot.out = ticket, keep, out
        self._next += 1
        return ticket

    # ---- input forms --------------------------------------------------------------------------------------
    def submit(self, x_host: torch.Tensor, out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32,
               frames_per_item: Optional[Sequence[int]] = None, trimmed: bool = False) -> int:
        """Enqueue one batch (CPU fp32 [B, imput_dim, T], pinned for overlap).  Returns a ticket for result().
        `out` (pinned CPU) and x_host must not be touched until result(ticket) returned."""
        if x_host.is_cuda:
            raise ValueError("submit takes a CPU tensor")
        lib = _lib.load()
        ticket, slot = self._begin()
        x_host = x_host.to(torch.float32).contiguous()
        B, _, T = x_host.shape
        out, n_out = self._out(out, B, T, out_dtype, frames_per_item, trimmed)
        with torch.cuda.device(self.device):
            self.gen._ensure_ready(self.device)
            h = self.gen._handle
            scratch = self._scratch(slot, lib, h, B, T, out)
            fpi = None
            if frames_per_item is not None:
                fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            fn = lib.sa_hifigan_synthesize_host_trimmed_async if trimmed else lib.sa_hifigan_synthesize_host_async
            _lib.check(fn(h, x_host.data_ptr(), B, T, fpi, out.data_ptr(), _OUT_DTYPE[out.dtype], scratch.data_ptr(),
                          scratch.numel(), slot.stream.cuda_stream))
        return self._finish(ticket, slot, (x_host, fpi), out, lib, h, x_host.numel() * 4, n_out * out.element_size())

    def submit_parts(self, bn: torch.Tensor, f0: torch.Tensor, spk_id: torch.Tensor, out: Optional[torch.Tensor] = None,
                     out_dtype: torch.dtype = torch.float32, frames_per_item: Optional[Sequence[int]] = None) -> int:
        """Like submit(), fed with the conditioning parts (CPU, pinned for overlap): bn [B, n_bn, T], f0 [B, 1, T] or
        [B, T], spk_id [B, n_spk] one-hot -- about half the H2D bytes of the concatenated tensor."""
        if bn.is_cuda or f0.is_cuda or spk_id.is_cuda:
            raise ValueError("submit_parts takes CPU tensors")
        lib = _lib.load()
        ticket, slot = self._begin()
        bn = bn.to(torch.float32).contiguous()
        B, n_bn, T = bn.shape
        f0 = f0.to(torch.float32).reshape(B, 1, -1).contiguous()
        spk = spk_id.to(torch.float32).reshape(B, -1).contiguous()
        if f0.shape[2] != T:
            raise ValueError(f"f0 has {f0.shape[2]} frames, bn {T}")
        out, n_out = self._out(out, B, T, out_dtype, frames_per_item, False)
        with torch.cuda.device(self.device):
            self.gen._ensure_ready(self.device)
            h = self.gen._handle
            scratch = self._scratch(slot, lib, h, B, T, out)
            fpi = None
            if frames_per_item is not None:
                fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            _lib.check(lib.sa_hifigan_synthesize_host_parts_async(h, bn.data_ptr(), n_bn, f0.data_ptr(), spk.data_ptr(),
                                                                  spk.shape[1], B, T, fpi, out.data_ptr(),
                                                                  _OUT_DTYPE[out.dtype], scratch.data_ptr(),
                                                                  scratch.numel(), slot.stream.cuda_stream))
        return self._finish(ticket, slot, (bn, f0, spk, fpi), out, lib, h, (bn.numel() + f0.numel() + spk.numel()) * 4,
                            n_out * out.element_size())

    def submit_vq(self, vq_idx: torch.Tensor, f0: torch.Tensor, spk_ids: torch.Tensor, frames_per_item: Sequence[int],
                  out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.int16) -> int:
        """Compact conditioning (gen.set_codebook first): vq_idx uint8 [B, T], f0 fp32 [B, T], spk_ids int32 [B], CPU,
        pinned for overlap.  The result is always trimmed (see module docstring)."""
        if vq_idx.is_cuda or f0.is_cuda or spk_ids.is_cuda:
            raise ValueError("submit_vq takes CPU tensors")
        lib = _lib.load()
        ticket, slot = self._begin()
        idx = vq_idx.to(torch.uint8).contiguous()
        B, T = idx.shape
        f0 = f0.to(torch.float32).reshape(B, -1).contiguous()
        spk = spk_ids.to(torch.int32).reshape(B).contiguous()
        if f0.shape[1] != T:
            raise ValueError(f"f0 has {f0.shape[1]} frames, vq_idx {T}")
        out, n_out = self._out(out, B, T, out_dtype, frames_per_item, True)
        with torch.cuda.device(self.device):
            self.gen._ensure_ready(self.device)
            self.gen._ensure_codebook()
            h = self.gen._handle
            scratch = self._scratch(slot, lib, h, B, T, out)
            fpi = (C.c_int32 * B)(*[int(v) for v in frames_per_item])
            _lib.check(lib.sa_hifigan_synthesize_host_vq_trimmed_async(h, idx.data_ptr(), f0.data_ptr(), spk.data_ptr(), B, T,
                                                                       fpi, out.data_ptr(), _OUT_DTYPE[out.dtype],
                                                                       scratch.data_ptr(), scratch.numel(),
                                                                       slot.stream.cuda_stream))
        return self._finish(ticket, slot, (idx, f0, spk, fpi), out, lib, h, idx.numel() + f0.numel() * 4 + spk.numel() * 4,
                            n_out * out.element_size())

    # ---- results ------------------------------------------------------------------------------------------
    def result(self, ticket: int) -> torch.Tensor:
        """Wait for the batch of `ticket` and return its waveform tensor (CPU)."""
        slot = self._slots[ticket % len(self._slots)]
        if slot.ticket != ticket:
            raise KeyError(f"ticket {ticket} is not in flight (already overwritten or never submitted)")
        slot.stream.synchronize()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().sa_hifigan_check(self.gen._handle, slot.stream.cuda_stream))
        out, slot.keep, slot.out, slot.ticket = slot.out, None, None, -1
        return out

    def drain(self) -> None:
        for slot in self._slots:
            if slot.ticket >= 0:
                slot.stream.synchronize()
                slot.keep, slot.out, slot.ticket = None, None, -1
