"""Make the reference pick up the B200 generator without touching its sources.

The model files resolve the generator class at ``Net.__init__`` time through the module
attribute ``hifigan.archi.CoreHifiGan`` (/root/reference/egs/vc/libritts/local/tuning/
hifigan.py:45, hifigan_clean.py, hifigan_m2o.py), and ``infer_helper.load_model`` executes the
model file before building the net (infer_helper.py:49-58).  Rebinding that one attribute
before ``torch.hub.load(..., 'anonymization')`` / ``load_model`` therefore swaps the synthesis
path and nothing else; checkpoints load unchanged because the state-dict keys are identical.
"""
from __future__ import annotations

_ORIGINAL = None
_ORIGINAL_YAAPT = None


def yaapt(_in, kwargs):
    """Same call as `satools.hifigan.yaapt.yaapt(_in, kwargs)` (yaapt.py:947-952): `_in` [B, n] waveforms, `kwargs` the
    option dict (`Net.f0_yaapt_opts`); returns the pitch per frame [B, n_frames] on `_in`'s device.  The whole batch runs on
    the GPU in one pass (satools_b200.yaapt_frontend.yaapt); a CPU input is moved to the current CUDA device and the result
    moved back -- `get_f0` is registered with compute_device="cpu" in the reference (hifigan.py:118).  Unknown option names are
    ignored as `_yaapt` ignores them (it only `kwargs.get`s the names it knows)."""
    import torch
    from . import yaapt_frontend as yf
    if not torch.cuda.is_available():
        raise RuntimeError("satools_b200 YAAPT needs a CUDA device: there is no CPU implementation in this package")
    opts = {k: float(v) for k, v in dict(kwargs).items() if k in yf.OPTION_NAMES or k == "frame_lengtht"}
    x = _in if _in.dim() == 2 else _in.reshape(-1, _in.shape[-1])
    dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return yf.yaapt(x.to(dev), **opts).to(_in.device)


def install(yaapt_too: bool = False) -> None:
    """satools.hifigan.archi.CoreHifiGan := satools_b200.CoreHifiGan; with yaapt_too also
    satools.hifigan.yaapt.yaapt := satools_b200.install.yaapt (the model files call it through the module attribute,
    egs/vc/libritts/local/tuning/hifigan.py:121)."""
    global _ORIGINAL, _ORIGINAL_YAAPT
    import satools.hifigan.archi as ref_archi  # the reference package must be importable
    from .archi import CoreHifiGan
    if ref_archi.CoreHifiGan is not CoreHifiGan:
        _ORIGINAL = ref_archi.CoreHifiGan
        ref_archi.CoreHifiGan = CoreHifiGan
    if yaapt_too:
        import satools.hifigan.yaapt as ref_yaapt
        if ref_yaapt.yaapt is not yaapt:
            _ORIGINAL_YAAPT = ref_yaapt.yaapt
            ref_yaapt.yaapt = yaapt


def uninstall() -> None:
    global _ORIGINAL, _ORIGINAL_YAAPT
    if _ORIGINAL is not None:
        import satools.hifigan.archi as ref_archi
        ref_archi.CoreHifiGan = _ORIGINAL
        _ORIGINAL = None
    if _ORIGINAL_YAAPT is not None:
        import satools.hifigan.yaapt as ref_yaapt
        ref_yaapt.yaapt = _ORIGINAL_YAAPT
        _ORIGINAL_YAAPT = None
