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


def install() -> None:
    """satools.hifigan.archi.CoreHifiGan := satools_b200.CoreHifiGan."""
    global _ORIGINAL
    import satools.hifigan.archi as ref_archi  # the reference package must be importable
    from .archi import CoreHifiGan
    if ref_archi.CoreHifiGan is not CoreHifiGan:
        _ORIGINAL = ref_archi.CoreHifiGan
        ref_archi.CoreHifiGan = CoreHifiGan


def uninstall() -> None:
    global _ORIGINAL
    if _ORIGINAL is not None:
        import satools.hifigan.archi as ref_archi
        ref_archi.CoreHifiGan = _ORIGINAL
        _ORIGINAL = None
