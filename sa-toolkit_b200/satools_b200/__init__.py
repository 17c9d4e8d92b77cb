"""satools_b200 -- B200 (sm_100a) implementation of SA-toolkit's HiFi-GAN synthesis hot path.

Public surface:
  CoreHifiGan          drop-in for satools.hifigan.archi.CoreHifiGan (archi.py)
  install()            rebind satools.hifigan.archi.CoreHifiGan to it (install.py)
  scheduler            length-balanced utterance sharding for multi-GPU runs (scheduler.py)
  synth                host-side batch driver: bucket, pad, convert, trim (synth.py)
  HostPipeline         two-deep H2D / generator / D2H pipeline over host batches (pipeline.py)
  GraphedForward       CoreHifiGan.graphed(B, T): forward for one fixed shape as a single CUDA graph launch (archi.py)
  yaapt_frontend       batched GPU front end of the YAAPT F0 extractor: band-pass biquads, NLFER energy, voiced flags
"""
from .archi import CoreHifiGan, GraphedForward, ResBlock1  # noqa: F401
from .install import install, uninstall  # noqa: F401
from . import scheduler  # noqa: F401
from . import synth  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from . import yaapt_frontend  # noqa: F401

__all__ = ["CoreHifiGan", "ResBlock1", "install", "uninstall", "scheduler", "synth", "HostPipeline", "GraphedForward", "yaapt_frontend"]
