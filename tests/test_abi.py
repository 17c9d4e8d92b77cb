"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares; the Python shim mirrors the reference interface.  No compute calls here."""
import copy
import os
import pickle
import re

import pytest
import torch

import helpers
from satools_b200 import _lib, CoreHifiGan

HEADER = os.path.join(helpers.ROOT, "include", "sa_hifigan.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_hifigan_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sa_hifigan.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and header disagree"
    assert lib.sa_hifigan_abi_version() == 1


def test_default_cfg_is_the_reference_generator():
    lib = _lib.load()
    cfg = _lib.Cfg()
    assert lib.sa_hifigan_default_cfg(cfg) == 0
    assert cfg.input_dim == 504 and cfg.initial_channels == 512 and cfg.n_stages == 5
    assert list(cfg.upsample_rates)[:5] == [5, 4, 4, 2, 2]
    assert list(cfg.upsample_kernels)[:5] == [11, 8, 8, 4, 4]
    assert list(cfg.resblock_kernels)[:3] == [3, 7, 11]
    assert [list(r)[:3] for r in cfg.resblock_dilations][:3] == [[1, 3, 5]] * 3


def test_null_arguments_are_errors_not_crashes():
    lib = _lib.load()
    assert lib.sa_hifigan_default_cfg(None) < 0
    assert b"NULL" in lib.sa_hifigan_last_error()
    assert lib.sa_hifigan_create(None, None) < 0
    assert lib.sa_hifigan_workspace_bytes(None, 1, 1) == 0
    lib.sa_hifigan_destroy(None)


def test_state_dict_keys_match_reference_layout():
    gen = helpers.seeded_generator(0)
    keys = list(gen.state_dict().keys())
    assert len(keys) == 291
    assert keys[:3] == ["conv_pre.bias", "conv_pre.weight_g", "conv_pre.weight_v"]
    assert "ups.4.weight_v" in keys and "resblocks.14.convs2.2.weight_g" in keys and "conv_post.bias" in keys
    sd = gen.state_dict()
    assert tuple(sd["ups.0.weight_g"].shape) == (512, 1, 1)          # ConvTranspose1d: norm axis = Cin
    assert tuple(sd["ups.0.weight_v"].shape) == (512, 256, 11)
    assert tuple(sd["conv_pre.weight_v"].shape) == (512, 504, 7)
    assert tuple(sd["resblocks.14.convs1.2.weight_v"].shape) == (16, 16, 11)
    other = CoreHifiGan(imput_dim=504)
    other.load_state_dict(sd, strict=True)                            # infer_helper.py:57-58


def test_remove_weight_norm_changes_keys_like_the_reference():
    gen = copy.deepcopy(helpers.seeded_generator(1))
    gen.remove_weight_norm()
    keys = set(gen.state_dict().keys())
    assert "conv_pre.weight" in keys and "conv_pre.weight_g" not in keys
    assert len(keys) == 2 * 97


def test_cpu_tensor_raises_no_fallback():
    gen = helpers.seeded_generator(0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gen(torch.zeros(1, 504, 8))


def test_istftnet_head_rejected():
    with pytest.raises(NotImplementedError):
        CoreHifiGan(imput_dim=504, iSTFTNetout=True)


def test_pickle_and_deepcopy_drop_native_state():
    gen = helpers.seeded_generator(0)
    clone = pickle.loads(pickle.dumps(gen))
    assert clone._handle is None
    assert helpers.state_sha256(clone.state_dict()) == helpers.state_sha256(gen.state_dict())


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsatools_hifigan.so")
    with pytest.raises(_lib.SaHifiganError, match="no CPU fallback"):
        _lib.load()
