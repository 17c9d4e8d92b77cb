"""CPU test of the drop-in seam: satools_b200.install() makes the REFERENCE's own model file
build its Net on satools_b200.CoreHifiGan, a reference-built state dict loads strictly, uninstall() restores.

Needs the reference tree (build container only; /root/reference does not exist on the GPU box): skipped there.
Recipe of SURVEY Appendix C: the model file is executed by path exactly as infer_helper.load_model does
(/root/reference/satools/satools/infer_helper.py:49-58), with load_model patched to return a random-init BN extractor
because released checkpoints need the network.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "satools", "satools")),
                                reason="reference tree not present (GPU box)")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def ref_env():
    os.environ.setdefault("SA_JIT_TWEAK", "true")
    sys.dont_write_bytecode = True
    added = os.path.join(REF, "satools")
    sys.path.insert(0, added)
    threads = torch.get_num_threads()
    import satools  # noqa: F401  (sets torch threads to 1, yaapt.py:27)
    torch.set_num_threads(threads)
    import satools.infer_helper
    tdnnf = _load(f"{REF}/egs/asr/librispeech/local/chain/tuning/tdnnf_vq.py", "tdnnf_vq_cfg_t")
    orig_load = satools.infer_helper.load_model
    satools.infer_helper.load_model = lambda *a, **k: tdnnf.build(
        types.SimpleNamespace(freeze_encoder="False", codebook_size=48))(output_dim=3280)
    yield
    satools.infer_helper.load_model = orig_load
    sys.path.remove(added)


def _build_net(tag):
    hf = _load(f"{REF}/egs/vc/libritts/local/tuning/hifigan.py", "hifigan_cfg_" + tag)
    Net = hf.build(types.SimpleNamespace(asrbn_model="x", f0_transformation=""))
    net = Net(utt2spk={f"u{i}": str(1000 + i) for i in range(247)})
    net.eval()
    return net


def test_install_swaps_the_generator_class_behind_the_reference_net(ref_env):
    import satools.hifigan.archi as ref_archi
    import satools_b200
    RefGen = ref_archi.CoreHifiGan
    torch.manual_seed(5)
    ref_net = _build_net("ref")                               # hifigan.py:45-49 on the reference class
    assert type(ref_net.hifigan) is RefGen
    ref_state = ref_net.state_dict()

    satools_b200.install()
    try:
        assert ref_archi.CoreHifiGan is satools_b200.CoreHifiGan
        satools_b200.install()                                # idempotent
        torch.manual_seed(5)
        net = _build_net("b200")
        assert isinstance(net.hifigan, satools_b200.CoreHifiGan)
        assert net.hifigan.imput_dim == 256 + 1 + 247         # hifigan.py:45-46
        # same parameter names, shapes and (same seed, same RNG draws) values as the reference-built Net
        state = net.state_dict()
        assert list(state.keys()) == list(ref_state.keys())
        for k in state:
            assert state[k].shape == ref_state[k].shape, k
        gen_keys = [k for k in state if k.startswith("hifigan.")]
        assert len(gen_keys) == 291
        for k in gen_keys:
            assert torch.equal(state[k], ref_state[k]), k
        # strict load of a reference checkpoint (infer_helper.py:57-58)
        perturbed = {k: (v + 0.25 if k == "hifigan.conv_post.bias" else v) for k, v in ref_state.items()}
        missing, unexpected = net.load_state_dict(perturbed, strict=True)
        assert not missing and not unexpected
        assert torch.equal(net.hifigan.conv_post.bias, ref_state["hifigan.conv_post.bias"] + 0.25)
        # the reference's call sites: remove_weight_norm (hifigan.py:51-52) and CPU input -> loud failure, no fallback
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            net.hifigan(torch.zeros(1, 504, 4))
        net.remove_weight_norm()
        assert "hifigan.conv_pre.weight" in net.state_dict()
    finally:
        satools_b200.uninstall()
    assert ref_archi.CoreHifiGan is RefGen
    satools_b200.uninstall()                                  # harmless when nothing is installed
    assert ref_archi.CoreHifiGan is RefGen
    assert type(_build_net("again").hifigan) is RefGen


def test_conditioning_assembly_of_the_reference_net_reaches_the_drop_in_unchanged(ref_env):
    """Net._forward (hifigan.py:83-102) builds x and calls self.hifigan(x) under autocast; with the drop-in installed
    the tensor that arrives at the boundary is the one the reference class would have seen (captured, no GPU needed)."""
    import numpy as np
    import satools_b200
    from satools_b200 import conditioning
    rng = np.random.default_rng(5)
    T = 16
    bn = torch.from_numpy(conditioning.codebook()[rng.integers(48, size=(2, T))]).permute(0, 2, 1).contiguous()
    f0 = torch.from_numpy((rng.random((2, T)) * 120 + 80).astype(np.float32))
    f0[:, 3:6] = 0.0
    seen = {}
    for tag in ("ref", "b200"):
        if tag == "b200":
            satools_b200.install()
        try:
            torch.manual_seed(3)
            net = _build_net("cap_" + tag)
            spk = net.get_spk_id(None, target=["1003", "1100"])
            net.hifigan.forward = lambda x, _t=tag: (seen.__setitem__(_t, x.detach().clone()),
                                                      (torch.zeros(x.shape[0], 1, 320 * x.shape[2] + 1), torch.empty(1)))[1]
            with torch.no_grad():
                y = net._forward(f0.clone(), bn, spk)
            assert y.dtype == torch.float32 and tuple(y.shape) == (2, 1, 320 * T + 1)
        finally:
            if tag == "b200":
                satools_b200.uninstall()
    assert torch.equal(seen["ref"], seen["b200"])
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "net_forward.npz"))
    np.testing.assert_array_equal(seen["b200"].numpy(), z["plain/x"])


def test_install_can_rebind_the_f0_extractor_too(ref_env):
    """`Net.get_f0` calls `hifigan.yaapt.yaapt(wav, opts)` through the module attribute (hifigan.py:121): install(yaapt_too=True)
    rebinds it to the batched GPU extractor, uninstall() restores the TorchScript original.  Without a GPU the rebound
    function fails loudly instead of falling back to a CPU path."""
    import satools.hifigan.yaapt as ref_yaapt
    import satools_b200
    import importlib
    inst = importlib.import_module("satools_b200.install")      # the module (satools_b200.install is the function)
    original = ref_yaapt.yaapt
    satools_b200.install(yaapt_too=True)
    try:
        assert ref_yaapt.yaapt is inst.yaapt
        net = _build_net("f0")
        wav = torch.zeros(1, 8000)
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="CUDA"):
                net.get_f0(wav)
    finally:
        satools_b200.uninstall()
    assert ref_yaapt.yaapt is original
