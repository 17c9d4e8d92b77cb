"""CPU tests: pin the oracles against the golden vectors minted from the reference itself
(oracle/make_golden.py), per op and end to end."""
import os

import numpy as np
import pytest
import torch

from oracle import hifigan_numpy as onp
from oracle import hifigan_torch_cpu as otc
from satools_b200 import conditioning
import helpers


@pytest.fixture(scope="module")
def kats():
    z = np.load(os.path.join(helpers.GOLDEN, "layer_kats.npz"))
    out = {}
    for k in z.files:
        name, leaf = k.split("/", 1)
        out.setdefault(name, {})[leaf] = z[k]
    return out


@pytest.mark.parametrize("name,dil", [("conv_k3_d1", 1), ("conv_k3_d5", 5), ("conv_k7_d3", 3),
                                      ("conv_k11_d5", 5), ("conv_k7_d1", 1)])
def test_conv1d_kat(kats, name, dil):
    c = kats[name]
    k = c["w"].shape[2]
    y = onp.conv1d(c["x"], c["w"], c["b"], dilation=dil, padding=onp.get_padding(k, dil))
    assert y.shape == c["y"].shape
    np.testing.assert_allclose(y, c["y"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,u", [("convt_k11_u5", 5), ("convt_k8_u4", 4), ("convt_k4_u2", 2)])
def test_conv_transpose_polyphase_kat(kats, name, u):
    c = kats[name]
    k = c["w"].shape[2]
    y = onp.conv_transpose1d(c["x"], c["w"], c["b"], stride=u, padding=(k - u) // 2)
    assert y.shape == c["y"].shape == (2, 6, 9 * u)
    np.testing.assert_allclose(y, c["y"], rtol=0, atol=1e-12)


def test_tail_reflect_pad_kat(kats):
    c = kats["tail"]
    h = onp.leaky_relu(c["x"], onp.FINAL_SLOPE)
    h = np.concatenate([h[:, :, 1:2], h], axis=2)
    y = np.tanh(onp.conv1d(h, c["w"], c["b"], padding=3))
    assert y.shape[-1] == c["x"].shape[-1] + 1
    np.testing.assert_allclose(y, c["y"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["wn_conv", "wn_convt"])
def test_weight_norm_fold_kat(kats, name):
    c = kats[name]
    np.testing.assert_allclose(onp.fold_weight_norm(c["g"], c["v"]), c["w"], rtol=1e-13, atol=0)


def test_resblock_kat(kats):
    c = kats["resblock_k7"]
    st = {k[2:]: v for k, v in c.items() if k.startswith("p_")}
    p = onp.folded_params(st)
    y = onp.resblock1(c["x"], [p[f"convs1.{m}"] for m in range(3)], [p[f"convs2.{m}"] for m in range(3)], 7)
    np.testing.assert_allclose(y, c["y"], rtol=0, atol=1e-12)


GEN_CASES = sorted(helpers.manifest()["generator"].items())


@pytest.mark.parametrize("name,meta", GEN_CASES)
def test_state_regenerates_bit_exact(name, meta):
    gen = helpers.seeded_generator(meta["seed"], meta.get("imput_dim", 504))
    assert len(gen.state_dict()) == 291
    assert helpers.state_sha256(gen.state_dict()) == meta["state_sha256"]


@pytest.mark.parametrize("name,meta", GEN_CASES)
def test_numpy_oracle_matches_reference(name, meta):
    g = np.load(os.path.join(helpers.GOLDEN, name + ".npz"))
    gen = helpers.seeded_generator(meta["seed"], meta.get("imput_dim", 504))
    x = helpers.case_input(name, meta)
    y, stages = onp.generator_forward(helpers.numpy_state(gen), x, return_stages=True)
    assert list(y.shape) == meta["y_shape"]
    assert helpers.max_abs(g["y_ref_fp64"], y) < 1e-12
    assert helpers.snr_db(g["y_ref_fp64"], y) > 200
    for i, s in enumerate(stages):
        np.testing.assert_allclose(helpers.stage_slices(s), g[f"stage{i}"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("name,meta", GEN_CASES)
def test_torch_cpu_port_matches_reference(name, meta):
    g = np.load(os.path.join(helpers.GOLDEN, name + ".npz"))
    gen = helpers.seeded_generator(meta["seed"], meta.get("imput_dim", 504))
    x = torch.from_numpy(helpers.case_input(name, meta))
    y = otc.generator_forward(otc.fold(gen.state_dict()), x).numpy()
    assert helpers.max_abs(g["y_ref_fp32"], y) < 2e-6      # fp32, fold order differs from the hook's
    assert helpers.snr_db(g["y_ref_fp64"], y) > 100


def test_chunked_equals_unchunked_oracle():
    """Receptive field of the generator is +-20 frames: windows with halo 20
    reproduce the full result."""
    from satools_b200 import scheduler
    gen = helpers.seeded_generator(0)
    st = helpers.numpy_state(gen)
    x = conditioning.batch(77, [70])
    full = onp.generator_forward(st, x)
    out = np.zeros_like(full)
    for rlo, rhi, klo, khi in scheduler.chunks(70, 25):
        y = onp.generator_forward(st, x[:, :, rlo:rhi])
        # output sample 1 + 320 f + s belongs to frame f; sample 0 is the reflect-pad extra
        lo, hi = 320 * (klo - rlo), 320 * (khi - rlo)
        if klo == 0:
            out[:, :, 0:1 + 320 * khi] = y[:, :, 0:1 + hi]
        else:
            out[:, :, 1 + 320 * klo:1 + 320 * khi] = y[:, :, 1 + lo:1 + hi]
    assert helpers.max_abs(full, out) < 1e-12


def test_reference_conditioning_layout():
    """Net._forward (hifigan.py:83-97) builds x as [BN 0..255 | F0 256 | speaker one-hot 257..]."""
    z = np.load(os.path.join(helpers.GOLDEN, "net_forward.npz"))
    for tag in ("plain", "quant_16_awgn_2"):
        x, bn, spk = z[f"{tag}/x"], z[f"{tag}/bn"], z[f"{tag}/spk"]
        assert x.shape == (2, 504, 16)
        np.testing.assert_array_equal(x[:, :256], bn)
        np.testing.assert_array_equal(x[:, 257:], np.repeat(spk[:, :, None].astype(np.float32), 16, axis=2))
        assert z[f"{tag}/y"].shape == (2, 1, 320 * 16 + 1)
    # quant+awgn changes only the F0 channel
    assert not np.array_equal(z["plain/x"][:, 256], z["quant_16_awgn_2/x"][:, 256])
    np.testing.assert_array_equal(z["plain/x"][:, :256], z["quant_16_awgn_2/x"][:, :256])


# ---- N3 last step: VectorQuantizerEMA assignment (oracle/vq_numpy.py against outputs of the reference's own module) ----
@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_vq_assign_oracle_matches_reference(case):
    from oracle import vq_numpy as ovq
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "vq_assign.npz"))
    x, cb = g[f"c{case}_inputs"], g[f"c{case}_codebook"]
    idx, q = ovq.assign(x, cb)
    np.testing.assert_array_equal(idx, g[f"c{case}_indices"])
    np.testing.assert_array_equal(q, g[f"c{case}_quantized"])
    assert ovq.margin(x, cb).min() > 1e-5            # no fixture row sits on an fp32 tie
