"""CPU tests for the U-Net variants: the oracle replays the reference's own outputs (tests/golden/<model>_seed42.npz
and variant_blocks_seeds11to17.npz, produced by oracle/pin_against_reference.py from the UNMODIFIED reference) and
our drop-in modules reproduce the reference's state_dict layout and default initialisation draw for draw."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

VARIANTS = {
    # name -> (module, class, #state_dict entries and #parameters measured on the reference)
    "AttentionUNet": ("UNetFamily.AttentionUNet", "AttentionUNet", 240, 34878573),
    "R2UNet": ("UNetFamily.R2UNet", "R2UNet", 174, 39091393),
    "R2AttentionUNet": ("UNetFamily.R2AttentionUNet", "R2AttentionUNet", 258, 39442925),
    "ResUNet": ("UNetFamily.ResUNet", "ResUNet", 145, 13043009),
    "NestedUNet": ("UNetFamily.UNetPP", "NestedUNet", 212, 9163329),
}


def _make(name, seed=42):
    import importlib

    mod, cls, _, _ = VARIANTS[name]
    torch.manual_seed(seed)
    return getattr(importlib.import_module(mod), cls)()


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", list(VARIANTS))
def test_state_dict_layout_and_init_match_reference(name):
    g = np.load(os.path.join(GOLDEN, f"{name.lower()}_seed42.npz"), allow_pickle=False)
    m = _make(name)
    sd = m.state_dict()
    _, cls, n_keys, n_params = VARIANTS[name]
    assert list(sd.keys()) == [str(k) for k in g["state_dict_keys"]]          # same names, same order
    assert len(sd) == n_keys and sum(p.numel() for p in m.parameters()) == n_params
    ours = np.array([float(v.double().abs().sum()) for v in sd.values()])
    assert np.array_equal(ours, g["init_abs_sum"])                             # same RNG draws as the reference
    assert m.n_channels == 3 and m.n_classes == 1
    assert type(m).__module__ == VARIANTS[name][0] and type(m).__qualname__ == cls   # pickle path (train.py:374)


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_replays_reference_forward(name):
    g = np.load(os.path.join(GOLDEN, f"{name.lower()}_seed42.npz"), allow_pickle=False)
    sd = {k: v.detach().clone() for k, v in _make(name).state_dict().items()}
    fwd = O.FORWARDS[name]
    with torch.no_grad():
        y = fwd(_t(g["images"]), sd, training=True)
        y_eval = fwd(_t(g["images"]), sd, training=False)
    assert np.array_equal(y.numpy(), g["logits_train"])
    assert np.array_equal(y_eval.numpy(), g["logits_eval_after_1_train_fwd"])
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        sd2 = {k: v.detach().clone() for k, v in _make(name).state_dict().items()}
        y16 = fwd(_t(g["images"]), sd2, training=True).float()
    assert np.array_equal(y16.numpy(), g["logits_train_bf16_autocast"])


@pytest.mark.parametrize("name", ["AttentionUNet", "ResUNet", "NestedUNet"])
def test_oracle_replays_reference_train_step(name):
    g = np.load(os.path.join(GOLDEN, f"{name.lower()}_seed42.npz"), allow_pickle=False)
    sd = {k: v.detach().clone() for k, v in _make(name).state_dict().items()}
    names = O.param_names(sd)
    assert names == [str(k) for k in g["step_param_names"]]
    opt_state = {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in names}
    loss, logits, grads = O.train_step(sd, opt_state, _t(g["step_images"]), _t(g["step_labels"]), 1e-3, bf16=False,
                                       model=name)
    assert np.array_equal(loss.numpy(), g["step_loss"])
    assert np.array_equal(logits.numpy(), g["step_logits"])
    assert np.allclose([float(grads[k].norm()) for k in names], g["step_gradnorm"], rtol=0, atol=0)


def test_oracle_replays_variant_blocks():
    g = np.load(os.path.join(GOLDEN, "variant_blocks_seeds11to17.npz"), allow_pickle=False)
    from UNetFamily.utils import unet_parts as parts

    x = _t(g["x16"])

    def sd_of(m):
        return {k: v.detach().clone() for k, v in m.state_dict().items()}

    with torch.no_grad():
        torch.manual_seed(11)
        assert np.array_equal(O.conv_block(x, sd_of(parts.conv_block(16, 24)), "", True).numpy(), g["conv_block_y"])
        torch.manual_seed(12)
        assert np.array_equal(O.up_conv(x, sd_of(parts.up_conv(16, 8)), "", True).numpy(), g["up_conv_y"])
        torch.manual_seed(13)
        assert np.array_equal(O.recurrent_block(x, sd_of(parts.Recurrent_block(16, t=2)), "", True, 2).numpy(), g["recurrent_y"])
        torch.manual_seed(14)
        assert np.array_equal(O.rrcnn_block(x, sd_of(parts.RRCNN_block(16, 24, t=2)), "", True, 2).numpy(), g["rrcnn_y"])
        torch.manual_seed(15)
        assert np.array_equal(O.attention_block(_t(g["att_g"]), x, sd_of(parts.Attention_block(16, 16, 8)), "", True).numpy(),
                              g["attention_y"])
        for stride, seed in ((1, 16), (2, 17)):
            torch.manual_seed(seed)
            y = O.residual_conv(x, sd_of(parts.ResidualConv(16, 24, stride, 1)), "", True, stride)
            assert np.array_equal(y.numpy(), g[f"residual_conv_s{stride}_y"])


def test_variants_refuse_cpu_tensors():
    m = _make("ResUNet")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 16, 16))
