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


# ------------------------------------------------------------------------------------------------ deep supervision
def test_deep_supervision_module_and_oracle_match_reference():
    """BASELINE.json configs[4] names UNet++ "deep supervision" (UNetPP.py:65-69, 93-102; hard-coded off in the
    reference).  Golden vectors come from the reference class with the attribute forced on: our module must create the
    same parameters with the same draws, and the oracle must replay outputs, the mean-of-heads loss and gradients."""
    from UNetFamily.UNetPP import NestedUNet

    g = np.load(os.path.join(GOLDEN, "nestedunet_ds_seed42.npz"), allow_pickle=False)
    torch.manual_seed(42)
    m = NestedUNet(3, 1, deepsupervision=True)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert list(sd.keys()) == [str(k) for k in g["keys"]] and "final4.bias" in sd and "final.weight" not in sd
    assert np.array_equal(sd["final1.weight"].numpy(), g["final1_weight"])            # same RNG draws
    assert NestedUNet().deepsupervision is False and "final.weight" in NestedUNet().state_dict()   # the default is the reference's
    images, labels = _t(g["images"]), _t(g["labels"])
    with torch.no_grad():
        y = O.FORWARDS["NestedUNetDS"](images, {k: v.clone() for k, v in sd.items()}, True)
    assert all(np.array_equal(y[k].numpy(), g[f"out{k + 1}_train"]) for k in range(4))
    s = {k: v.clone() for k, v in sd.items()}
    for k in O.param_names(s):
        s[k].requires_grad_(True)
    _, loss, _, _ = O.forward_loss(s, images, labels, bf16=False, training=True, model="NestedUNetDS")
    loss.backward()
    assert float(loss) == float(g["loss"])
    assert np.array_equal(s["final1.weight"].grad.numpy(), g["grad_final1_weight"])
    assert np.array_equal(s["conv0_0.conv.0.weight"].grad.numpy(), g["grad_conv0_0_w"])
