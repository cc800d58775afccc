"""CPU tests: the oracle replays the reference's own outputs (tests/golden/*.npz, produced by
oracle/pin_against_reference.py from the UNMODIFIED reference) and our drop-in modules reproduce the
reference's default initialisation and state_dict layout.  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _our_unet(seed=42):
    from UNetFamily.UNet import UNet

    torch.manual_seed(seed)
    return UNet(3, 1)


def test_state_dict_layout_matches_reference():
    m = _our_unet()
    sd = m.state_dict()
    assert len(sd) == 118                                   # SURVEY.md §8b [measured on the reference]
    assert sum(p.numel() for p in m.parameters()) == 31037633
    for k in ("inc.double_conv.0.weight", "inc.double_conv.1.running_mean", "inc.double_conv.4.num_batches_tracked",
              "down4.maxpool_conv.1.double_conv.3.weight", "up1.up.weight", "up1.up.bias",
              "up4.conv.double_conv.4.bias", "outc.conv.weight", "outc.conv.bias"):
        assert k in sd, k
    assert sd["up1.up.weight"].shape == (1024, 512, 2, 2)   # ConvTranspose2d layout [Cin, Cout, 2, 2]
    assert m.n_channels == 3 and m.n_classes == 1
    assert type(m).__module__ == "UNetFamily.UNet" and type(m).__qualname__ == "UNet"   # pickle path


def test_oracle_forward_replays_reference_with_our_default_init():
    """Same seed -> our modules draw the same weights as the reference -> oracle output == golden (bit-exact)."""
    g = _load("unet_forward_seed42.npz")
    m = _our_unet(42)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        y = O.unet_forward(_t(g["images"]), sd, training=True)
    assert np.array_equal(y.numpy(), g["logits_train"])
    assert np.array_equal(sd["inc.double_conv.1.running_mean"].numpy(), g["running_mean_inc1"])
    assert np.array_equal(sd["up4.conv.double_conv.4.running_var"].numpy(), g["running_var_up4_4"])
    assert int(sd["inc.double_conv.1.num_batches_tracked"]) == int(g["num_batches_tracked"]) == 1
    with torch.no_grad():
        y_eval = O.unet_forward(_t(g["images"]), sd, training=False)
    assert np.array_equal(y_eval.numpy(), g["logits_eval_after_1_train_fwd"])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_oracle_train_step_replays_reference(mode):
    g = _load(f"unet_trainstep_seed42_{mode}.npz")
    m = _our_unet(42)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    names = O.param_names(sd)
    assert list(g["param_names"]) == names
    opt_state = {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in names}
    for step in range(2):
        gen = torch.Generator().manual_seed(100 + step)
        images = torch.rand(2, 3, 32, 32, generator=gen)
        labels = (torch.rand(2, 1, 32, 32, generator=gen) < 0.12).float()
        loss, logits, grads = O.train_step(sd, opt_state, images, labels, 1e-3, bf16=(mode == "bf16"))
        assert np.array_equal(loss.float().numpy(), g[f"loss{step}"])
        assert np.array_equal(logits.float().numpy(), g[f"logits{step}"])
        norms = np.array([float(grads[k].float().norm()) for k in names])
        assert np.array_equal(norms, g[f"gradnorm{step}"])
        for key in g.files:
            if key.startswith(f"grad{step}:"):
                assert np.array_equal(grads[key.split(":", 1)[1]].float().numpy(), g[key]), key
            if key.startswith(f"param{step}:"):
                assert np.array_equal(sd[key.split(":", 1)[1]].float().numpy(), g[key]), key


def test_oracle_blocks_replay_reference():
    from UNetFamily.utils.unet_parts import DoubleConv, Down, OutConv, Up

    g = _load("blocks_seeds3to6.npz")
    x = _t(g["dc_x"])
    torch.manual_seed(3)
    dc = DoubleConv(8, 16)
    with torch.no_grad():
        assert np.array_equal(O.double_conv(x, {k: v.clone() for k, v in dc.state_dict().items()}, "", True).numpy(), g["dc_y"])
    torch.manual_seed(4)
    dn = Down(8, 16)
    with torch.no_grad():
        assert np.array_equal(O.down(x, {k: v.clone() for k, v in dn.state_dict().items()}, "", True).numpy(), g["down_y"])
    torch.manual_seed(5)
    up = Up(16, 8)
    with torch.no_grad():
        y = O.up(_t(g["up_x1"]), _t(g["up_x2"]), {k: v.clone() for k, v in up.state_dict().items()}, "", True)
    assert np.array_equal(y.numpy(), g["up_y"])
    torch.manual_seed(6)
    oc = OutConv(8, 1)
    with torch.no_grad():
        assert np.array_equal(O.out_conv(x, dict(oc.state_dict()), "").numpy(), g["outc_y"])


def test_oracle_dice_cases():
    g = _load("dice_cases.npz")
    for name in ("rand", "empty", "full", "out_of_range"):
        p, t = _t(g[f"{name}_p"]), _t(g[f"{name}_t"])
        assert np.array_equal(O.dice_coeff(p, t, reduce_batch_first=True).numpy(), g[f"{name}_coeff_batch"])
        assert np.array_equal(O.dice_coeff(p[:, None], t[:, None]).numpy(), g[f"{name}_coeff_per_item"])
        assert np.array_equal(O.dice_loss(p, t).numpy(), g[f"{name}_loss"])
        # dice_coeff_numpy restates dice_coeff (which clamps to [0,1]); dice_loss additionally clamps to [1e-7, 1-1e-7]
        assert abs(O.dice_coeff_numpy(g[f"{name}_p"], g[f"{name}_t"]) - float(g[f"{name}_coeff_batch"])) < 1e-6
    assert float(g["empty_coeff_batch"]) == 1.0            # empty-mask branch (dice_score.py:35)


def test_oracle_maxpool_indices():
    g = _load("maxpool_indices.npz")
    vals, idx = O.maxpool2x2_with_indices_numpy(g["x"])
    assert np.array_equal(idx, g["indices"])
    assert np.array_equal(np.nan_to_num(vals, nan=-7.0), np.nan_to_num(g["values"], nan=-7.0))
    # the documented rule: all-zero 4x4 plane -> first element of every window
    z = np.zeros((1, 1, 4, 4), dtype=np.float32)
    assert O.maxpool2x2_with_indices_numpy(z)[1].ravel().tolist() == [0, 2, 8, 10]


def test_product_path_has_no_cpu_fallback():
    m = _our_unet()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))


def test_max_unpool_restatement_matches_torch():
    """SegNet's pool / unpool pair (SegNet.py:89-138): the numpy restatement of F.max_unpool2d against torch itself on
    CPU, on indices produced by F.max_pool2d(kernel 2, stride 2) including ties and NaN."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 5, 8, 12, generator=g)
    x[0, 0, 0, :4] = 1.0                       # ties
    x[1, 2, 3, 5] = float("nan")
    p, idx = F.max_pool2d(x, 2, 2, return_indices=True)
    z = torch.randn(p.shape, generator=g)
    ref = F.max_unpool2d(z, idx, 2, 2)
    got = O.max_unpool2x2_numpy(z.numpy(), idx.numpy())
    assert np.array_equal(got, ref.numpy())
    pv, pidx = O.maxpool2x2_with_indices_numpy(x.numpy())
    assert np.array_equal(pidx, idx.numpy())
    assert np.array_equal(O.max_unpool2x2_numpy(pv, pidx), F.max_unpool2d(p, idx, 2, 2).numpy(), equal_nan=True)
