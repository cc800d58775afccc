"""Sliding-window inference with overlap averaging against the oracle's restatement of
evaluate.py:28-96."""
import numpy as np
import pytest
import torch


def test_patch_grid_matches_reference_formula():
    from jcfszxc_unet_b200.inference import patch_positions

    assert patch_positions(100, 90, 64, 0.5) == [(0, 0), (32, 0)]
    assert patch_positions(256, 256, 256, 0.5) == [(0, 0)]
    assert patch_positions(100, 100, 128, 0.5) == []                      # image smaller than a patch: nothing to do
    assert len(patch_positions(512, 512, 256, 0.75)) == 25               # step 64: 5 x 5


class _Linear(torch.nn.Module):
    """A stand-in 'model' whose logits are an exact function of the patch: isolates the tiling arithmetic."""

    def forward(self, x):
        return (x * torch.tensor([0.5, -1.0, 2.0], device=x.device).view(1, 3, 1, 1)).sum(1, keepdim=True) - 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,p,ov,bs", [(96, 80, 32, 0.5, 4), (100, 90, 64, 0.5, 3), (70, 130, 32, 0.75, 16), (40, 40, 64, 0.5, 4)])
def test_tiling_arithmetic_vs_reference(h, w, p, ov, bs):
    from jcfszxc_unet_b200.inference import predict_full_image
    from oracle import unet_oracle as O

    img = np.random.RandomState(h + w).rand(h, w, 3).astype(np.float32)
    m = _Linear()
    got = predict_full_image(m, "cuda:0", img, p, ov, bs)
    ref = O.predict_full_image(lambda b: m(b), img, p, ov, bs)
    assert got.shape == ref.shape == (1, h, w) and got.dtype == np.float64
    assert np.abs(got - ref).max() <= 2e-7, np.abs(got - ref).max()       # fp32 sigmoid: 1-2 ulp between expf and ATen
    # pixels no window reaches stay zero (the reference's `where=count != 0`)
    assert (got[0, -1, -1] == 0.0) == (ref[0, -1, -1] == 0.0)


@pytest.mark.gpu
def test_full_pipeline_with_unet_vs_oracle_forward():
    from jcfszxc_unet_b200.inference import predict_full_image
    from oracle import unet_oracle as O
    from UNetFamily.UNet import UNet

    torch.manual_seed(42)
    m = UNet(3, 1)
    g = torch.Generator().manual_seed(3)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(0.6 + 0.8 * torch.rand(mod.num_features, generator=g))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to("cuda:0")
    img = np.random.RandomState(5).rand(96, 112, 3).astype(np.float32)
    got = predict_full_image(m, "cuda:0", img, 64, 0.5, 4)
    ref = O.predict_full_image(lambda b: O.unet_forward(b, sd, training=False), img, 64, 0.5, 4)
    err = np.abs(got - ref).max()
    print("sliding-window probabilities: max abs err", err)
    assert err <= 2e-2        # bf16 forward against the fp32 oracle, on probabilities in [0, 1]
