"""Training-batch assembly: host logic on the CPU, the gather kernel on the GPU, both against
the oracle's restatement of train.py:126-155, 200-253."""
import numpy as np
import pytest
import torch


def _pools(seed, n=3, h=96, w=80, c=3):
    r = np.random.RandomState(seed)
    images = r.rand(n, h, w, c).astype(np.float32)
    masks = (r.rand(n, h, w) < 0.05).astype(np.uint8)
    masks[0, 0, 0] = 1          # corner and border pixels: filtered out
    masks[1, h - 1, w - 1] = 1
    labels = (r.rand(n, h, w) < 0.12).astype(np.float32)
    return images, masks, labels


def test_sample_map_matches_oracle():
    from jcfszxc_unet_b200.sampler import filtered_sample_map
    from oracle import unet_oracle as O

    _, masks, _ = _pools(1)
    for p in (16, 48, 64):
        ours, ref = filtered_sample_map(masks, p), O.patch_sample_map(masks, p)
        assert all(np.array_equal(a, b) for a, b in zip(ours, ref))
        half = p // 2
        assert (ours[1] >= half).all() and (ours[1] < masks.shape[1] - half).all()
        assert (ours[2] >= half).all() and (ours[2] < masks.shape[2] - half).all()


@pytest.mark.gpu
@pytest.mark.parametrize("patch,batch", [(16, 5), (48, 16), (64, 3)])
def test_gather_bit_exact_vs_reference_loop(patch, batch):
    from jcfszxc_unet_b200.sampler import PatchSampler
    from oracle import unet_oracle as O

    images, masks, labels = _pools(7)
    s = PatchSampler(images, masks, labels, patch, device="cuda:0")
    got_i, got_l = s.sample(batch, rng=np.random.RandomState(123))
    ref_i, ref_l = O.patch_batch(images, labels, O.patch_sample_map(masks, patch), batch, patch, np.random.RandomState(123))
    assert got_i.shape == ref_i.shape and got_i.stride() == ref_i.stride()      # channels_last, as train.py:248-252
    assert got_l.shape == ref_l.shape
    assert torch.equal(got_i.cpu(), ref_i) and torch.equal(got_l.cpu(), ref_l)


@pytest.mark.gpu
def test_gather_rejects_bad_centres_and_feeds_the_model():
    from jcfszxc_unet_b200.sampler import PatchSampler
    from UNetFamily.UNet import UNet

    images, masks, labels = _pools(9)
    s = PatchSampler(images, masks, labels, 32, device="cuda:0")
    with pytest.raises(ValueError):
        s.gather(np.array([[0, 5, 40]], dtype=np.int32))        # x too close to the border
    with pytest.raises(ValueError):
        s.gather(np.array([[7, 40, 40]], dtype=np.int32))       # no such image
    x, y = s.sample(2, rng=np.random.RandomState(0))
    torch.manual_seed(0)
    m = UNet(3, 1).to("cuda:0").train()
    with torch.no_grad():
        out = m(x)                                               # the channels_last batch is consumed as is
    assert out.shape == (2, 1, 32, 32) and torch.isfinite(out).all() and y.shape == (2, 1, 32, 32)
