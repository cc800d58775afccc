"""Data-parallel host logic on 2 CPU ranks (gloo): jcfszxc_unet_b200/dp.py is device-agnostic, so the
sharding, the loss-sum all-reduce, the SUM gradient all-reduce, the shared clip coefficient and the
in-sync optimizer update are exercised here with the oracle standing in for the CUDA kernels.

Checked against ONE process computing the same global step (weights shared, BatchNorm statistics per
shard = the per-rank-BN semantics of the default DP mode):
  * the loss every rank reports == the single-process global loss (one BCE mean + one dice ratio over the
    WHOLE batch, utils/dice_score.py:13-38 with a 3-D input);
  * sum-reduced gradients == the single-process gradients;  * both ranks end with identical parameters.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_inputs():
    g = torch.Generator().manual_seed(2024)
    images = torch.rand(4, 3, 32, 32, generator=g)
    labels = (torch.rand(4, 1, 32, 32, generator=g) < 0.12).float()
    return images, labels


def _loss_sums(logits, labels):
    """The four sums the head kernel produces (loss.cu): sum BCE, sum p*y, sum p, sum y with p clamped."""
    bce = F.binary_cross_entropy_with_logits(logits, labels, reduction="sum")
    p = torch.clamp(torch.sigmoid(logits), 1e-7, 1 - 1e-7)
    return torch.stack([bce, (p * labels).sum(), p.sum(), labels.sum()]).double()


def _loss_from_sums(s, npix):
    """loss_finalize_kernel: 0.5 * mean BCE + 0.5 * (1 - dice), dice = (2I+eps)/(P+Y+eps), eps = 1e-5."""
    eps = 1e-5
    inter, sets = 2.0 * s[1], s[2] + s[3]
    sets = torch.where(sets < eps, inter, sets)
    return 0.5 * s[0] / npix + 0.5 * (1.0 - (inter + eps) / (sets + eps))


def _init_sd():
    from UNetFamily.UNet import UNet

    torch.manual_seed(42)
    return {k: v.detach().clone() for k, v in UNet(3, 1).state_dict().items()}


def _single_process_reference():
    """One process, two shards with shared weights and per-shard BN statistics, ONE global loss."""
    from oracle import unet_oracle as O

    sd = _init_sd()
    names = O.param_names(sd)
    for k in names:
        sd[k].requires_grad_(True)
    images, labels = _global_inputs()
    sums = 0
    for lo in (0, 2):
        logits = O.unet_forward(images[lo:lo + 2], sd, training=True)
        sums = sums + _loss_sums(logits, labels[lo:lo + 2])
    loss = _loss_from_sums(sums, labels.numel())
    loss.backward()
    grads = {k: sd[k].grad.clone() for k in names}
    flat = torch.cat([g.flatten() for g in grads.values()])
    (clipped,), total = O.clip_grad_norm([flat], 1.0)
    return float(loss), grads, float(total), clipped


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from jcfszxc_unet_b200.dp import DataParallel
    from oracle import unet_oracle as O

    dp = DataParallel(sync_bn=False, sync_loss=True)
    assert dp.enabled and dp.world == 2 and dp.rank == rank
    images, labels = _global_inputs()
    lo, hi = dp.shard(images.shape[0])
    assert (lo, hi) == (rank * 2, rank * 2 + 2)
    sd = _init_sd()
    flat_p = torch.cat([sd[k].flatten() for k in O.param_names(sd)])
    if rank == 1:
        flat_p += 1.0                      # replicas must start identical: rank 0's copy wins
    dp.broadcast_(flat_p, 0)
    names = O.param_names(sd)
    off = 0
    for k in names:
        n = sd[k].numel()
        sd[k] = flat_p[off:off + n].view(sd[k].shape).clone().requires_grad_(True)
        off += n
    # forward on the shard, local loss sums, all-reduce -> every rank holds the global sums
    logits = O.unet_forward(images[lo:hi], sd, training=True)
    local = _loss_sums(logits, labels[lo:hi])
    sums = local.detach().clone()
    npix = dp.reduce_loss_sums(sums, labels[lo:hi].numel())
    assert npix == labels.numel()
    # back-propagate the GLOBAL loss restricted to this rank's pixels: global = local + (others, constant)
    loss = _loss_from_sums(local + (sums - local.detach()), npix)
    loss.backward()
    flat_g = torch.cat([sd[k].grad.flatten() for k in names])
    # the Trainer's bucketed, overlapped reduction (decoder tail first, encoder head second) must equal ONE all-reduce
    bucketed = flat_g.clone()
    cut = bucketed.numel() // 2 + 3
    h_tail = dp.reduce_grads_async(bucketed[cut:])
    h_head = dp.reduce_grads_async(bucketed[:cut])
    dp.wait(h_tail)
    dp.wait(h_head)
    gscale = dp.reduce_grads(flat_g)       # SUM, not mean: the loss was already global
    assert gscale == 1.0
    assert torch.equal(bucketed, flat_g)
    (clipped,), total = O.clip_grad_norm([flat_g * gscale], 1.0)
    # optimizer on the flat buffers, identical on every rank
    p = flat_p.clone()
    sq, buf = torch.zeros_like(p), torch.zeros_like(p)
    O.rmsprop_step(p, clipped, sq, buf, 1e-3)
    # SyncBN plumbing: per-channel [2, C] sums are all-reduced and the count scales with the world
    dp_bn = DataParallel(sync_bn=True)
    bn_sums = torch.full((2, 8), float(rank + 1), dtype=torch.float64)
    count = dp_bn.reduce_bn_sums(bn_sums, 100)
    assert count == 200 and torch.all(bn_sums == 3.0)
    assert dp.max_over_ranks(float(rank), torch.device("cpu")) == 1.0
    torch.save({"loss": float(loss), "grad": flat_g, "total": float(total), "params": p}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_matches_single_process(tmp_path):
    from oracle import unet_oracle as O

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    ref_loss, ref_grads, ref_total, ref_clipped = _single_process_reference()
    # every rank reports the global loss
    assert abs(r0["loss"] - ref_loss) < 1e-6 and abs(r1["loss"] - ref_loss) < 1e-6
    # reduced gradients identical on both ranks and equal to the single-process gradient
    assert torch.equal(r0["grad"], r1["grad"])
    ref_flat = torch.cat([g.flatten() for g in ref_grads.values()])
    assert torch.allclose(r0["grad"], ref_flat, rtol=1e-4, atol=1e-7), (r0["grad"] - ref_flat).abs().max()
    assert abs(r0["total"] - ref_total) <= 1e-5 * ref_total
    # replicas stay in sync after the update (bit-identical), and the update is train.py:299-300 applied to the
    # reduced gradient (RMSprop's first step is ~sign(g)*10*lr, so it is compared on the SAME gradient bits)
    assert torch.equal(r0["params"], r1["params"])
    sd = _init_sd()
    p = torch.cat([sd[k].flatten() for k in O.param_names(sd)])
    sq, buf = torch.zeros_like(p), torch.zeros_like(p)
    (clipped,), _ = O.clip_grad_norm([r0["grad"]], 1.0)
    O.rmsprop_step(p, clipped, sq, buf, 1e-3)
    assert torch.equal(r0["params"], p)
    assert torch.allclose(clipped, ref_clipped, rtol=1e-4, atol=1e-7)


def test_shard_requires_divisible_batch():
    from jcfszxc_unet_b200.dp import DataParallel

    dp = DataParallel()
    assert not dp.enabled and dp.shard(16) == (0, 16)
    dp.world = 3
    with pytest.raises(ValueError):
        dp.shard(16)
