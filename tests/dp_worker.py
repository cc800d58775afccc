"""Worker of tests/test_gpu_dp.py: one process per GPU under torch.distributed.run (NCCL over NVLink).

Checks, on real hardware, what tests/test_dp_gloo.py checks for the host logic on CPU:
  1. after 3 data-parallel steps the replicas are bit-identical (parameters, RMSprop state), in eager mode, with the
     step captured as ONE CUDA graph including the NCCL calls, and with per-segment graphs — and all three agree bit
     for bit with each other;
  2. the loss of a data-parallel step with sync_bn=True over the global batch equals the loss of ONE process training
     on the whole global batch (exact single-device BatchNorm statistics and ONE dice ratio over the global batch);
  3. the bucket schedule covers the whole flat gradient exactly once and leaves < 2 MB for after the backward.
Prints one line `DP_WORKER_OK {...}` on rank 0; any failed assertion exits non-zero.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _say(rank, msg):
    print(f"[dp_worker rank {rank}] {msg}", file=sys.stderr, flush=True)


def main():
    import faulthandler

    faulthandler.dump_traceback_later(150, exit=True)      # a hang prints every thread's stack and ends the process
    from jcfszxc_unet_b200.dp import DataParallel
    from jcfszxc_unet_b200.trainer import Trainer
    from UNetFamily.UNet import UNet

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per, size, lr = 2, 64, 1e-3
    g = torch.Generator().manual_seed(5)
    images = torch.rand(3, per * world, 3, size, size, generator=g)                 # 3 steps x global batch
    labels = (torch.rand(3, per * world, 1, size, size, generator=g) < 0.12).float()
    lo, hi = rank * per, (rank + 1) * per

    def run(graph, env=None, sync_bn=False, buckets=6):
        for k, v in (env or {}).items():
            os.environ[k] = v
        torch.manual_seed(42 + rank)            # replicas start from DIFFERENT weights: the broadcast must fix that
        m = UNet(3, 1).to(dev).train()
        tr = Trainer(m, lr=lr, use_cuda_graph=graph, dp=DataParallel(sync_bn=sync_bn), grad_buckets=buckets)
        losses = [float(tr.step(images[s, lo:hi].to(dev), labels[s, lo:hi].to(dev))) for s in range(3)]
        torch.cuda.synchronize()
        for k in (env or {}):
            os.environ.pop(k)
        made.append(tr)
        return tr, losses

    made = []

    def in_sync(t):
        mine = torch.cat([t.flat_p, t.sq, t.buf]).view(torch.int32)
        all_ = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(all_, mine)
        return all(torch.equal(a, all_[0]) for a in all_)

    report = {}
    _say(rank, "eager data-parallel run")
    tr_e, l_e = run(False)
    assert in_sync(tr_e), "eager data-parallel replicas diverged"
    # 3. bucket schedule
    cuts = tr_e._cuts
    covered = sorted(r for _, rs in cuts for r in rs)
    pos = 0
    for a, b in covered:
        assert a == pos, f"gap or overlap in the bucket schedule at {pos}: {(a, b)}"
        pos = b
    assert pos == tr_e.flat_g.numel() and len(cuts) >= 4, (pos, len(cuts))
    tail_mb = sum(b - a for a, b in cuts[-1][1]) * 4 / 2**20
    assert tail_mb < 2.0, f"the bucket reduced after the backward is {tail_mb:.2f} MB"
    report["buckets_mb"] = [round(sum(b - a for a, b in rs) * 4 / 2**20, 2) for _, rs in cuts]
    # 1. graph modes
    _say(rank, "single-graph run")
    tr_g, l_g = run(True)
    report["graph_mode"] = tr_g.graph_mode
    assert in_sync(tr_g), "graph-mode replicas diverged"
    assert l_g == l_e, (l_g, l_e)
    assert torch.equal(tr_g.flat_p, tr_e.flat_p) and torch.equal(tr_g.sq, tr_e.sq), "single-graph DP != eager DP"
    _say(rank, "per-segment graphs run")
    tr_s, l_s = run(True, env={"UNETK_DP_GRAPH": "0"})
    assert tr_s.graph_mode == "segments"
    assert l_s == l_e and torch.equal(tr_s.flat_p, tr_e.flat_p), "per-segment graphs != eager DP"
    _say(rank, "one-bucket run")
    tr_1, l_1 = run(True, buckets=1)
    # one bucket = nothing overlaps the backward = no SMs set aside (Trainer.sm_reserve): other split-K orders in the
    # weight gradients, so equal up to summation order only
    assert all(abs(a - b) <= 1e-4 * max(1.0, abs(b)) for a, b in zip(l_1, l_e)), (l_1, l_e)
    assert torch.allclose(tr_1.flat_p, tr_e.flat_p, rtol=0, atol=5e-3), "bucket count changed the result"
    tr_0, l_0 = run(True, env={"UNETK_DP_SM_RESERVE": "0"})
    assert in_sync(tr_0) and tr_0.sm_reserve == 0
    assert all(abs(a - b) <= 1e-4 * max(1.0, abs(b)) for a, b in zip(l_0, l_e)), (l_0, l_e)
    del tr_g, tr_s, tr_1, tr_0
    # 2. global batch, sync_bn: equal to one process on the whole batch
    _say(rank, "sync_bn run")
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(150, exit=True)
    tr_b, l_b = run(False, sync_bn=True)
    assert in_sync(tr_b)
    _say(rank, "single-process reference run")
    torch.manual_seed(42)                        # rank 0's initial weights are what the broadcast distributed
    m1 = UNet(3, 1).to(dev).train()
    t1 = Trainer(m1, lr=lr, use_cuda_graph=False, dp=DataParallel(enabled=False))
    made.append(t1)
    l_1p = [float(t1.step(images[s].to(dev), labels[s].to(dev))) for s in range(3)]
    report["loss_dp_syncbn"], report["loss_single_process"] = l_b, l_1p
    # same arithmetic, another summation order (per-rank partial sums): fp32/bf16 rounding differences only
    assert abs(l_b[0] - l_1p[0]) <= 2e-4 * max(1.0, abs(l_1p[0])), (l_b, l_1p)
    for a, b in zip(l_b, l_1p):
        assert abs(a - b) <= 2e-2 * max(1.0, abs(b)), (l_b, l_1p)
    # gradient norm of the first step is the global one on every rank
    _say(rank, "done")
    dist.barrier()
    if rank == 0:
        print("DP_WORKER_OK " + json.dumps(report), flush=True)
    # captured graphs hold NCCL kernels: release them before the communicator goes away (Trainer.close)
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(60, exit=True)
    for t in made:
        t.close()
    del tr_e, tr_b, t1, made
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()
    _say(rank, "process group destroyed")


if __name__ == "__main__":
    main()
