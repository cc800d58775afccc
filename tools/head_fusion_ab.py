import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_variants as T
from jcfszxc_unet_b200 import builders
from jcfszxc_unet_b200.trainer import Trainer
DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "AttentionUNet"
out = {}
for fuse in ("1", "0"):
    os.environ["UNETK_FUSE_HEAD"] = fuse
    m = T._make(name).to(DEV).train()
    tr = Trainer(m, lr=1e-3, use_cuda_graph=False, builder=getattr(builders, T.BUILDERS[name]))
    losses, grads = [], None
    for step in range(3):
        im, lb = T._inputs(100 + step, 2, 32, 32)
        losses.append(float(tr.step(im.to(DEV), lb.to(DEV))))
        if step == 0:
            grads = {k: tr.grad_views[id(p)].detach().clone() for k, p in m.named_parameters()}
            logits = tr.plan.head.logits.clone()
    out[fuse] = (losses, grads, logits)
    print("fuse", fuse, "prod", tr.plan.head.prod is not None, losses)
print("logits equal", torch.equal(out["1"][2], out["0"][2]))
worst = []
for k, g in out["0"][1].items():
    d = ((out["1"][1][k] - g).norm() / (g.norm() + 1e-20)).item()
    worst.append((d, k, g.norm().item()))
for d, k, n in sorted(worst, reverse=True)[:12]:
    print(f"{d:.3e} {k} |g|={n:.3e}")
