"""n_classes > 1: the K-channel OutConv head through the nn.Module surface, and the Dice
coefficient / loss of utils/dice_score.py (dice_coeff, multiclass_dice_coeff, dice_loss) as CUDA reductions, all
against the oracle (oracle/unet_oracle.py, pinned to the reference) on identical inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


@pytest.mark.parametrize("k", [2, 3, 5])
def test_unet_multiclass_forward_backward_vs_oracle(k):
    """UNet(3, n_classes=k): logits and CrossEntropy gradients (train.py:124 picks nn.CrossEntropyLoss for k > 1)."""
    from oracle import unet_oracle as O
    from UNetFamily.UNet import UNet

    torch.manual_seed(42)
    m = UNet(3, k).to(DEV).train()
    assert m.n_classes == k
    sd = {n: v.detach().clone() for n, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 64, 48, generator=g).to(DEV)
    y = torch.randint(0, k, (2, 64, 48), generator=g).to(DEV)
    logits = m(x)
    assert logits.shape == (2, k, 64, 48) and logits.dtype == torch.float32
    loss = F.cross_entropy(logits, y)
    loss.backward()

    def oracle_run(bf16):
        s = {n: v.clone() for n, v in sd.items()}
        names = O.param_names(s)
        for n in names:
            s[n].requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            lg = O.unet_forward(x, s, training=True)
        ls = F.cross_entropy(lg.float(), y)
        gr = torch.autograd.grad(ls, [s[n] for n in names])
        return lg.float().detach(), dict(zip(names, gr))

    lg32, g32 = oracle_run(False)
    lg16, g16 = oracle_run(True)
    # same acceptance rule as tests/test_gpu_unet.py: BASELINE tolerance (2e-2) or as close to fp32 as stock bf16 autocast
    l_err, l_ref = _rel(logits, lg32), _rel(lg16, lg32)
    print(f"k={k}: logits ours {l_err:.3g} | stock bf16 {l_ref:.3g}")
    assert l_err <= max(2e-2, 1.25 * l_ref)
    for name in ("outc.conv.weight", "outc.conv.bias", "up4.conv.double_conv.3.weight", "inc.double_conv.0.weight"):
        ours = dict(m.named_parameters())[name].grad
        e, e_ref = _rel(ours, g32[name]), _rel(g16[name], g32[name])
        print(f"  d{name}: ours {e:.3g} | stock bf16 {e_ref:.3g}")
        assert e <= max(3e-2, 2.0 * e_ref), name


def test_head_multi_kernels_exact():
    """The K-class head alone against fp32 torch on the same bf16 activations: forward, dx, dw, db (+ accumulate)."""
    from jcfszxc_unet_b200 import _lib

    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(8)
    for (n, h, w, c, k) in [(2, 9, 7, 64, 3), (1, 16, 16, 32, 8), (3, 5, 5, 128, 2)]:
        x = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
        wt = torch.randn(k, c, device=DEV, generator=g) / c ** 0.5
        b = torch.randn(k, device=DEV, generator=g)
        logits = torch.empty(n, k, h, w, device=DEV)
        _lib.call("unetk_head_multi_fwd", x.data_ptr(), c, wt.data_ptr(), b.data_ptr(), logits.data_ptr(), n, h * w, c, k, s)
        ref = torch.einsum("nhwc,kc->nkhw", x.float(), wt) + b.view(1, k, 1, 1)
        assert torch.allclose(logits, ref, rtol=1e-5, atol=1e-5)
        dl = torch.randn(n, k, h, w, device=DEV, generator=g)
        dx = torch.empty_like(x)
        dw = torch.full((k, c), 2.0, device=DEV)
        db = torch.full((k,), -1.0, device=DEV)
        partial = torch.empty(lib.unetk_head_multi_partial_floats(n * h * w, c, k), device=DEV)
        _lib.call("unetk_head_multi_bwd", x.data_ptr(), c, wt.data_ptr(), dl.data_ptr(), 0.5, dx.data_ptr(), c,
                  dw.data_ptr(), db.data_ptr(), 1, n, h * w, c, k, partial.data_ptr(), s)
        torch.cuda.synchronize()
        dxr = 0.5 * torch.einsum("nkhw,kc->nhwc", dl, wt)
        assert _rel(dx, dxr) <= 6e-3      # one bf16 rounding
        assert torch.allclose(dw, 2.0 + 0.5 * torch.einsum("nkhw,nhwc->kc", dl, x.float()), rtol=1e-4, atol=1e-4)
        assert torch.allclose(db, -1.0 + 0.5 * dl.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-4)


def test_trainer_rejects_multiclass():
    from jcfszxc_unet_b200.trainer import Trainer
    from UNetFamily.UNet import UNet

    m = UNet(3, 2).to(DEV).train()
    tr = Trainer(m, use_cuda_graph=False)
    with pytest.raises(NotImplementedError):
        tr.step(torch.rand(1, 3, 32, 32, device=DEV), torch.zeros(1, 1, 32, 32, device=DEV))


@pytest.mark.parametrize("shape,rbf", [((4, 40, 36), True), ((4, 40, 36), False), ((33, 29), False), ((2, 3, 24, 20), False)])
def test_dice_coeff_vs_oracle(shape, rbf):
    from oracle import unet_oracle as O
    from utils.dice_score import dice_coeff

    g = torch.Generator().manual_seed(sum(shape))
    p = (torch.rand(*shape, generator=g) * 1.4 - 0.2)       # values outside [0, 1] exercise the clamp
    t = (torch.rand(*shape, generator=g) < 0.3).float()
    if len(shape) == 3 and not rbf:
        t[1] = 0; p[1] = 0                                     # one empty sample: the sets_sum := inter branch
    ours = dice_coeff(p.to(DEV), t.to(DEV), reduce_batch_first=rbf)
    ref = O.dice_coeff(p, t, reduce_batch_first=rbf)
    assert abs(float(ours) - float(ref)) <= 1e-6, (float(ours), float(ref))


def test_multiclass_dice_and_loss_vs_oracle_with_gradient():
    from oracle import unet_oracle as O
    from utils.dice_score import dice_loss, multiclass_dice_coeff

    g = torch.Generator().manual_seed(21)
    logits = torch.randn(3, 4, 32, 28, generator=g) * 3
    labels = torch.randint(0, 4, (3, 32, 28), generator=g)
    onehot = F.one_hot(labels, 4).permute(0, 3, 1, 2).float()
    probs = F.softmax(logits, dim=1)
    ours = multiclass_dice_coeff(probs.to(DEV), onehot.to(DEV), reduce_batch_first=True)
    ref = O.multiclass_dice_coeff(probs, onehot, reduce_batch_first=True)
    assert abs(float(ours) - float(ref)) <= 1e-6
    ours_nb = multiclass_dice_coeff(probs.to(DEV), onehot.to(DEV))
    assert abs(float(ours_nb) - float(O.multiclass_dice_coeff(probs, onehot))) <= 1e-6
    # dice_loss, both modes, with the gradient w.r.t. the prediction (saturated probabilities hit the 1e-7 clamp)
    for multiclass, (pp, tt) in ((True, (probs, onehot)), (False, (torch.sigmoid(logits[:, 0] * 8), onehot[:, 1]))):
        a = pp.clone().to(DEV).requires_grad_(True)
        b = pp.clone().requires_grad_(True)
        lo = dice_loss(a, tt.to(DEV), multiclass=multiclass)
        lr = O.dice_loss(b, tt, multiclass=multiclass)
        assert abs(float(lo) - float(lr)) <= 1e-6
        lo.backward()
        lr.backward()
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-9), (a.grad.cpu() - b.grad).abs().max()
    # the golden dice cases of the reference itself (tests/golden/dice_cases.npz) through the CUDA op
    import os

    import numpy as np

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "dice_cases.npz"))
    from utils.dice_score import dice_coeff

    for case in sorted({k.rsplit("_", 1)[0] for k in gold.files if k.endswith("_p")}):
        p = torch.from_numpy(gold[case + "_p"]).to(DEV)
        t = torch.from_numpy(gold[case + "_t"]).to(DEV)
        assert abs(float(dice_coeff(p, t, reduce_batch_first=True)) - float(gold[case + "_coeff_batch"])) <= 1e-6, case
        assert abs(float(dice_coeff(p, t, reduce_batch_first=False)) - float(gold[case + "_coeff_per_item"])) <= 1e-6, case
        assert abs(float(dice_loss(p, t)) - float(gold[case + "_loss"])) <= 1e-6, case
