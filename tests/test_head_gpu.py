"""GPU parity of the fused projection-head tail: bn_contrastive_loss(u1, u2, bn) against
contrastive_loss(bn(u1), bn(u2)) -- reference models/simclr.py:38-39 feeding objective.py -- evaluated two ways: in fp64
on the CPU (torch batch_norm + the oracle's closed forms, chained through autograd) and with nn.BatchNorm1d + this
repository's own (separately verified) loss on the GPU.  Loss, accuracy, dL/du, dL/dgamma, dL/dbeta and the module's
running statistics."""
import copy

import numpy as np
import pytest
import torch

import contrastive_oracle as oracle
import pytorch_simclr_b200 as sb

pytestmark = pytest.mark.gpu
TOL = {"bf16": (2e-3, 1e-2), "fp32": (1e-5, 1e-4)}


@pytest.fixture(autouse=True)
def _restore():
    yield
    sb.set_precision("auto")


def _fp64_reference(kind, u1, u2, gamma, beta, eps, tau):
    """loss(BN(u1), BN(u2)) and its gradients in fp64: BatchNorm by torch autograd, the loss by the oracle."""
    a = u1.double().requires_grad_(True)
    c = u2.double().requires_grad_(True)
    g = gamma.double().requires_grad_(True)
    bt = beta.double().requires_grad_(True)
    z1 = torch.nn.functional.batch_norm(a, None, None, g, bt, True, 0.0, eps)
    z2 = torch.nn.functional.batch_norm(c, None, None, g, bt, True, 0.0, eps)
    fn = oracle.ntxent_closed_form if kind == "ntxent" else oracle.modified_closed_form
    ref = fn(z1.detach(), z2.detach(), temperature=tau)
    torch.autograd.backward([z1, z2], [torch.from_numpy(ref.grad1), torch.from_numpy(ref.grad2)])
    return ref, a.grad.numpy(), c.grad.numpy(), g.grad.numpy(), bt.grad.numpy()


@pytest.mark.parametrize("kind", ["ntxent", "modified"])
@pytest.mark.parametrize("b,d,precision", [(512, 128, "bf16"), (512, 128, "fp32"), (4096, 128, "bf16"), (300, 100, "bf16")])
def test_head_tail_training_matches_batchnorm_plus_loss(kind, b, d, precision):
    tau = 0.5
    gen = torch.Generator().manual_seed(b + d)
    base = torch.randn(b, d, generator=gen)
    u1 = 3.0 * (base + 0.7 * torch.randn(b, d, generator=gen)) + 1.5          # not centred, not unit variance
    u2 = 3.0 * (base + 0.7 * torch.randn(b, d, generator=gen)) - 0.5
    bn = torch.nn.BatchNorm1d(d).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(d, generator=gen) + 0.5)
        bn.bias.copy_(0.3 * torch.randn(d, generator=gen))
    bn_ref = copy.deepcopy(bn)
    fused = sb.bn_contrastive_loss if kind == "ntxent" else sb.bn_modified_contrastive_loss
    plain = sb.contrastive_loss if kind == "ntxent" else sb.modified_contrastive_loss
    sb.set_precision(precision)
    ltol, gtol = TOL[precision]

    a = u1.cuda().requires_grad_(True)
    c = u2.cuda().requires_grad_(True)
    loss, acc = fused(a, c, bn, temperature=tau)
    loss.backward()
    torch.cuda.synchronize()

    ref, g1, g2, gg, gb = _fp64_reference(kind, u1, u2, bn_ref.weight.detach().cpu(), bn_ref.bias.detach().cpu(), bn.eps, tau)
    assert float(loss) == pytest.approx(ref.loss, rel=ltol)
    assert acc == ref.acc
    scale = max(np.abs(g1).max(), np.abs(g2).max())
    assert np.abs(a.grad.cpu().numpy() - g1).max() < gtol * scale
    assert np.abs(c.grad.cpu().numpy() - g2).max() < gtol * scale
    assert np.abs(bn.weight.grad.cpu().numpy() - gg).max() < gtol * max(np.abs(gg).max(), 1e-12)
    assert np.abs(bn.bias.grad.cpu().numpy() - gb).max() < gtol * max(np.abs(gb).max(), 1e-12)

    # the unfused path on the GPU: nn.BatchNorm1d (two calls, as the training loop makes them) + the drop-in loss
    a2 = u1.cuda().requires_grad_(True)
    c2 = u2.cuda().requires_grad_(True)
    loss2, acc2 = plain(bn_ref(a2), bn_ref(c2), temperature=tau)
    loss2.backward()
    torch.cuda.synchronize()
    assert float(loss) == pytest.approx(float(loss2), rel=ltol)
    assert acc == acc2
    assert torch.allclose(bn.running_mean, bn_ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var, bn_ref.running_var, rtol=1e-4, atol=1e-6)
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked) == 2
    assert float((a.grad - a2.grad).abs().max()) < gtol * scale


def test_head_tail_eval_mode_and_no_grad():
    b, d, tau = 256, 128, 0.5
    gen = torch.Generator().manual_seed(7)
    u1 = torch.randn(b, d, generator=gen) * 2 + 1
    u2 = u1 + 0.5 * torch.randn(b, d, generator=gen)
    bn = torch.nn.BatchNorm1d(d).cuda()
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(d, generator=gen))
        bn.running_var.copy_(torch.rand(d, generator=gen) + 0.5)
        bn.weight.copy_(torch.rand(d, generator=gen) + 0.5)
    bn.eval()
    sb.set_precision("fp32")
    with torch.no_grad():
        loss, acc = sb.bn_contrastive_loss(u1.cuda(), u2.cuda(), bn, temperature=tau)
        loss_ref, acc_ref = sb.contrastive_loss(bn(u1.cuda()), bn(u2.cuda()), temperature=tau)
    assert float(loss) == pytest.approx(float(loss_ref), rel=1e-5) and acc == acc_ref
    # eval-mode gradients: dL/du = scale * dL/dz
    a = u1.cuda().requires_grad_(True)
    loss, _ = sb.bn_contrastive_loss(a, u2.cuda(), bn, temperature=tau)
    loss.backward()
    a2 = u1.cuda().requires_grad_(True)
    loss2, _ = sb.contrastive_loss(bn(a2), bn(u2.cuda()), temperature=tau)
    loss2.backward()
    torch.cuda.synchronize()
    assert float((a.grad - a2.grad).abs().max()) < 1e-4 * float(a2.grad.abs().max())
    assert int(bn.num_batches_tracked) == 0
