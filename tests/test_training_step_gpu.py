"""The loss swapped into a SimCLR-style training step (BASELINE.json configs[1], scaled down): encoder + projection
head stay PyTorch modules, `from objective import contrastive_loss` resolves to this repository (the drop-in of
reference utils/model_utils.py:2,115-123).  Parity of the loss value, of the parameter gradients and of the weights after
a few Adam steps against the same step with the reference arithmetic restated in torch on the GPU."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def reference_contrastive_loss(x1, x2, temperature=1.0):
    """objective.py:23-53 restated with torch ops (test oracle for the training step; dense 2B x 2B logits)."""
    b = x1.shape[0]
    x1, x2 = F.normalize(x1, p=2, dim=1), F.normalize(x2, p=2, dim=1)           # :25-30
    eye = torch.eye(b, device=x1.device)
    aa = x1 @ x1.t() / temperature - eye * 1e9                                  # :35,39
    bb = x2 @ x2.t() / temperature - eye * 1e9                                  # :36,40
    ab = x1 @ x2.t() / temperature                                              # :42
    ba = x2 @ x1.t() / temperature                                              # :43
    logits = torch.cat((torch.cat((ab, aa), 1), torch.cat((bb, ba), 1)), 0)     # :48
    labels = torch.arange(2 * b, device=x1.device)                              # :49
    loss = F.cross_entropy(logits, labels)                                      # :47,50
    acc = 100.0 * (logits.argmax(1) == labels).float().mean().item()            # :51-53
    return loss, acc


class TinySimCLR(nn.Module):
    """Stand-in for models/simclr.py: conv encoder -> g(): Linear-BN-ReLU-Linear-BN (models/simclr.py:33-46)."""

    def __init__(self, dim=128):
        super().__init__()
        self.f = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.AdaptiveAvgPool2d(4),
                               nn.Flatten(), nn.Linear(512, 256), nn.ReLU())
        self.g = nn.Sequential(nn.Linear(256, 256, bias=False), nn.BatchNorm1d(256), nn.ReLU(),
                               nn.Linear(256, dim, bias=False), nn.BatchNorm1d(dim))

    def forward(self, x):
        return self.g(self.f(x))


def _train(model, loss_fn, batches, accum_steps, tau):
    # the reference uses Adam (utils/model_utils.py:96); Adam normalises every gradient element to a unit-size step, so
    # elements whose true gradient is zero (anything in front of a BatchNorm) would random-walk on rounding noise and
    # make a weight comparison meaningless: SGD keeps the trajectory comparison about the loss
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-6)
    losses, accs = [], []
    grads = None
    for i, (xa, xb) in enumerate(batches):
        z1, z2 = model(xa), model(xb)                                            # :113-114
        loss, acc = loss_fn(z1, z2, temperature=tau)                             # :115
        loss /= accum_steps                                                      # :116 (in place on the returned tensor)
        losses.append(loss.item())                                               # :117
        accs.append(acc)
        loss.backward()                                                          # :120
        if (i + 1) % accum_steps == 0:
            if grads is None:
                grads = [p.grad.detach().clone() for p in model.parameters()]
            opt.step()                                                           # :122
            opt.zero_grad()
    return losses, accs, grads


@pytest.mark.parametrize("accum_steps", [1, 2])
def test_loss_swapped_into_a_training_step(accum_steps):
    from objective import contrastive_loss          # the drop-in module at the repository root
    import pytorch_simclr_b200 as sb
    sb.set_precision("auto")                         # float32 embeddings -> fp32-grade arithmetic
    torch.manual_seed(0)
    tau, batch = 0.5, 256
    model_a = TinySimCLR().cuda()
    model_b = copy.deepcopy(model_a)
    gen = torch.Generator().manual_seed(1)
    batches = []
    for _ in range(4):
        base = torch.randn(batch, 3, 16, 16, generator=gen)
        batches.append(((base + 0.1 * torch.randn(batch, 3, 16, 16, generator=gen)).cuda(),
                        (base + 0.1 * torch.randn(batch, 3, 16, 16, generator=gen)).cuda()))
    la, aa, ga = _train(model_a, contrastive_loss, batches, accum_steps, tau)
    lb, ab, gb = _train(model_b, reference_contrastive_loss, batches, accum_steps, tau)
    # first step: identical weights -> the loss is the same number and the parameter gradients agree
    assert la[0] == pytest.approx(lb[0], rel=1e-5)
    assert aa[0] == pytest.approx(ab[0], abs=100.0 / (2 * batch) + 1e-6)
    # (parameters in front of a BatchNorm have a mathematically zero gradient: compare on the global gradient scale)
    scale = max(float(y.abs().max()) for y in gb)
    for x, y in zip(ga, gb):
        assert float((x - y).abs().max()) / scale < 1e-3
    # after four optimiser steps the two runs are still on the same trajectory
    assert la[-1] == pytest.approx(lb[-1], rel=2e-3)
    for pa, pb in zip(model_a.parameters(), model_b.parameters()):
        assert torch.allclose(pa, pb, rtol=0, atol=2e-3 * (float(pb.abs().max()) + 1e-6) + 1e-5)
