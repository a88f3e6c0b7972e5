"""SimCLR pretrain step with the loss swapped in (BASELINE.json configs[1] and the first half of configs[4]).

    python bench/pretrain_step.py [--config cifar|stl10|both] [--steps 8] [--warmup 3]

The encoder and projection head are the reference's PyTorch modules restated here (the reference checkout does not
travel to the GPU box): ResNet-50 with the CIFAR stem -- 3x3 stride-1 conv, no max-pool (models/resnets.py:8-36) -- or
the stock stem for 96x96 inputs, followed by g = Linear(2048,2048)-BN-ReLU-Linear(2048,128,no bias)-BN
(models/simclr.py:27-41).  The step is the reference's inner loop (utils/model_utils.py:110-123, accum_steps = 1):
two separate forwards (per-view BatchNorm statistics), loss, `loss /= accum_steps`, `loss.item()`, backward, Adam
(lr 1e-3, weight decay 1e-6, pretrain.py:80).  Synthetic inputs.  Timed with CUDA events around whole steps, once with
`objective.contrastive_loss` of this repository (the drop-in) and once with the reference's loss arithmetic restated
in eager torch ops (objective.py:23-53); the loss alone (forward + backward on detached embeddings) is timed next to
them; further arms: the projection head's last BatchNorm fused into the loss, the whole step captured in a CUDA graph
(accuracy left on the device), and the eager step under bf16 autocast + channels_last.  One JSON line per configuration.  The loss is ~0.01 % of the step's FLOPs: the point of this
harness is that swapping it in changes nothing else and removes the loss's share of the step.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torchvision.models.resnet import Bottleneck, ResNet  # noqa: E402


class Encoder(ResNet):
    """ResNet-50 trunk up to the global average pool (reference models/resnets.py:8-36)."""

    def __init__(self, cifar_stem: bool):
        super().__init__(block=Bottleneck, layers=[3, 4, 6, 3])
        self.cifar_stem = cifar_stem
        if cifar_stem:
            self.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(64)
        del self.fc

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        if not self.cifar_stem:
            x = self.maxpool(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.avgpool(x)


class SimCLR(nn.Module):
    """f + 2-layer projection head g (reference models/simclr.py:27-46); forward returns (h, z)."""

    def __init__(self, cifar_stem: bool, feature_dim=2048, out_dim=128):
        super().__init__()
        self.f = Encoder(cifar_stem)
        self.g = nn.Sequential(nn.Flatten(), nn.Linear(feature_dim, feature_dim), nn.BatchNorm1d(feature_dim),
                               nn.ReLU(inplace=True), nn.Linear(feature_dim, out_dim, bias=False), nn.BatchNorm1d(out_dim))

    def forward(self, x):
        h = self.f(x)
        return h, self.g(h)


def reference_loss(x1, x2, temperature=1.0):
    """objective.py:23-53 restated with eager torch ops (dense 2B x 2B logits) -- the arithmetic the drop-in replaces."""
    b = x1.shape[0]
    x1, x2 = F.normalize(x1, p=2, dim=1), F.normalize(x2, p=2, dim=1)
    eye = torch.eye(b, device=x1.device)
    aa = x1 @ x1.t() / temperature - eye * 1e9
    bb = x2 @ x2.t() / temperature - eye * 1e9
    ab = x1 @ x2.t() / temperature
    ba = x2 @ x1.t() / temperature
    logits = torch.cat((torch.cat((ab, aa), 1), torch.cat((bb, ba), 1)), 0)
    labels = torch.arange(2 * b, device=x1.device)
    loss = F.cross_entropy(logits, labels)
    acc = 100.0 * (logits.argmax(1) == labels).sum().item() / (2 * b)
    return loss, acc


def time_steps(model, opt, loss_fn, x1, x2, tau, steps, warmup, fused_head=False):
    def one():
        if fused_head:
            # the projection head up to (not including) its final BatchNorm1d; that module goes into the loss
            # (pytorch_simclr_b200.bn_contrastive_loss: BatchNorm apply / backward inside the loss kernels)
            u1 = model.g[:-1](model.f(x1))
            u2 = model.g[:-1](model.f(x2))
            loss, acc = loss_fn(u1, u2, model.g[-1], temperature=tau)
        else:
            _, z1 = model(x1)                                   # utils/model_utils.py:113
            _, z2 = model(x2)                                   # :114
            loss, acc = loss_fn(z1, z2, temperature=tau)        # :115
        loss /= 1                                           # :116 (accum_steps = 1)
        val = loss.item()                                   # :117
        loss.backward()                                     # :120
        opt.step()                                          # :122
        opt.zero_grad()                                     # :123
        return val, acc

    for _ in range(warmup):
        one()
    ms = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        val, acc = one()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2], val, acc


def time_graph_captured(model, loss_fn, x1, x2, tau, steps, warmup):
    """The same step captured ONCE in a CUDA graph and replayed: two forwards, the drop-in loss with the
    accuracy left on the device (`set_lazy_accuracy`: no host synchronisation inside the step), backward, capturable Adam.
    The loss value and the accuracy are read after the replay -- the reference's `loss.item()` (:117) moves behind the step."""
    import pytorch_simclr_b200 as sb
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6, capturable=True)
    sb.set_lazy_accuracy(True)
    try:
        def one():
            _, z1 = model(x1)
            _, z2 = model(x2)
            loss, acc = loss_fn(z1, z2, temperature=tau)
            loss.backward()
            opt.step()
            return loss, acc

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(3, warmup)):
                opt.zero_grad(set_to_none=True)
                one()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            loss, acc = one()
        for _ in range(2):
            graph.replay()
        ms = []
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            graph.replay()
            val = loss.item()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms.sort()
        return ms[len(ms) // 2], val, float(acc)
    finally:
        sb.set_lazy_accuracy(False)


def time_autocast_steps(model, opt, loss_fn, x1, x2, tau, steps, warmup):
    """The eager step with bf16 autocast and channels_last activations (what a user would switch on for a B200); the
    embeddings reach the loss as bf16 tensors, i.e. on its bf16-input path."""
    model = model.to(memory_format=torch.channels_last)
    x1 = x1.contiguous(memory_format=torch.channels_last)
    x2 = x2.contiguous(memory_format=torch.channels_last)

    def one():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, z1 = model(x1)
            _, z2 = model(x2)
        loss, acc = loss_fn(z1, z2, temperature=tau)
        val = loss.item()
        loss.backward()
        opt.step()
        opt.zero_grad()
        return val, acc

    for _ in range(warmup):
        one()
    ms = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        val, acc = one()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2], val, acc


def time_loss_only(loss_fn, b, d, tau, steps=30):
    g = torch.Generator(device="cuda").manual_seed(3)
    z1 = torch.randn(b, d, device="cuda", generator=g)
    z2 = torch.randn(b, d, device="cuda", generator=g)

    def one():
        a, c = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        loss, acc = loss_fn(a, c, temperature=tau)
        loss.backward()
        return loss.item()

    for _ in range(5):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps


def run(name, batch, res, cifar_stem, steps, warmup, tau=0.5, quiet=False):
    from objective import contrastive_loss           # the drop-in module at the repository root
    from pytorch_simclr_b200 import bn_contrastive_loss
    torch.manual_seed(0)
    out = {"config": name, "batch": batch, "resolution": res, "stem": "cifar 3x3 s1, no maxpool" if cifar_stem else "7x7 s2 + maxpool",
           "temperature": tau, "optimizer": "Adam(lr=1e-3, weight_decay=1e-6)", "data": "synthetic", "dtype": "fp32 (cuDNN TF32 convs: torch default)"}
    g = torch.Generator(device="cuda").manual_seed(1)
    x1 = torch.randn(batch, 3, res, res, device="cuda", generator=g)
    x2 = torch.randn(batch, 3, res, res, device="cuda", generator=g)
    for label, fn in (("ours", contrastive_loss), ("ours_fused_head_tail", bn_contrastive_loss),
                      ("reference_arithmetic", reference_loss)):
        torch.manual_seed(0)
        model = SimCLR(cifar_stem).cuda().train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6)
        fused = fn is bn_contrastive_loss
        ms, val, acc = time_steps(model, opt, fn, x1, x2, tau, steps, warmup, fused_head=fused)
        out[label] = {"ms_per_step": ms, "images_per_s": batch / (ms * 1e-3), "last_loss": val, "last_acc": acc}
        if not fused:
            out[label]["loss_only_ms"] = time_loss_only(fn, batch, 128, tau)
        del model, opt
        torch.cuda.empty_cache()
    # the same step without host synchronisation, captured in a CUDA graph; and with bf16 autocast
    for label in ("ours_graph_captured", "ours_bf16_autocast_channels_last", "reference_arithmetic_bf16_autocast_channels_last"):
        try:
            torch.manual_seed(0)
            model = SimCLR(cifar_stem).cuda().train()
            if label == "ours_graph_captured":
                ms, val, acc = time_graph_captured(model, contrastive_loss, x1, x2, tau, steps, warmup)
            else:
                opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6)
                fn = contrastive_loss if label.startswith("ours") else (lambda a, c, temperature: reference_loss(a.float(), c.float(), temperature))
                ms, val, acc = time_autocast_steps(model, opt, fn, x1, x2, tau, steps, warmup)
                del opt
            out[label] = {"ms_per_step": ms, "images_per_s": batch / (ms * 1e-3), "last_loss": val, "last_acc": acc}
            del model
        except Exception as e:                                  # an arm that cannot run must not take the others down
            out[label] = {"skipped": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    out["step_ratio_ours_over_reference"] = out["ours"]["ms_per_step"] / out["reference_arithmetic"]["ms_per_step"]
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="both", choices=["cifar", "stl10", "both"])
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    if args.config in ("cifar", "both"):
        run("pretrain step, ResNet-50 CIFAR stem + 2-layer head, batch 512, 32x32 (BASELINE configs[1])", 512, 32, True,
            args.steps, args.warmup)
    if args.config in ("stl10", "both"):
        run("pretrain step, ResNet-50 + 2-layer head, batch 256, 96x96 STL-10 shape (BASELINE configs[4])", 256, 96, False,
            args.steps, args.warmup)
