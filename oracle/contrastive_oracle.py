"""CPU oracle for the contrastive-objective hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, what the reference's ``objective.py`` computes.  It exists so that
the CUDA path can be checked; it is never imported by the product package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it.

Parity pinning
--------------
The reference ships no tests and no golden vectors, so the oracle is pinned
against outputs of the reference itself: ``oracle/make_golden.py`` imports the unmodified
``/root/reference/objective.py`` in the build container and writes ``tests/golden/*.npz``;
``tests/test_oracle.py`` checks every function below against those fixtures (and against the live
reference whenever ``/root/reference`` is present).

Two restatements per loss:

* ``*_dense_port``   -- the reference's algorithm step by step (materialised M x M logits, torch
  CPU ops, autograd for the gradients).  fp32.  This is what ``bench.py`` times as the CPU baseline
  (``cpu_baseline.kind == "port"``) because the reference is Python and does not travel to the GPU box.
* ``*_closed_form``  -- the closed forms of SURVEY.md Appendix A evaluated blockwise in fp64, never
  holding more than ``block x M`` scores.  This is the oracle for sizes the dense port cannot run
  (2N >= 32768) and the high-precision yardstick for the tolerance tests.

Notation: B images, M = 2B views, Z = [x_batch1 ; x_batch2], pos(r) = (r + B) mod M.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

SELF_MASK = 1e9        # reference objective.py:21  (VERY_LARGE_NUM)
L2_EPS = 1e-12         # F.normalize default eps, reference objective.py:26-27
SOFTPLUS_BETA = 0.8    # reference objective.py:70-71
SOFTPLUS_THRESHOLD = 20.0  # torch default threshold used by F.softplus
CLAMP_MIN = 1e-4       # reference objective.py:87-88


# --------------------------------------------------------------------------------------------
# Dense ports (fp32, torch CPU).  Same sequence of operations as the reference.
# --------------------------------------------------------------------------------------------

def ntxent_dense_port(x_batch1: torch.Tensor, x_batch2: torch.Tensor, temperature: float = 1.0,
                      normalize: bool = True, weight: Optional[torch.Tensor] = None):
    """NT-Xent exactly as reference objective.py:23-55 computes it.

    Returns ``(loss, acc)``: loss is a 0-d tensor attached to autograd, acc a python float in [0,100].
    """
    n = x_batch1.shape[0]
    # objective.py:25-30 -- optional row-wise L2 normalisation
    u = F.normalize(x_batch1, p=2, dim=1) if normalize else x_batch1
    v = F.normalize(x_batch2, p=2, dim=1) if normalize else x_batch2
    eye = torch.eye(n, device=u.device)
    # objective.py:35-36,39-40 -- same-view similarities with the self term pushed to -1e9
    uu = u.matmul(u.t()) / temperature - eye * SELF_MASK
    vv = v.matmul(v.t()) / temperature - eye * SELF_MASK
    # objective.py:42-43 -- cross-view similarities
    uv = u.matmul(v.t()) / temperature
    vu = v.matmul(u.t()) / temperature
    # objective.py:48-49 -- row r<B sees [uv | uu], row B+i sees [vv | vu]; target column is the row index
    top = torch.cat((uv, uu), dim=1)
    bottom = torch.cat((vv, vu), dim=1)
    logits = torch.cat((top, bottom), dim=0)
    target = torch.arange(2 * n, device=u.device)
    # objective.py:47,50 -- weighted mean cross entropy
    loss = F.cross_entropy(logits, target, weight=weight, reduction="mean")
    # objective.py:51-53 -- first-argmax accuracy
    hits = int((logits.argmax(dim=1) == target).sum())
    return loss, 100.0 * hits / (2 * n)


def modified_dense_port(x_batch1: torch.Tensor, x_batch2: torch.Tensor, **kwargs):
    """The probabilistic ("--new_loss" / --modified_loss) variant as reference objective.py:68-98."""
    tau = kwargs.get("temperature", 1.0)                        # objective.py:68
    a = F.softplus(x_batch1, beta=SOFTPLUS_BETA)                 # objective.py:70-71
    b = F.softplus(x_batch2, beta=SOFTPLUS_BETA)
    n = a.shape[0]
    a = F.normalize(a, p=1, dim=1)                               # objective.py:77-78
    b = F.normalize(b, p=1, dim=1)
    idx = torch.arange(n, device=a.device)
    target = torch.cat((idx, idx))                               # objective.py:80-81
    ab = torch.clamp(a.matmul(b.t()) * n, min=CLAMP_MIN)         # objective.py:87-88
    ba = torch.clamp(b.matmul(a.t()) * n, min=CLAMP_MIN)
    logits = torch.cat((ab.log() / tau, ba.log() / tau), dim=0)  # objective.py:89-93
    loss = F.cross_entropy(logits, target, reduction="mean")     # objective.py:92,94
    hits = int((logits.argmax(dim=1) == target).sum())           # objective.py:95-96
    return loss, 100.0 * hits / (2 * n)


# --------------------------------------------------------------------------------------------
# Closed forms (fp64, blockwise).  SURVEY.md Appendix A.1 / A.2.
# --------------------------------------------------------------------------------------------

@dataclass
class OracleResult:
    loss: float
    correct: int           # number of rows whose first-argmax is the positive
    acc: float             # 100 * correct / M
    grad1: Optional[np.ndarray]
    grad2: Optional[np.ndarray]
    lse: np.ndarray        # per-row log-sum-exp (natural log), length M
    row_loss: np.ndarray   # per-row loss L_r, length M
    margin: Optional[np.ndarray] = None   # per row: best negative logit - positive logit (how close the argmax decision is)


def _as_f64(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().to(torch.float32).cpu().numpy() if x.dtype == torch.bfloat16 else x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def first_argmax_is_positive(scores: np.ndarray, rows: np.ndarray, n: int) -> np.ndarray:
    """Reference tie rule (objective.py:48-51): ``Tensor.max`` returns the first maximal index in the
    *permuted* column order [view-2 block | view-1 block] for rows of view 1 and [view-2 | view-1]
    for rows of view 2 (both are "other block as laid out by the two torch.cat calls").

    ``scores`` is [len(rows), M] in natural column order with the diagonal already set to -inf.
    Row r < n sees columns ordered [n..2n-1, 0..n-1]; row r >= n sees [n..2n-1, 0..n-1] as well
    (objective.py:48-49: top = [ab | aa], bottom = [bb | ba]).
    """
    m = 2 * n
    perm = np.concatenate((np.arange(n, m), np.arange(0, n)))
    permuted = scores[:, perm]
    pred = permuted.argmax(axis=1)                  # numpy argmax = first maximal index
    # label of row r is r itself in the permuted frame: row r<n -> column ab[r,r] at index r;
    # row n+i -> column ba[i,i] at index n+i.
    return pred == rows


def ntxent_closed_form(x_batch1, x_batch2, temperature: float = 1.0, normalize: bool = True,
                       weight=None, grad_output: float = 1.0, need_grad: bool = True,
                       block: int = 1024) -> OracleResult:
    """NT-Xent loss, accuracy count and input gradients from Appendix A.1, in fp64, blockwise.

    Follows reference objective.py:23-53; backward is the analytic derivative of that graph.
    """
    z = np.concatenate((_as_f64(x_batch1), _as_f64(x_batch2)), axis=0)
    m, d = z.shape
    n = m // 2
    if normalize:
        nrm = np.maximum(np.sqrt((z * z).sum(axis=1)), L2_EPS)      # objective.py:26-27
        zh = z / nrm[:, None]
    else:
        nrm = np.ones(m)
        zh = z
    w = np.ones(m) if weight is None else _as_f64(weight)
    wsum = w.sum()
    inv_tau = 1.0 / temperature
    pos = (np.arange(m) + n) % m

    lse = np.empty(m)
    s_pos = np.empty(m)
    margin = np.empty(m)
    correct = 0
    for r0 in range(0, m, block):
        r1 = min(m, r0 + block)
        rows = np.arange(r0, r1)
        s = zh[r0:r1] @ zh.T * inv_tau                               # objective.py:35-36,42-43
        s_pos[r0:r1] = s[np.arange(r1 - r0), pos[r0:r1]]
        s[np.arange(r1 - r0), rows] = -np.inf                         # objective.py:39-40 (exact in fp32 too)
        mx = s.max(axis=1)
        lse[r0:r1] = mx + np.log(np.exp(s - mx[:, None]).sum(axis=1))
        correct += int(first_argmax_is_positive(s, rows, n).sum())
        neg = s.copy()
        neg[np.arange(r1 - r0), pos[r0:r1]] = -np.inf
        margin[r0:r1] = neg.max(axis=1) - s_pos[r0:r1] if m > 2 else -np.inf
    row_loss = lse - s_pos
    loss = float((w * row_loss).sum() / wsum)                         # objective.py:47,50

    g1 = g2 = None
    if need_grad:
        g = grad_output * w / wsum
        dzh = np.zeros_like(zh)
        for r0 in range(0, m, block):
            r1 = min(m, r0 + block)
            rows = np.arange(r0, r1)
            s = zh[r0:r1] @ zh.T * inv_tau
            # W[r,c] = g_r P[r,c] + g_c P[c,r]  (S is symmetric, so P[c,r] = exp(S[r,c] - lse_c))
            wmat = g[r0:r1, None] * np.exp(s - lse[r0:r1, None]) + g[None, :] * np.exp(s - lse[None, :])
            wmat[np.arange(r1 - r0), rows] = 0.0
            wmat[np.arange(r1 - r0), pos[r0:r1]] -= g[r0:r1] + g[pos[r0:r1]]
            dzh[r0:r1] = wmat @ zh * inv_tau
        if normalize:
            dz = (dzh - zh * (zh * dzh).sum(axis=1, keepdims=True)) / nrm[:, None]
            # rows whose norm was clamped by eps have a zero Jacobian through the clamp's max()
            clamped = np.sqrt((z * z).sum(axis=1)) < L2_EPS
            dz[clamped] = dzh[clamped] / L2_EPS
        else:
            dz = dzh
        g1, g2 = dz[:n], dz[n:]
    return OracleResult(loss, correct, 100.0 * correct / m, g1, g2, lse, row_loss, margin)


def ntxent_row_sample_check(x_batch1, x_batch2, temperature: float, sample_rows, grad_output: float = 1.0,
                            block: int = 2048, threads: Optional[int] = None):
    """Blockwise fp64 NT-Xent at sizes where nothing M x M can exist (2N = 65536): the exact global loss, the
    first-argmax count over ALL rows, and the input gradients of the rows in ``sample_rows`` (indices into
    Z = [x_batch1; x_batch2]).  Same closed forms as ``ntxent_closed_form`` (reference objective.py:23-53, normalised,
    unweighted); torch CPU float64 so that the elementwise passes use every host core.

    Returns (loss, correct, grads[len(sample_rows), d]).
    """
    if threads:
        torch.set_num_threads(threads)
    z = torch.cat((torch.as_tensor(_as_f64(x_batch1)), torch.as_tensor(_as_f64(x_batch2))), 0)
    m, d = z.shape
    n = m // 2
    nrm = z.norm(dim=1).clamp_min(L2_EPS)                          # objective.py:26-27
    zh = z / nrm[:, None]
    inv_tau = 1.0 / temperature
    lse = torch.empty(m, dtype=torch.float64)
    s_pos = torch.empty(m, dtype=torch.float64)
    correct = 0
    perm = torch.cat((torch.arange(n, m), torch.arange(0, n)))     # objective.py:48-49 column order
    for r0 in range(0, m, block):
        r1 = min(m, r0 + block)
        rows = torch.arange(r0, r1)
        s = (zh[r0:r1] @ zh.T) * inv_tau                            # objective.py:35-36,42-43
        pos = (rows + n) % m
        s_pos[r0:r1] = s[torch.arange(r1 - r0), pos]
        s[torch.arange(r1 - r0), rows] = -float("inf")              # objective.py:39-40
        lse[r0:r1] = torch.logsumexp(s, dim=1)
        # first maximal index in the reference's permuted order; the label of row r is r in that frame
        pred = s[:, perm].argmax(dim=1)
        correct += int((pred == rows).sum())
    loss = float((lse - s_pos).mean())                              # objective.py:47,50
    sample = torch.as_tensor(np.asarray(sample_rows, dtype=np.int64))
    g = grad_output / m
    grads = torch.empty((len(sample), d), dtype=torch.float64)
    for i0 in range(0, len(sample), 256):
        rows = sample[i0:i0 + 256]
        k = len(rows)
        s = (zh[rows] @ zh.T) * inv_tau
        w = g * (torch.exp(s - lse[rows][:, None]) + torch.exp(s - lse[None, :]))
        w[torch.arange(k), rows] = 0.0
        w[torch.arange(k), (rows + n) % m] -= 2.0 * g
        dzh = (w @ zh) * inv_tau
        zr = zh[rows]
        grads[i0:i0 + k] = (dzh - zr * (zr * dzh).sum(dim=1, keepdim=True)) / nrm[rows][:, None]
    return loss, correct, grads.numpy()


def _softplus64(x: np.ndarray) -> np.ndarray:
    bx = SOFTPLUS_BETA * x
    out = np.where(bx > SOFTPLUS_THRESHOLD, x, np.log1p(np.exp(np.minimum(bx, SOFTPLUS_THRESHOLD))) / SOFTPLUS_BETA)
    return out


def modified_closed_form(x_batch1, x_batch2, temperature: float = 1.0, grad_output: float = 1.0,
                         need_grad: bool = True, block: int = 1024) -> OracleResult:
    """Probabilistic loss from Appendix A.2 in fp64, blockwise (reference objective.py:68-98).

    Row r of the 2B x B logit matrix is view-1 item r against all view-2 items for r < B, and view-2
    item r-B against all view-1 items otherwise (objective.py:87-93: ``ba`` is ``ab`` transposed).
    """
    x1 = _as_f64(x_batch1)
    x2 = _as_f64(x_batch2)
    n, d = x1.shape
    m = 2 * n
    s1, s2 = _softplus64(x1), _softplus64(x2)
    l1a = np.maximum(np.abs(s1).sum(axis=1), L2_EPS)
    l1b = np.maximum(np.abs(s2).sum(axis=1), L2_EPS)
    p1, p2 = s1 / l1a[:, None], s2 / l1b[:, None]
    inv_tau = 1.0 / temperature
    src = (p1, p2)

    lse = np.empty(m)
    a_pos = np.empty(m)
    margin = np.empty(m)
    correct = 0
    for view in (0, 1):
        mine, other = src[view], src[1 - view]
        for r0 in range(0, n, block):
            r1 = min(n, r0 + block)
            q = np.maximum(mine[r0:r1] @ other.T * n, CLAMP_MIN)         # objective.py:87-88
            a = np.log(q) * inv_tau                                       # objective.py:89-90
            mx = a.max(axis=1)
            lse[view * n + r0:view * n + r1] = mx + np.log(np.exp(a - mx[:, None]).sum(axis=1))
            a_pos[view * n + r0:view * n + r1] = a[np.arange(r1 - r0), np.arange(r0, r1)]
            correct += int((a.argmax(axis=1) == np.arange(r0, r1)).sum())  # objective.py:95-96
            neg = a.copy()
            neg[np.arange(r1 - r0), np.arange(r0, r1)] = -np.inf
            margin[view * n + r0:view * n + r1] = neg.max(axis=1) - a_pos[view * n + r0:view * n + r1] if n > 1 else -np.inf
    row_loss = lse - a_pos
    loss = float(row_loss.mean())                                          # objective.py:92-94

    g1 = g2 = None
    if need_grad:
        gscale = grad_output / m
        dp = [np.zeros_like(p1), np.zeros_like(p2)]
        for view in (0, 1):
            mine, other = src[view], src[1 - view]
            lse_mine = lse[view * n:(view + 1) * n]
            lse_other = lse[(1 - view) * n:(2 - view) * n]
            for r0 in range(0, n, block):
                r1 = min(n, r0 + block)
                praw = mine[r0:r1] @ other.T
                q = np.maximum(praw * n, CLAMP_MIN)
                a = np.log(q) * inv_tau
                da = np.exp(a - lse_mine[r0:r1, None]) + np.exp(a - lse_other[None, :])
                da[np.arange(r1 - r0), np.arange(r0, r1)] -= 2.0
                live = (praw * n >= CLAMP_MIN)
                dpm = gscale * da * inv_tau / np.where(live, praw, 1.0) * live
                dp[view][r0:r1] = dpm @ other
        grads = []
        for view, (x, s, p, l1) in enumerate(((x1, s1, p1, l1a), (x2, s2, p2, l1b))):
            ds = (dp[view] - (dp[view] * p).sum(axis=1, keepdims=True)) / l1[:, None]
            sig = 1.0 / (1.0 + np.exp(-SOFTPLUS_BETA * x))
            grads.append(ds * np.where(SOFTPLUS_BETA * x > SOFTPLUS_THRESHOLD, 1.0, sig))
        g1, g2 = grads
    return OracleResult(loss, correct, 100.0 * correct / m, g1, g2, lse, row_loss, margin)


# --------------------------------------------------------------------------------------------
# Helpers shared by tests / bench
# --------------------------------------------------------------------------------------------

def make_embeddings(n: int, d: int, seed: int = 0, kind: str = "iid", noise: float = 0.5,
                    bf16_representable: bool = False):
    """Seeded synthetic embeddings on the CPU.

    ``iid``: two independent randn batches (accuracy ~ 0 %).  ``correlated``: z_k = base + noise*randn
    (accuracy ~ 100 %).  ``bf16_representable`` rounds to bf16 and back so the bf16-input contract can
    feed identical values to the fp32 reference.
    """
    gen = torch.Generator().manual_seed(seed)
    if kind == "iid":
        z1 = torch.randn(n, d, generator=gen)
        z2 = torch.randn(n, d, generator=gen)
    elif kind == "correlated":
        base = torch.randn(n, d, generator=gen)
        z1 = base + noise * torch.randn(n, d, generator=gen)
        z2 = base + noise * torch.randn(n, d, generator=gen)
    else:
        raise ValueError(kind)
    if bf16_representable:
        z1 = z1.to(torch.bfloat16).to(torch.float32)
        z2 = z2.to(torch.bfloat16).to(torch.float32)
    return z1, z2


def dense_port_with_grads(fn, z1: torch.Tensor, z2: torch.Tensor, grad_output: float = 1.0, **kw):
    """Run a dense port forward+backward; returns (loss float, acc, grad1, grad2) as numpy."""
    a = z1.detach().clone().requires_grad_(True)
    b = z2.detach().clone().requires_grad_(True)
    loss, acc = fn(a, b, **kw)
    (loss * grad_output).backward()
    return float(loss.detach()), acc, a.grad.numpy(), b.grad.numpy()


def algorithmic_flops(m: int, d: int, modified: bool = False) -> float:
    """6*M^2*d for NT-Xent fwd+bwd, 3*M^2*d for the modified loss."""
    return (3.0 if modified else 6.0) * float(m) * float(m) * float(d)
