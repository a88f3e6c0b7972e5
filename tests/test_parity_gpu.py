"""GPU parity tests: the CUDA path (through the public objective API -> C ABI) against

  * the committed golden fixtures produced by the unmodified reference (tests/golden, oracle/make_golden.py),
  * the fp64 closed-form oracle on seeded inputs at sizes it finishes in seconds,
  * size-independent properties at the full BASELINE size (2N = 8192).

Tolerances are BASELINE.json's north_star contracts, each test running in both arithmetic modes:
  precision "bf16" (bf16 tensor-core operands, fp32 accumulate): loss within 2e-3 relative, gradients within 1e-2 of
                   max|reference gradient|;
  precision "fp32" (split hi+lo bf16 operands, the fp32/TF32-grade path): loss within 1e-5 relative, gradients within
                   1e-4 of max|reference gradient|;
accuracy (an integer count, objective.py:51-53 / :95-97) exact in BOTH modes: the bf16 mode re-scores in exact fp32 every
negative whose tensor-core score lies within the bf16 error bound of the exact positive (candidates recorded by the
forward tile kernel, decided by the forward finalize kernel).
"""
import os

import numpy as np
import pytest
import torch

import contrastive_oracle as oracle
import pytorch_simclr_b200 as sb
from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu

LOSS_RTOL = 2e-3
GRAD_TOL = 1e-2
TOL = {"bf16": (2e-3, 1e-2), "fp32": (1e-5, 1e-4)}     # precision mode -> (loss rtol, gradient tolerance)
PRECISIONS = ["bf16", "fp32"]


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    sb.set_precision("auto")
    sb.set_eager_backward(True)
    sb.set_deterministic(None)
    sb.set_lazy_accuracy(False)


def _mode_for(precision, dtype, d):
    """The mode a case can run in: bf16 inputs and d > 128 only have the bf16 contract."""
    if precision == "fp32" and (dtype != torch.float32 or d > 128):
        pytest.skip("fp32-grade mode needs float32 inputs and d <= 128")
    return precision


def _grad_err(g, ref):
    scale = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(g.astype(np.float64) - ref).max() / scale)


def _run(fn, z1, z2, grad_output=1.0, dtype=torch.float32, precision=None, **kw):
    if precision is not None:
        sb.set_precision(precision)
    a = z1.to(device="cuda", dtype=dtype).requires_grad_(True)
    b = z2.to(device="cuda", dtype=dtype).requires_grad_(True)
    loss, acc = fn(a, b, **kw)
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.is_cuda and isinstance(acc, float)
    if grad_output != 1.0:
        loss = loss * grad_output
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()) / grad_output, acc, a.grad.float().cpu().numpy(), b.grad.float().cpu().numpy()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("path", golden_files("ntxent"), ids=os.path.basename)
def test_ntxent_matches_reference_fixture(path, precision):
    g = load_golden(path)
    kw = dict(temperature=float(g["temperature"]), normalize=bool(g["normalize"]))
    if g["weight"].size:
        kw["weight"] = torch.from_numpy(g["weight"]).cuda()
    dtype = torch.bfloat16 if bool(g["bf16"]) else torch.float32
    _mode_for(precision, dtype, g["z1"].shape[1])
    LOSS_RTOL, GRAD_TOL = TOL[precision]
    # the fixtures are the reference's own fp32 results: their rounding noise (~1e-6 of the gradient scale) is the
    # floor of the comparison
    loss, acc, g1, g2 = _run(sb.contrastive_loss, torch.from_numpy(g["z1"]), torch.from_numpy(g["z2"]),
                             float(g["grad_output"]), dtype, precision, **kw)
    assert loss == pytest.approx(float(g["loss"]), rel=LOSS_RTOL, abs=2e-6)
    if "ties" in path:
        # duplicated rows: the tie rule of objective.py:51 decides; bf16 rounding keeps exact ties exact
        assert acc == float(g["acc"])
    else:
        assert acc == float(g["acc"])
    if float(np.abs(g["grad1"]).max()) > 0:
        tol = GRAD_TOL if bool(g["normalize"]) else 4 * GRAD_TOL   # unnormalised logits reach +-100
        assert _grad_err(g1, g["grad1"]) < tol
        assert _grad_err(g2, g["grad2"]) < tol
    else:
        assert np.abs(g1).max() < 1e-6 and np.abs(g2).max() < 1e-6


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("path", golden_files("modified"), ids=os.path.basename)
def test_modified_matches_reference_fixture(path, precision):
    g = load_golden(path)
    kw = {} if bool(g["default_tau"]) else dict(temperature=float(g["temperature"]))
    dtype = torch.bfloat16 if bool(g["bf16"]) else torch.float32
    _mode_for(precision, dtype, g["z1"].shape[1])
    LOSS_RTOL, GRAD_TOL = TOL[precision]
    loss, acc, g1, g2 = _run(sb.modified_contrastive_loss, torch.from_numpy(g["z1"]), torch.from_numpy(g["z2"]),
                             float(g["grad_output"]), dtype, precision, **kw)
    assert loss == pytest.approx(float(g["loss"]), rel=LOSS_RTOL, abs=2e-6)
    assert acc == float(g["acc"])
    if float(np.abs(g["grad1"]).max()) > 0:
        assert _grad_err(g1, g["grad1"]) < GRAD_TOL
        assert _grad_err(g2, g["grad2"]) < GRAD_TOL


@pytest.mark.parametrize("b,d,tau,kind,dtype", [
    (512, 128, 0.5, "iid", torch.float32),            # BASELINE configs[0] shape
    (1000, 256, 0.5, "iid", torch.float32),
    (777, 100, 0.1, "correlated", torch.float32),     # ragged rows and ragged feature dim
    (2048, 64, 0.5, "correlated", torch.bfloat16),
    (4096, 128, 0.5, "iid", torch.bfloat16),          # BASELINE metric shape, bf16 in
    (4096, 128, 0.1, "correlated", torch.float32),
])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_ntxent_against_fp64_oracle(b, d, tau, kind, dtype, precision):
    _mode_for(precision, dtype, d)
    LOSS_RTOL, GRAD_TOL = TOL[precision]
    z1, z2 = oracle.make_embeddings(b, d, seed=b + d, kind=kind, bf16_representable=(dtype == torch.bfloat16))
    ref = oracle.ntxent_closed_form(z1, z2, temperature=tau)
    loss, acc, g1, g2 = _run(sb.contrastive_loss, z1, z2, 1.0, dtype, precision, temperature=tau)
    assert loss == pytest.approx(ref.loss, rel=LOSS_RTOL)
    assert acc == ref.acc                # the count is exact in both modes
    assert _grad_err(g1, ref.grad1) < GRAD_TOL and _grad_err(g2, ref.grad2) < GRAD_TOL


@pytest.mark.parametrize("b,d,tau,kind,dtype", [
    (512, 128, 0.5, "iid", torch.float32),
    (777, 100, 1.0, "correlated", torch.float32),
    (4096, 128, 0.5, "iid", torch.bfloat16),          # BASELINE configs[2]
    (4096, 128, 0.1, "correlated", torch.bfloat16),   # general pow path
])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_modified_against_fp64_oracle(b, d, tau, kind, dtype, precision):
    _mode_for(precision, dtype, d)
    LOSS_RTOL, GRAD_TOL = TOL[precision]
    z1, z2 = oracle.make_embeddings(b, d, seed=b + d + 1, kind=kind, bf16_representable=(dtype == torch.bfloat16))
    ref = oracle.modified_closed_form(z1, z2, temperature=tau)
    loss, acc, g1, g2 = _run(sb.modified_contrastive_loss, z1, z2, 1.0, dtype, precision, temperature=tau)
    assert loss == pytest.approx(ref.loss, rel=LOSS_RTOL)
    # the modified loss compares products of probabilities that agree to three digits: only the exact re-scoring of the
    # near-ties makes the bf16 mode's count exact
    assert acc == ref.acc
    assert _grad_err(g1, ref.grad1) < GRAD_TOL and _grad_err(g2, ref.grad2) < GRAD_TOL


def test_general_two_exp_backward_path_small_temperature():
    """tau = 0.02 makes 2*log2(e)/tau > 80, which disables the one-exp (bounded score) backward form."""
    z1, z2 = oracle.make_embeddings(300, 128, seed=4, kind="correlated", noise=1.0)
    ref = oracle.ntxent_closed_form(z1, z2, temperature=0.02)
    loss, acc, g1, g2 = _run(sb.contrastive_loss, z1, z2, precision="bf16", temperature=0.02)
    assert loss == pytest.approx(ref.loss, rel=LOSS_RTOL, abs=1e-5)
    assert acc == ref.acc
    assert _grad_err(g1, ref.grad1) < 3 * GRAD_TOL      # logits span +-50: bf16 operands, documented in DESIGN.md


def test_drop_in_behaviours_the_callers_rely_on():
    """reference utils/model_utils.py:24-36 (eval under no_grad) and :115-120 (in-place division, backward)."""
    z1, z2 = oracle.make_embeddings(64, 128, seed=2)
    a = z1.cuda().requires_grad_(True)
    b = z2.cuda().requires_grad_(True)
    with torch.no_grad():
        loss, acc = sb.contrastive_loss(a, b, temperature=0.5)
        assert not loss.requires_grad
        loss /= 8
    loss, acc = sb.contrastive_loss(a, b, temperature=0.5)
    assert loss.requires_grad and loss.grad_fn is not None
    full = float(loss.detach())
    loss /= 8                                   # model_utils.py:116, in place on the returned tensor
    assert float(loss.detach()) == pytest.approx(full / 8, rel=1e-6)
    loss.backward()
    ref = oracle.ntxent_closed_form(z1, z2, temperature=0.5, grad_output=1.0 / 8)
    assert _grad_err(a.grad.cpu().numpy(), ref.grad1) < GRAD_TOL
    # positional call exactly like the reference's callers, modified loss ignores unknown kwargs (:68)
    l2, acc2 = sb.modified_contrastive_loss(a, b, temperature=0.5, normalize=False, bogus=1)
    assert torch.isfinite(l2) and 0.0 <= acc2 <= 100.0


def test_fp16_and_fp64_inputs_are_computed_in_fp32():
    z1, z2 = oracle.make_embeddings(96, 64, seed=8, kind="correlated")
    ref = oracle.ntxent_closed_form(z1, z2, temperature=0.5)
    for dt in (torch.float64, torch.float16):
        a = z1.to("cuda", dt).requires_grad_(True)
        b = z2.to("cuda", dt).requires_grad_(True)
        loss, acc = sb.contrastive_loss(a, b, temperature=0.5)
        loss.backward()
        assert a.grad.dtype == dt
        assert float(loss.detach()) == pytest.approx(ref.loss, rel=5e-3)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_full_size_properties_2n8192(precision):
    """Size-independent properties at the BASELINE metric shape (2N = 8192, d = 128)."""
    LOSS_RTOL, GRAD_TOL = TOL[precision]
    sb.set_precision(precision)
    b, d, tau = 4096, 128, 0.5
    z1, z2 = oracle.make_embeddings(b, d, seed=21, kind="correlated", noise=1.0)
    loss, acc, g1, g2 = _run(sb.contrastive_loss, z1, z2, temperature=tau)
    # (1) loss(z1,z2) == loss(z2,z1) and the gradients swap
    loss_s, acc_s, h1, h2 = _run(sb.contrastive_loss, z2, z1, temperature=tau)
    assert loss_s == pytest.approx(loss, rel=1e-6)
    assert np.abs(h2 - g1).max() <= 1e-3 * np.abs(g1).max()
    # (2) scale invariance under normalisation (exact power-of-two scales)
    loss_k, _, k1, _ = _run(sb.contrastive_loss, 4.0 * z1, 0.25 * z2, temperature=tau)
    assert loss_k == pytest.approx(loss, rel=1e-6)
    assert np.abs(4.0 * k1 - g1).max() <= 1e-3 * np.abs(g1).max()
    # (3) the gradient of a normalised row is orthogonal to the row:  z_r . dz_r = 0
    ortho = np.abs((z1.numpy() * g1).sum(1)) / (np.linalg.norm(z1.numpy(), axis=1) * np.linalg.norm(g1, axis=1) + 1e-30)
    assert ortho.max() < 1e-3
    # (4) consistent relabelling of the images leaves the loss unchanged and permutes the gradients
    perm = torch.from_numpy(np.random.default_rng(0).permutation(b))
    loss_p, acc_p, p1, _ = _run(sb.contrastive_loss, z1[perm], z2[perm], temperature=tau)
    assert loss_p == pytest.approx(loss, rel=1e-6)
    assert acc_p == acc
    assert np.abs(p1 - g1[perm.numpy()]).max() <= 1e-3 * np.abs(g1).max()
    # (5) gradients sum: sum_r dL/dz_r . z_r == 0 was (3); total loss bounded by log(2N-1) + 2/tau
    assert 0.0 < loss < np.log(2 * b - 1) + 2.0 / tau
    # (6) the oracle agrees at this size too (blockwise fp64, a few seconds)
    ref = oracle.ntxent_closed_form(z1, z2, temperature=tau)
    assert loss == pytest.approx(ref.loss, rel=LOSS_RTOL)
    assert acc == ref.acc
    assert _grad_err(g1, ref.grad1) < GRAD_TOL


@pytest.mark.parametrize("prec", [0, 1], ids=["bf16", "fp32"])
def test_row_sharded_kernels_equal_single_shot(prec):
    """Emulate R = 4 ranks on one GPU through the staged C ABI (row_offset / b_global), no collectives:
    per-rank stats must add up to, and per-rank gradients must equal, the single-shot result."""
    from pytorch_simclr_b200 import functional as F

    class FakeGather:
        def __init__(self, full_operand, full_lse2, b_glob, row_off, world):
            self.full_operand, self.full_lse2, self.b_glob, self.row_off = full_operand, full_lse2, b_glob, row_off
            self.world = world

        def operand(self, operand_local, b):
            return self.full_operand, self.b_glob, self.row_off

        def reduce(self, stats, loss):
            return loss, stats

        def rowvec(self, vec, b):
            return self.full_lse2

        def col_scale(self, *a):
            raise AssertionError

    b, d, tau, ranks = 1024, 128, 0.5, 4
    z1, z2 = oracle.make_embeddings(b, d, seed=5, kind="correlated", noise=1.0)
    x1, x2 = z1.cuda(), z2.cuda()
    loss, stats, rowvec, saved = F.run_forward(F.LOSS_NTXENT, x1, x2, tau, True, None, None, False, prec)
    g1, g2 = F.run_backward(saved, x1, x2, None)
    bl = b // ranks
    tot = torch.zeros(3, device="cuda")
    for r in range(ranks):
        sl = slice(r * bl, (r + 1) * bl)
        gather = FakeGather(saved.operand_cols, rowvec[2], b, r * bl, ranks)
        l_r, st_r, _, sv_r = F.run_forward(F.LOSS_NTXENT, x1[sl].contiguous(), x2[sl].contiguous(), tau, True, None, gather,
                                           False, prec)
        tot += st_r[:3]
        h1, h2 = F.run_backward(sv_r, x1[sl].contiguous(), x2[sl].contiguous(), None)
        assert torch.allclose(h1, g1[sl], rtol=1e-4, atol=1e-7 * float(g1.abs().max()) + 1e-12)
        assert torch.allclose(h2, g2[sl], rtol=1e-4, atol=1e-7 * float(g1.abs().max()) + 1e-12)
    assert float(tot[0] / tot[1]) == pytest.approx(float(loss), rel=1e-6)
    assert float(tot[2]) == float(stats[2])


@pytest.mark.parametrize("kind", [0, 1], ids=["ntxent", "modified"])
@pytest.mark.parametrize("b,d,precision", [(4096, 128, "bf16"), (4096, 128, "fp32"), (300, 100, "bf16"), (2048, 256, "bf16")])
def test_fused_step_equals_staged_calls_over_changing_inputs(kind, b, d, precision):
    """simclr_forward_backward (deferred statistics, operand loads ahead of the forward finalize kernel, input rows
    loaded ahead of griddepcontrol.wait) against the staged prepare / forward / backward calls, on a sequence of
    DIFFERENT inputs replayed back to back from one CUDA graph -- the situation in which a stale operand, column
    vector or statistic of the previous step would show."""
    from pytorch_simclr_b200 import functional as F
    from pytorch_simclr_b200.runner import ContrastiveStep
    tau, n_sets = 0.5, 3
    step = ContrastiveStep(kind, b, d, tau, True, torch.float32, "cuda", precision=precision)
    xs = []
    for s in range(n_sets):
        z1, z2 = oracle.make_embeddings(b, d, seed=100 + s, kind="correlated" if s % 2 else "iid", noise=1.0)
        xs.append((z1.cuda(), z2.cuda()))
    go = torch.tensor([0.125], device="cuda")
    outs = [(torch.empty(b, d, device="cuda"), torch.empty(b, d, device="cuda"), torch.empty(4, device="cuda"))
            for _ in range(2 * n_sets)]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step.step(go, xs[0][0], xs[0][1], outs[0][0], outs[0][1])          # warm-up outside the graph
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for i in range(2 * n_sets):
            x1, x2 = xs[i % n_sets]
            step.step(go, x1, x2, outs[i][0], outs[i][1])
            outs[i][2].copy_(step.stats)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    for i in range(2 * n_sets):
        x1, x2 = xs[i % n_sets]
        prec = F.resolve_precision(x1, False, precision)
        loss, stats, _rv, saved = F.run_forward(kind, x1, x2, tau, True, None, None, True, prec)      # staged calls
        g1, g2 = F.run_backward(saved, x1, x2, go)
        torch.cuda.synchronize()
        assert torch.equal(outs[i][2][:3], stats[:3]), f"step {i}: statistics differ"
        scale = float(g1.abs().max())
        assert float((outs[i][0] - g1).abs().max()) <= 2e-6 * scale, f"step {i}"
        assert float((outs[i][1] - g2).abs().max()) <= 2e-6 * scale, f"step {i}"
    # and against the oracle (the staged path is itself checked above; this pins the fused one independently)
    z1, z2 = xs[0][0].cpu(), xs[0][1].cpu()
    ref = (oracle.ntxent_closed_form if kind == 0 else oracle.modified_closed_form)(z1, z2, temperature=tau, grad_output=0.125)
    ltol, gtol = TOL[precision]
    assert float(outs[0][2][3]) == pytest.approx(ref.loss, rel=ltol)
    assert _grad_err(outs[0][0].cpu().numpy(), ref.grad1) < gtol


@pytest.mark.parametrize("kind", ["ntxent", "modified"])
@pytest.mark.parametrize("b,d,precision", [(512, 128, "fp32"), (4096, 128, "bf16"), (777, 100, "bf16")])
def test_eager_and_staged_autograd_paths_agree(kind, b, d, precision):
    """The autograd Function either runs the fused step at forward time and scales the kept gradients in backward()
    (default) or the staged forward / backward calls: same loss, accuracy and gradients, including the in-place
    division of the loss (utils/model_utils.py:116) and a second backward() over a retained graph."""
    fn = sb.contrastive_loss if kind == "ntxent" else sb.modified_contrastive_loss
    z1, z2 = oracle.make_embeddings(b, d, seed=31 + b, kind="correlated", noise=1.0)
    sb.set_precision(precision)
    res = {}
    for eager in (True, False):
        sb.set_eager_backward(eager)
        a = z1.cuda().requires_grad_(True)
        c = z2.cuda().requires_grad_(True)
        loss, acc = fn(a, c, temperature=0.5)
        loss /= 8
        loss.backward(retain_graph=True)
        g1 = a.grad.clone()
        loss.backward()                      # accumulates a second, identical contribution
        torch.cuda.synchronize()
        assert float((a.grad - 2 * g1).abs().max()) <= 1e-5 * float(g1.abs().max())     # staged: a recomputation
        res[eager] = (float(loss.detach()), acc, g1.cpu(), c.grad.cpu() / 2)
    assert res[True][0] == pytest.approx(res[False][0], rel=1e-7)
    assert res[True][1] == res[False][1]
    scale = float(res[False][2].abs().max())
    assert float((res[True][2] - res[False][2]).abs().max()) <= 3e-6 * scale
    assert float((res[True][3] - res[False][3]).abs().max()) <= 3e-6 * scale
    ref = (oracle.ntxent_closed_form if kind == "ntxent" else oracle.modified_closed_form)(z1, z2, temperature=0.5,
                                                                                            grad_output=1.0 / 8)
    assert _grad_err(res[True][2].numpy(), ref.grad1) < TOL[precision][1]


def test_first_argmax_tie_with_an_earlier_column_of_the_positive_tile():
    """objective.py:51 -- Tensor.max returns the FIRST maximal index.  A view-2 row identical to the positive, placed in
    the same 128-column tile as the positive but in an earlier 16-column chunk, wins the tie (the row is NOT counted as
    correct); placed later it loses (the row IS counted).  The oracle applies the reference rule exactly."""
    b, d = 256, 128
    z1, z2 = oracle.make_embeddings(b, d, seed=77, kind="correlated", noise=0.05)
    z2 = z2.clone()
    z2[130] = z2[200]          # duplicate BEFORE the positive of row 200 (same tile 128..255, chunk 0 vs chunk 4)
    z2[250] = z2[140]          # duplicate AFTER the positive of row 140
    ref = oracle.ntxent_closed_form(z1, z2, temperature=0.5)
    for precision in PRECISIONS:
        loss, acc, g1, g2 = _run(sb.contrastive_loss, z1, z2, precision=precision, temperature=0.5)
        assert acc == ref.acc, precision
        assert loss == pytest.approx(ref.loss, rel=TOL[precision][0])


def _near_tie_batch(b, d, seed, ndup=96, sigma=0.08):
    """Image ndup + i is a near-copy of image i (i < ndup): 4 * ndup rows whose best negative is within ~1e-3 of the
    positive in cosine similarity, on either side of it -- far inside bf16 resolution (2^-8)."""
    gen = torch.Generator().manual_seed(seed)
    base = torch.randn(b, d, generator=gen)
    base[ndup:2 * ndup] = base[:ndup]
    return base + sigma * torch.randn(b, d, generator=gen), base + sigma * torch.randn(b, d, generator=gen)


@pytest.mark.parametrize("kind", ["ntxent", "modified"])
@pytest.mark.parametrize("b,d", [(1024, 128), (4096, 128), (777, 100)])
def test_accuracy_count_is_exact_under_near_ties_in_bf16_mode(kind, b, d):
    """Adversarial for the bf16 tensor-core scores: hundreds of rows whose best negative is inside bf16 resolution of the
    positive.  The count must equal the fp64 oracle's (objective.py:51-53 / :95-97) in both arithmetic modes.  The batch
    is chosen (first seed that qualifies) so that no decision is closer than 3e-6 -- below that the reference's own fp32
    arithmetic is no longer order-stable against fp64 either."""
    ref_fn = oracle.ntxent_closed_form if kind == "ntxent" else oracle.modified_closed_form
    for seed in range(16):
        z1, z2 = _near_tie_batch(b, d, seed)
        ref = ref_fn(z1, z2, temperature=0.5, need_grad=False)
        gap = np.abs(ref.margin * 0.5)                  # similarity units (NT-Xent) / log-ratio units (modified)
        if gap.min() > 3e-6:
            break
    else:
        pytest.fail("no seed gives a resolvable batch")
    assert (gap < 2.0 ** -8).sum() >= 300              # the bf16 scores cannot decide these rows
    assert 50.0 < ref.acc < 99.0
    fn = sb.contrastive_loss if kind == "ntxent" else sb.modified_contrastive_loss
    for precision in PRECISIONS:
        sb.set_precision(precision)
        with torch.no_grad():
            loss, acc = fn(z1.cuda(), z2.cuda(), temperature=0.5)
        assert acc == ref.acc, (precision, acc, ref.acc)
        # and through the training path (fused begin / finish calls)
        a = z1.cuda().requires_grad_(True)
        c = z2.cuda().requires_grad_(True)
        loss, acc = fn(a, c, temperature=0.5)
        assert acc == ref.acc, (precision, "training path", acc, ref.acc)


def test_more_near_ties_than_candidate_slots_falls_back_gracefully():
    """Sixteen near-copies of every image: more negatives inside the band than the 8 candidate slots of a row.  Such
    rows keep the tensor-core decision (documented in include/simclr_b200.h): the count stays within a few rows of the
    oracle's and nothing overflows."""
    b, d = 512, 128
    gen = torch.Generator().manual_seed(5)
    base = torch.randn(b // 16, d, generator=gen)
    owner = torch.arange(b) % (b // 16)
    z1 = base[owner] + 1e-3 * torch.randn(b, d, generator=gen)
    z2 = base[owner] + 1e-3 * torch.randn(b, d, generator=gen)
    ref = oracle.ntxent_closed_form(z1, z2, temperature=0.5, need_grad=False)
    sb.set_precision("bf16")
    for _ in range(2):                                          # twice: the candidate counters must be left clean
        with torch.no_grad():
            loss, acc = sb.contrastive_loss(z1.cuda(), z2.cuda(), temperature=0.5)
        assert abs(acc - ref.acc) * 2 * b / 100.0 <= 0.25 * 2 * b
        assert loss.item() == pytest.approx(ref.loss, rel=2e-3)


@pytest.mark.parametrize("kind", [0, 1], ids=["ntxent", "modified"])
@pytest.mark.parametrize("b,d", [(4096, 128), (1000, 100), (2048, 256)])
def test_deterministic_mode_gives_bit_identical_gradients(kind, b, d):
    """SIMCLR_FLAG_DETERMINISTIC (the reference's cudnn.deterministic / manual_seed switch, pretrain.py:59-61): per-(CTA,
    segment) accumulator slots merged in a fixed order.  Repeated runs -- staged calls, the fused step, a CUDA graph
    replayed under different load -- give torch.equal gradients, and they agree with the default (reduce-add) mode to
    fp32 round-off."""
    from pytorch_simclr_b200.runner import ContrastiveStep
    z1, z2 = oracle.make_embeddings(b, d, seed=3 * b + d, kind="correlated", noise=1.0)
    det = ContrastiveStep(kind, b, d, 0.5, True, torch.float32, "cuda", "bf16", deterministic=True)
    ref = ContrastiveStep(kind, b, d, 0.5, True, torch.float32, "cuda", "bf16", deterministic=False)
    for s in (det, ref):
        s.x1.copy_(z1)
        s.x2.copy_(z2)
    ref.step()
    torch.cuda.synchronize()
    outs = []
    noise = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for rep in range(6):
        det.grad1.zero_()
        det.grad2.zero_()
        if rep % 2:
            noise.zero_()                      # perturb timing / cache state between repetitions
        if rep < 3:
            det.step()
        else:
            det.step_staged()
        torch.cuda.synchronize()
        outs.append((det.grad1.clone(), det.grad2.clone(), float(det.loss)))
    for i, (g1, g2, loss) in enumerate(outs[1:]):
        assert loss == outs[0][2], (i, loss, outs[0][2])
        assert torch.equal(g1, outs[0][0]), (i, float((g1 - outs[0][0]).abs().max()), int((g1 != outs[0][0]).sum()))
        assert torch.equal(g2, outs[0][1]), (i, float((g2 - outs[0][1]).abs().max()), int((g2 != outs[0][1]).sum()))
    scale = float(ref.grad1.abs().max())
    assert float((outs[0][0] - ref.grad1).abs().max()) <= 2e-6 * scale
    assert float((outs[0][1] - ref.grad2).abs().max()) <= 2e-6 * scale
    # and through the public API under torch.use_deterministic_algorithms
    sb.set_precision("bf16")
    sb.set_deterministic(True)
    fn = sb.contrastive_loss if kind == 0 else sb.modified_contrastive_loss
    grads = []
    for rep in range(3):
        a = z1.cuda().requires_grad_(True)
        c = z2.cuda().requires_grad_(True)
        loss, acc = fn(a, c, temperature=0.5)
        loss.backward()
        torch.cuda.synchronize()
        grads.append((a.grad.clone(), c.grad.clone()))
    assert all(torch.equal(g[0], grads[0][0]) and torch.equal(g[1], grads[0][1]) for g in grads[1:])
    assert torch.equal(grads[0][0], outs[0][0])


def test_lazy_accuracy_returns_a_device_tensor_and_allows_graph_capture():
    """set_lazy_accuracy(True): no host synchronisation inside the call, so forward + loss + backward of the public API
    can be captured into one CUDA graph (the two .item() syncs of utils/model_utils.py:117 and
    objective.py:52 are what keeps the reference's step host-bound)."""
    b, d = 1024, 128
    z1, z2 = oracle.make_embeddings(b, d, seed=9, kind="correlated", noise=1.0)
    ref = oracle.ntxent_closed_form(z1, z2, temperature=0.5)
    sb.set_precision("bf16")
    sb.set_lazy_accuracy(True)
    a = z1.cuda().requires_grad_(True)
    c = z2.cuda().requires_grad_(True)
    loss, acc = sb.contrastive_loss(a, c, temperature=0.5)
    assert isinstance(acc, torch.Tensor) and acc.is_cuda and acc.dim() == 0
    loss.backward()
    assert float(acc) == ref.acc
    # whole step under stream capture
    side = torch.cuda.Stream()
    xa, xc = z1.cuda(), z2.cuda()
    ga, gc = torch.zeros_like(xa), torch.zeros_like(xc)
    out = torch.zeros(2, device="cuda")
    with torch.cuda.stream(side):
        for _ in range(2):                                   # warm-up (allocator, library state)
            p, q = xa.clone().requires_grad_(True), xc.clone().requires_grad_(True)
            l, ac = sb.contrastive_loss(p, q, temperature=0.5)
            l.backward()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        p, q = xa.clone().requires_grad_(True), xc.clone().requires_grad_(True)
        l, ac = sb.contrastive_loss(p, q, temperature=0.5)
        l.backward()
        ga.copy_(p.grad)
        gc.copy_(q.grad)
        out[0].copy_(l.detach())
        out[1].copy_(ac)
    xa.copy_(z2)                                             # new inputs, replay
    xc.copy_(z1)
    graph.replay()
    torch.cuda.synchronize()
    ref2 = oracle.ntxent_closed_form(z2, z1, temperature=0.5)
    assert float(out[0]) == pytest.approx(ref2.loss, rel=2e-3) and float(out[1]) == ref2.acc
    assert _grad_err(ga.cpu().numpy(), ref2.grad1) < 1e-2
