"""CPU tests: pin the oracle (dense ports + fp64 closed forms) to the reference's own outputs.

Fixtures in tests/golden were produced by oracle/make_golden.py from the unmodified reference
objective.py (reference has no tests of its own: SURVEY.md section 4).
"""
import os

import numpy as np
import pytest
import torch

import contrastive_oracle as oracle
from conftest import golden_files, load_golden


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("path", golden_files("ntxent"), ids=os.path.basename)
def test_ntxent_closed_form_matches_reference_fixture(path):
    g = load_golden(path)
    w = g["weight"] if g["weight"].size else None
    res = oracle.ntxent_closed_form(g["z1"], g["z2"], temperature=float(g["temperature"]),
                                    normalize=bool(g["normalize"]), weight=w, grad_output=float(g["grad_output"]))
    assert res.loss == pytest.approx(float(g["loss"]), rel=2e-6, abs=1e-7)
    assert res.acc == pytest.approx(float(g["acc"]), abs=1e-9)      # integer count: exact
    # fixtures are fp32 autograd results; fp64 closed form must agree to fp32 round-off
    assert _rel(res.grad1, g["grad1"]) < 2e-5
    assert _rel(res.grad2, g["grad2"]) < 2e-5


@pytest.mark.parametrize("path", golden_files("ntxent"), ids=os.path.basename)
def test_ntxent_dense_port_matches_reference_fixture(path):
    g = load_golden(path)
    torch.set_num_threads(1)
    w = torch.from_numpy(g["weight"]) if g["weight"].size else None
    kw = dict(temperature=float(g["temperature"]), normalize=bool(g["normalize"]))
    if w is not None:
        kw["weight"] = w
    loss, acc, g1, g2 = oracle.dense_port_with_grads(oracle.ntxent_dense_port, torch.from_numpy(g["z1"]),
                                                     torch.from_numpy(g["z2"]), float(g["grad_output"]), **kw)
    assert loss == pytest.approx(float(g["loss"]), rel=1e-6, abs=1e-7)
    assert acc == float(g["acc"])
    assert _rel(g1, g["grad1"]) < 1e-5
    assert _rel(g2, g["grad2"]) < 1e-5


@pytest.mark.parametrize("path", golden_files("modified"), ids=os.path.basename)
def test_modified_closed_form_matches_reference_fixture(path):
    g = load_golden(path)
    res = oracle.modified_closed_form(g["z1"], g["z2"], temperature=float(g["temperature"]),
                                      grad_output=float(g["grad_output"]))
    assert res.loss == pytest.approx(float(g["loss"]), rel=2e-6, abs=1e-7)
    assert res.acc == pytest.approx(float(g["acc"]), abs=1e-9)
    assert _rel(res.grad1, g["grad1"]) < 2e-5
    assert _rel(res.grad2, g["grad2"]) < 2e-5


@pytest.mark.parametrize("path", golden_files("modified"), ids=os.path.basename)
def test_modified_dense_port_matches_reference_fixture(path):
    g = load_golden(path)
    torch.set_num_threads(1)
    kw = {} if bool(g["default_tau"]) else dict(temperature=float(g["temperature"]))
    loss, acc, g1, g2 = oracle.dense_port_with_grads(oracle.modified_dense_port, torch.from_numpy(g["z1"]),
                                                     torch.from_numpy(g["z2"]), float(g["grad_output"]), **kw)
    assert loss == pytest.approx(float(g["loss"]), rel=1e-6, abs=1e-7)
    assert acc == float(g["acc"])
    assert _rel(g1, g["grad1"]) < 1e-5
    assert _rel(g2, g["grad2"]) < 1e-5


@pytest.mark.parametrize("n,d,tau,kind", [(96, 128, 0.5, "iid"), (300, 64, 0.1, "correlated"), (33, 256, 1.0, "iid")])
def test_oracle_against_live_reference(reference_objective, n, d, tau, kind):
    """Only in the build container: fresh seeds, not the committed fixtures."""
    z1, z2 = oracle.make_embeddings(n, d, seed=n + d, kind=kind)
    a = z1.clone().requires_grad_(True)
    b = z2.clone().requires_grad_(True)
    loss, acc = reference_objective.contrastive_loss(a, b, temperature=tau)
    loss.backward()
    res = oracle.ntxent_closed_form(z1, z2, temperature=tau)
    assert res.loss == pytest.approx(float(loss), rel=2e-6)
    assert res.acc == acc
    assert _rel(res.grad1, a.grad.numpy()) < 2e-5
    a2 = z1.clone().requires_grad_(True)
    b2 = z2.clone().requires_grad_(True)
    loss, acc = reference_objective.modified_contrastive_loss(a2, b2, temperature=tau)
    loss.backward()
    res = oracle.modified_closed_form(z1, z2, temperature=tau)
    assert res.loss == pytest.approx(float(loss), rel=2e-6)
    assert res.acc == acc
    assert _rel(res.grad2, b2.grad.numpy()) < 2e-5


def test_blockwise_is_block_size_independent():
    z1, z2 = oracle.make_embeddings(150, 64, seed=5, kind="correlated")
    a = oracle.ntxent_closed_form(z1, z2, temperature=0.5, block=1024)
    b = oracle.ntxent_closed_form(z1, z2, temperature=0.5, block=37)
    assert a.loss == pytest.approx(b.loss, rel=1e-13)
    assert a.correct == b.correct
    assert np.allclose(a.grad1, b.grad1, rtol=1e-11, atol=1e-15)
    m1 = oracle.modified_closed_form(z1, z2, temperature=0.5, block=1024)
    m2 = oracle.modified_closed_form(z1, z2, temperature=0.5, block=41)
    assert m1.loss == pytest.approx(m2.loss, rel=1e-13)
    assert np.allclose(m1.grad2, m2.grad2, rtol=1e-11, atol=1e-15)


def test_size_independent_properties():
    """Properties used at full BASELINE sizes on the GPU, validated here on the oracle itself."""
    z1, z2 = oracle.make_embeddings(64, 128, seed=9)
    base = oracle.ntxent_closed_form(z1, z2, temperature=0.5)
    swapped = oracle.ntxent_closed_form(z2, z1, temperature=0.5)
    assert base.loss == pytest.approx(swapped.loss, rel=1e-12)          # loss(z1,z2) == loss(z2,z1)
    assert np.allclose(base.grad1, swapped.grad2, rtol=1e-9, atol=1e-14)
    scaled = oracle.ntxent_closed_form(4.0 * z1, 0.25 * z2, temperature=0.5)  # exact power-of-two scales
    assert base.loss == pytest.approx(scaled.loss, rel=1e-12)           # scale invariance under normalize
    # normalised rows: gradient is orthogonal to the input row (z_r . dz_r == 0)
    assert np.abs((z1.numpy() * base.grad1).sum(axis=1)).max() < 1e-12
    perm = np.random.default_rng(0).permutation(64)
    permuted = oracle.ntxent_closed_form(z1[perm], z2[perm], temperature=0.5)
    assert base.loss == pytest.approx(permuted.loss, rel=1e-12)         # consistent relabelling


def test_row_sample_checker_matches_the_closed_form():
    """bench.py's multi-GPU parity block relies on oracle.ntxent_row_sample_check (the only form that fits 2N = 65536):
    pinned here against ntxent_closed_form, which is pinned against the reference's own outputs above."""
    for seed, kind, b, d, tau in ((3, "iid", 96, 64, 0.5), (4, "correlated", 200, 128, 0.1)):
        z1, z2 = oracle.make_embeddings(b, d, seed=seed, kind=kind)
        full = oracle.ntxent_closed_form(z1, z2, temperature=tau, grad_output=0.25)
        rows = np.array([0, 1, b - 1, b, b + 7, 2 * b - 1])
        loss, correct, grads = oracle.ntxent_row_sample_check(z1, z2, tau, rows, grad_output=0.25, block=64)
        assert loss == pytest.approx(full.loss, rel=1e-12)
        assert correct == full.correct
        ref = np.concatenate((full.grad1, full.grad2))[rows]
        assert np.allclose(np.asarray(grads), ref, rtol=1e-9, atol=1e-14)


def test_row_sample_checker_counts_exact_ties_like_the_reference():
    """first-argmax rule (objective.py:51): a duplicate of the positive that PRECEDES it in the reference's column order
    takes the hit away, one that follows does not."""
    z1, z2 = oracle.make_embeddings(32, 16, seed=8)
    z1, z2 = z1.clone(), z2.clone()
    z2[20] = z2[3]          # row 3 of view 1: positive is column (view 2, 3); the copy at (view 2, 20) FOLLOWS it
    z2[1] = z2[9]           # row 9 of view 1: the copy at (view 2, 1) PRECEDES its positive (view 2, 9)
    full = oracle.ntxent_closed_form(z1, z2, temperature=0.5)
    _, correct, _ = oracle.ntxent_row_sample_check(z1, z2, 0.5, np.array([0]))
    assert correct == full.correct
