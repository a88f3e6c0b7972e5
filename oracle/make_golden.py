"""Generate tests/golden/*.npz by running the UNMODIFIED reference objective.py.  TEST INFRASTRUCTURE.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/ntxent_*.npz, modified_*.npz

Every fixture stores the seeded inputs, the reference's loss, accuracy and input gradients (fp32,
torch CPU autograd) and the call arguments.  The GPU box has no /root/reference, so the parity
tests there compare against these files and against oracle/contrastive_oracle.py (itself checked
against these files in tests/test_oracle.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("REF_PATH", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")

sys.path.insert(0, HERE)
from contrastive_oracle import make_embeddings  # noqa: E402


def _load_reference():
    sys.path.insert(0, REF)
    import objective  # the reference's own module
    assert os.path.realpath(objective.__file__).startswith(os.path.realpath(REF)), objective.__file__
    return objective


def _run(fn, z1, z2, grad_output, **kw):
    a = z1.clone().requires_grad_(True)
    b = z2.clone().requires_grad_(True)
    loss, acc = fn(a, b, **kw)
    loss = loss * grad_output if grad_output != 1.0 else loss
    loss.backward()
    return float(loss.detach()) / grad_output, float(acc), a.grad.numpy(), b.grad.numpy()


NTXENT_CASES = [
    # name, B, d, tau, normalize, kind, noise, weight, grad_output, bf16
    ("b1_d128", 1, 128, 0.5, True, "iid", 0.5, False, 1.0, False),
    ("b2_d128", 2, 128, 0.5, True, "iid", 0.5, False, 1.0, False),
    ("b5_d64_ragged", 5, 64, 0.5, True, "iid", 0.5, False, 1.0, False),
    ("b16_d128_tau1", 16, 128, 1.0, True, "iid", 0.5, False, 1.0, False),
    ("b64_d128", 64, 128, 0.5, True, "iid", 0.5, False, 1.0, False),
    ("b64_d128_corr", 64, 128, 0.5, True, "correlated", 0.5, False, 1.0, False),
    ("b64_d128_corr_tau01", 64, 128, 0.1, True, "correlated", 0.1, False, 1.0, False),
    ("b100_d128_ragged", 100, 128, 0.5, True, "correlated", 0.5, False, 1.0, False),
    ("b64_d128_weight", 64, 128, 0.5, True, "iid", 0.5, True, 1.0, False),
    ("b64_d128_nonorm", 64, 128, 0.5, False, "iid", 0.5, False, 1.0, False),
    ("b64_d128_accum8", 64, 128, 0.5, True, "correlated", 0.5, False, 0.125, False),
    ("b40_d256", 40, 256, 0.1, True, "correlated", 0.5, False, 1.0, False),
    ("b200_d128_bf16", 200, 128, 0.5, True, "correlated", 0.5, False, 1.0, True),
    ("b512_d128_cfg1", 512, 128, 0.5, True, "iid", 0.5, False, 1.0, False),
    ("b512_d128_cfg1_corr", 512, 128, 0.5, True, "correlated", 0.5, False, 1.0, False),
]

MODIFIED_CASES = [
    # name, B, d, tau, kind, noise, grad_output, bf16
    ("b1_d128", 1, 128, 1.0, "iid", 0.5, 1.0, False),
    ("b5_d64_ragged", 5, 64, 0.5, "iid", 0.5, 1.0, False),
    ("b64_d128_default_tau", 64, 128, None, "iid", 0.5, 1.0, False),
    ("b64_d128_tau05", 64, 128, 0.5, "correlated", 0.5, 1.0, False),
    ("b64_d128_tau01", 64, 128, 0.1, "correlated", 0.5, 1.0, False),
    ("b100_d128_ragged", 100, 128, 0.5, "correlated", 0.5, 0.125, False),
    ("b200_d128_bf16", 200, 128, 0.5, "correlated", 0.5, 1.0, True),
    ("b512_d128", 512, 128, 0.5, "iid", 0.5, 1.0, False),
]


def main():
    ref = _load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order -> reproducible fixtures
    for i, (name, n, d, tau, norm, kind, noise, use_w, go, bf16) in enumerate(NTXENT_CASES):
        z1, z2 = make_embeddings(n, d, seed=100 + i, kind=kind, noise=noise, bf16_representable=bf16)
        kw = dict(temperature=tau, normalize=norm)
        w = None
        if use_w:
            w = torch.rand(2 * n, generator=torch.Generator().manual_seed(7)) + 0.25
            kw["weight"] = w
        loss, acc, g1, g2 = _run(ref.contrastive_loss, z1, z2, go, **kw)
        np.savez_compressed(os.path.join(OUT, f"ntxent_{name}.npz"), z1=z1.numpy(), z2=z2.numpy(),
                            temperature=tau, normalize=norm, weight=(w.numpy() if w is not None else np.zeros(0)),
                            grad_output=go, loss=loss, acc=acc, grad1=g1, grad2=g2, bf16=bf16)
        print(f"ntxent_{name}: loss={loss:.6f} acc={acc:.3f}")
    # exact-tie fixture: duplicated rows exercise the first-argmax rule (objective.py:51)
    z = torch.randn(4, 32, generator=torch.Generator().manual_seed(3))
    z1 = z[[0, 0, 1, 2]].clone()
    z2 = z[[0, 0, 1, 3]].clone()
    loss, acc, g1, g2 = _run(ref.contrastive_loss, z1, z2, 1.0, temperature=0.5, normalize=True)
    np.savez_compressed(os.path.join(OUT, "ntxent_ties_b4_d32.npz"), z1=z1.numpy(), z2=z2.numpy(), temperature=0.5,
                        normalize=True, weight=np.zeros(0), grad_output=1.0, loss=loss, acc=acc, grad1=g1, grad2=g2,
                        bf16=False)
    print(f"ntxent_ties: loss={loss:.6f} acc={acc:.3f}")
    # exact ties INSIDE the positive's 128-column tile, earlier and later than the positive (two image blocks, so that the
    # sm_100a kernels see the tie in an ordinary chunk of a masked tile): view-2 row 130 duplicates row 200 (the
    # duplicate precedes the positive of row 200 -> the reference's first-argmax picks it), row 250 duplicates row 140
    # (the duplicate follows the positive of row 140 -> the positive wins)
    z1, z2 = make_embeddings(256, 128, seed=77, kind="correlated", noise=0.05)
    z2 = z2.clone()
    z2[130] = z2[200]
    z2[250] = z2[140]
    loss, acc, g1, g2 = _run(ref.contrastive_loss, z1, z2, 1.0, temperature=0.5, normalize=True)
    np.savez_compressed(os.path.join(OUT, "ntxent_ties_b256_d128_same_tile.npz"), z1=z1.numpy(), z2=z2.numpy(),
                        temperature=0.5, normalize=True, weight=np.zeros(0), grad_output=1.0, loss=loss, acc=acc, grad1=g1,
                        grad2=g2, bf16=False)
    print(f"ntxent_ties_same_tile: loss={loss:.6f} acc={acc:.3f}")

    for i, (name, n, d, tau, kind, noise, go, bf16) in enumerate(MODIFIED_CASES):
        z1, z2 = make_embeddings(n, d, seed=300 + i, kind=kind, noise=noise, bf16_representable=bf16)
        kw = {} if tau is None else dict(temperature=tau)
        loss, acc, g1, g2 = _run(ref.modified_contrastive_loss, z1, z2, go, **kw)
        np.savez_compressed(os.path.join(OUT, f"modified_{name}.npz"), z1=z1.numpy(), z2=z2.numpy(),
                            temperature=(1.0 if tau is None else tau), default_tau=(tau is None), grad_output=go,
                            loss=loss, acc=acc, grad1=g1, grad2=g2, bf16=bf16)
        print(f"modified_{name}: loss={loss:.6f} acc={acc:.3f}")


if __name__ == "__main__":
    main()
