"""GPU bring-up diagnostics (run on the B200 box through gpurun).  Each stage runs in its own process so a
trapped kernel cannot poison the next stage.  Everything is logged to gpurun_out/first_light.log.

    python tests/first_light.py            # all stages
    python tests/first_light.py STAGE ...  # selected stages
"""
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))


def stage_selftest():
    import numpy as np
    import torch
    from pytorch_simclr_b200 import _lib
    lib = _lib.load_debug()
    torch.manual_seed(0)
    a = torch.randn(128, 128).to(torch.bfloat16).cuda()
    b = torch.randn(128, 128).to(torch.bfloat16).cuda()
    out = torch.full((3, 128, 128), float("nan"), device="cuda")
    rc = lib.simclr_selftest_umma(a.data_ptr(), b.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    print("rc", rc)
    torch.cuda.synchronize()
    af, bf = a.float().cpu().double().numpy(), b.float().cpu().double().numpy()
    ref = [af @ bf.T, af @ bf, af @ bf]
    o = out.cpu().double().numpy()
    for i, name in enumerate(["A*B^T (SS, K/K)", "A*B (SS, K/MN)", "A*B (TS, tmem A)"]):
        err = np.abs(o[i] - ref[i])
        print(f"{name}: max abs err {np.nanmax(err):.3e} nan={np.isnan(o[i]).sum()} ref scale {np.abs(ref[i]).max():.2f}")
        if not (np.nanmax(err) < 1e-2):
            bad = np.argwhere(~(err < 1e-2))
            print("   first bad entries", bad[:8].tolist(), "values", [float(o[i][tuple(x)]) for x in bad[:4]],
                  "expected", [float(ref[i][tuple(x)]) for x in bad[:4]])
            # does it match a transposed / permuted reference?
            for alt_name, alt in (("ref^T", ref[i].T), ("A^T B", af.T @ bf), ("A B^T", af @ bf.T), ("A^T B^T", af.T @ bf.T)):
                print(f"   vs {alt_name}: {np.nanmax(np.abs(o[i] - alt)):.3e}")


def _compare(loss_kind, b, d, tau, kind="iid", normalize=True, dtype="f32", weight=False, grad_out=1.0, seed=0):
    import numpy as np
    import torch
    import contrastive_oracle as oracle
    import pytorch_simclr_b200 as sb
    z1, z2 = oracle.make_embeddings(b, d, seed=seed, kind=kind, bf16_representable=(dtype == "bf16"))
    w = None
    if weight:
        w = torch.rand(2 * b, generator=torch.Generator().manual_seed(1)) + 0.25
    if loss_kind == 0:
        ref = oracle.ntxent_closed_form(z1, z2, temperature=tau, normalize=normalize, weight=w, grad_output=grad_out)
    else:
        ref = oracle.modified_closed_form(z1, z2, temperature=tau, grad_output=grad_out)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    x1 = z1.to(tdt).cuda().requires_grad_(True)
    x2 = z2.to(tdt).cuda().requires_grad_(True)
    t0 = time.time()
    if loss_kind == 0:
        loss, acc = sb.contrastive_loss(x1, x2, temperature=tau, normalize=normalize, weight=None if w is None else w.cuda())
    else:
        loss, acc = sb.modified_contrastive_loss(x1, x2, temperature=tau)
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    dt = time.time() - t0
    g1 = x1.grad.float().cpu().double().numpy()
    g2 = x2.grad.float().cpu().double().numpy()
    gmax = max(np.abs(ref.grad1).max(), np.abs(ref.grad2).max())
    e1 = np.abs(g1 - ref.grad1).max() / gmax
    e2 = np.abs(g2 - ref.grad2).max() / gmax
    lrel = abs(float(loss) - ref.loss) / max(abs(ref.loss), 1e-30)
    print(f"loss_kind={loss_kind} B={b} d={d} tau={tau} {kind} norm={normalize} {dtype} w={weight} go={grad_out}: "
          f"loss {float(loss):.6f} ref {ref.loss:.6f} rel {lrel:.2e} | acc {acc:.4f} ref {ref.acc:.4f} | "
          f"grad rel-to-max err {e1:.2e} {e2:.2e} | nan {np.isnan(g1).sum() + np.isnan(g2).sum()} | {dt*1e3:.1f} ms")
    return lrel, e1, e2


def stage_ntxent_small():
    _compare(0, 64, 128, 0.5)
    _compare(0, 64, 128, 0.5, kind="correlated")
    _compare(0, 1, 128, 0.5)
    _compare(0, 5, 64, 0.5)
    _compare(0, 100, 128, 0.5, kind="correlated")
    _compare(0, 64, 128, 0.5, weight=True)
    _compare(0, 64, 128, 0.5, normalize=False)
    _compare(0, 64, 128, 0.5, grad_out=0.125)
    _compare(0, 40, 256, 0.1, kind="correlated")
    _compare(0, 64, 100, 0.5)


def stage_ntxent_medium():
    _compare(0, 512, 128, 0.5)
    _compare(0, 512, 128, 0.5, kind="correlated", dtype="bf16")
    _compare(0, 700, 128, 0.1, kind="correlated")
    _compare(0, 1000, 256, 0.5)
    _compare(0, 2048, 64, 0.5)


def stage_ntxent_large():
    _compare(0, 4096, 128, 0.5)
    _compare(0, 4096, 128, 0.5, kind="correlated", dtype="bf16")
    _compare(0, 4096, 128, 0.1, kind="correlated")


def stage_modified():
    _compare(1, 64, 128, 0.5)
    _compare(1, 64, 128, 1.0)
    _compare(1, 64, 128, 0.1, kind="correlated")
    _compare(1, 5, 64, 0.5)
    _compare(1, 100, 128, 0.5, kind="correlated", grad_out=0.125)
    _compare(1, 512, 128, 0.5)
    _compare(1, 4096, 128, 0.5, dtype="bf16")
    _compare(1, 4096, 128, 0.1, kind="correlated", dtype="bf16")


def stage_timing():
    import torch
    import contrastive_oracle as oracle
    from pytorch_simclr_b200 import functional as F
    for loss_kind in (0, 1):
        for b, d in ((512, 128), (4096, 128), (4096, 256), (16384, 128)):
            z1, z2 = oracle.make_embeddings(b, d, seed=0)
            x1, x2 = z1.cuda(), z2.cuda()
            for _ in range(3):
                F.contrastive_forward_backward(loss_kind, x1, x2, 0.5)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            n = 20
            ev[0].record()
            for _ in range(n):
                F.contrastive_forward_backward(loss_kind, x1, x2, 0.5)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / n
            m = 2 * b
            flops = (6 if loss_kind == 0 else 3) * m * m * d
            print(f"loss_kind={loss_kind} 2N={m} d={d}: {ms*1e3:.1f} us fwd+bwd, {m/ms*1e3:.3e} views/s, "
                  f"{flops/ms*1e-9:.1f} TFLOP/s algorithmic")


STAGES = {k[len("stage_"):]: v for k, v in list(globals().items()) if k.startswith("stage_")}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(REPO, "gpurun_out", "first_light.log"), "a")
    for name in names:
        hdr = f"===== stage {name} ====="
        print(hdr, flush=True)
        log.write(hdr + "\n")
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], capture_output=True,
                               text=True, timeout=240)
            out = p.stdout + p.stderr[-4000:] + f"\n[exit {p.returncode}]\n"
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            out += "\n[TIMEOUT]\n"
        print(out, flush=True)
        log.write(out)
        log.flush()
