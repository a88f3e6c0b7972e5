"""world_size-2 gloo tests (CPU) of the row-sharded global-batch formulation and its collective plumbing.

The product kernels only run on a B200, so here the per-rank kernel is emulated in fp64 numpy on exactly
the buffers `RowShardGather` produces (view-padded layouts), and the result is compared with the
single-process oracle on the gathered batch.  What this pins down: the gather layout, row offsets, the
stats reduction, the weighted column scale, and that the symmetric backward form needs no reduce-scatter.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import contrastive_oracle as oracle
from pytorch_simclr_b200 import functional as F
from pytorch_simclr_b200.distributed import RowShardGather, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _emulate_rank(op_rows, op_cols, pos_dot, b_loc, b_glob, row_off, tau, w_rows=None):
    """fp64 emulation of simclr_forward on view-padded buffers: returns lse (natural log) per local slot,
    and the local stats [sum w L, sum w, correct]."""
    bl_pad, bg_pad = op_rows.shape[0] // 2, op_cols.shape[0] // 2
    lse = np.zeros(2 * bl_pad)
    stats = np.zeros(3)
    col_valid = np.zeros(2 * bg_pad, dtype=bool)
    col_valid[:b_glob] = True
    col_valid[bg_pad:bg_pad + b_glob] = True
    for v in (0, 1):
        for i in range(b_loc):
            slot = v * bl_pad + i
            g = row_off + i
            s = op_cols @ op_rows[slot] / tau
            mask = col_valid.copy()
            mask[v * bg_pad + g] = False
            pos = (1 - v) * bg_pad + g
            mask[pos] = False
            s_pos = pos_dot[slot] / tau
            terms = np.concatenate((s[mask], [s_pos]))
            m = terms.max()
            lse[slot] = m + np.log(np.exp(terms - m).sum())
            w = 1.0 if w_rows is None else w_rows[v * b_loc + i]
            stats[0] += w * (lse[slot] - s_pos)
            stats[1] += w
            # reference column order [view-2 block | view-1 block], first maximal index wins
            order = np.concatenate((np.arange(bg_pad, bg_pad + b_glob), np.arange(0, b_glob)))
            sc = s[order].copy()
            sc[np.where(order == v * bg_pad + g)[0]] = -np.inf
            stats[2] += float(order[int(np.argmax(sc))] == pos)
    return lse, stats


def _emulate_backward(op_rows, op_cols, lse_cols, col_scale, b_loc, b_glob, row_off, tau):
    """Symmetric-form row gradient d loss / d zhat for the local rows (view-padded [2*bl_pad, d])."""
    bl_pad, bg_pad = op_rows.shape[0] // 2, op_cols.shape[0] // 2
    out = np.zeros_like(op_rows)
    for v in (0, 1):
        for i in range(b_loc):
            slot = v * bl_pad + i
            g = row_off + i
            me = v * bg_pad + g
            pos = (1 - v) * bg_pad + g
            s = op_cols @ op_rows[slot] / tau
            w = col_scale[me] * np.exp(s - lse_cols[me]) + col_scale * np.exp(s - lse_cols)
            w[col_scale == 0] = 0.0
            w[me] = 0.0
            w[pos] -= col_scale[me] + col_scale[pos]
            out[slot] = (w @ op_cols) / tau
    return out


def _worker(rank, world, port, b_glob, d, tau, use_weight, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z1, z2 = oracle.make_embeddings(b_glob, d, seed=3, kind="correlated")
        wfull = None
        if use_weight:
            wfull = torch.rand(2 * b_glob, generator=torch.Generator().manual_seed(5), dtype=torch.float64) + 0.25
        row_off, b_loc = shard_rows(b_glob, world, rank)
        x1 = z1[row_off:row_off + b_loc].double()
        x2 = z2[row_off:row_off + b_loc].double()
        # what simclr_prepare would write (fp64 here): normalised operands in the local view-padded layout
        bl_pad = F.pad_rows(b_loc)
        op = torch.zeros(2 * bl_pad, d, dtype=torch.float64)
        op[:b_loc] = x1 / x1.norm(dim=1, keepdim=True)
        op[bl_pad:bl_pad + b_loc] = x2 / x2.norm(dim=1, keepdim=True)
        pos_dot = torch.zeros(2 * bl_pad, dtype=torch.float64)
        pos_dot[:b_loc] = (op[:b_loc] * op[bl_pad:bl_pad + b_loc]).sum(1)
        pos_dot[bl_pad:bl_pad + b_loc] = pos_dot[:b_loc]

        gather = RowShardGather()
        op_cols, bg, off = gather.operand(op, b_loc)
        assert (bg, off) == (b_glob, row_off)
        bg_pad = F.pad_rows(b_glob)
        assert op_cols.shape == (2 * bg_pad, d)
        # layout: view-major, ranks in order inside a view, zero padding
        full1 = z1.double() / z1.double().norm(dim=1, keepdim=True)
        full2 = z2.double() / z2.double().norm(dim=1, keepdim=True)
        assert torch.allclose(op_cols[:b_glob], full1, atol=1e-15)
        assert torch.allclose(op_cols[bg_pad:bg_pad + b_glob], full2, atol=1e-15)
        assert op_cols[b_glob:bg_pad].abs().sum() == 0 and op_cols[bg_pad + b_glob:].abs().sum() == 0

        w_loc = None
        if use_weight:
            w_loc = torch.cat((wfull[row_off:row_off + b_loc], wfull[b_glob + row_off:b_glob + row_off + b_loc]))
        lse, stats = _emulate_rank(op.numpy(), op_cols.numpy(), pos_dot.numpy(), b_loc, b_glob, row_off, tau,
                                   None if w_loc is None else w_loc.numpy())
        g_loss, g_stats = gather.reduce(torch.tensor(np.concatenate((stats, [stats[0] / stats[1]]))), None)
        assert not g_loss._is_view()
        lse_cols = gather.rowvec(torch.from_numpy(lse), b_loc)
        if use_weight:
            col_scale = gather.col_scale(w_loc, g_stats, b_loc)
        else:
            col_scale = torch.zeros(2 * bg_pad, dtype=torch.float64)
            col_scale[:b_glob] = 0.5 / b_glob
            col_scale[bg_pad:bg_pad + b_glob] = 0.5 / b_glob
        dzh = _emulate_backward(op.numpy(), op_cols.numpy(), lse_cols.numpy(), col_scale.numpy(), b_loc, b_glob,
                                row_off, tau)

        ref = oracle.ntxent_closed_form(z1, z2, temperature=tau, weight=None if wfull is None else wfull.numpy())
        assert float(g_loss) == pytest.approx(ref.loss, rel=1e-12)
        assert int(round(float(g_stats[2]))) == ref.correct
        # back through the normalisation, compare with the oracle's rows of this rank
        for v, (x, refg) in enumerate(((x1, ref.grad1), (x2, ref.grad2))):
            zh = op[v * bl_pad: v * bl_pad + b_loc].numpy()
            dz_hat = dzh[v * bl_pad: v * bl_pad + b_loc]
            nrm = x.norm(dim=1, keepdim=True).numpy()
            dz = (dz_hat - zh * (zh * dz_hat).sum(1, keepdims=True)) / nrm
            assert np.abs(dz - refg[row_off:row_off + b_loc]).max() < 1e-13 * max(1.0, np.abs(refg).max() * 1e13) * 1e-0 + 1e-14
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("b_glob,d,tau,use_weight", [(12, 16, 0.5, False), (150, 32, 0.1, False), (20, 8, 0.5, True)])
def test_row_sharded_global_batch_world2(tmp_path, b_glob, d, tau, use_weight):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, b_glob, d, tau, use_weight, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_rows():
    assert shard_rows(512, 8, 3) == (192, 64)
    with pytest.raises(ValueError):
        shard_rows(10, 4, 0)
