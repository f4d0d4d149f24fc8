"""Multi-process host logic on CPU: world_size-2 gloo runs of the exchange steps in
snekmer_b200/dist.py (sharding, basis-table merge, learn sum, top-2 gather), checked against
the single-process oracle.  The compute that surrounds them is CUDA-only; here the per-rank
inputs come from the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import skm_oracle as O
from snekmer_b200 import dist as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _seqs(n=240, seed=5):
    rng = np.random.default_rng(seed)
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWYX"))
    return ["".join(rng.choice(aa, size=int(rng.integers(0, 90)))) for _ in range(n)]


def test_shard_bounds_balanced_and_covering():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 500, size=1000)
    off = np.concatenate([[0], np.cumsum(lens)])
    for parts in (1, 2, 3, 8):
        b = D.shard_bounds(off, parts)
        assert len(b) == parts and b[0][0] == 0 and b[-1][1] == 1000
        assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
        res = [off[hi] - off[lo] for lo, hi in b]
        assert max(res) - min(res) <= 2 * lens.max()
    assert D.shard_bounds(np.array([0]), 4) == [(0, 0)] * 4
    assert D.shard_bounds(np.array([0, 0, 0]), 2)[-1][1] == 2


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    D.init("gloo")
    try:
        a, k, mf = 5, 3, 1
        seqs = _seqs()
        lut, syms = O.build_lut(a)
        res, off = O.pack(seqs)
        lo, hi = D.shard_bounds(off, world)[rank]
        S = len(syms) ** k
        # ---- basis: per-rank tables with GLOBAL first positions, then the exchange -------------
        si, pos, code, valid = O.window_codes(res[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], lut, len(syms), k)
        base, total = D.exclusive_prefix(int(off[hi] - off[lo]))
        assert base == int(off[lo]) and total == int(off[-1])
        count = torch.zeros(S, dtype=torch.int64)
        first = torch.full((S,), -1, dtype=torch.int64)
        c = code[valid].astype(np.int64)
        g = (off[lo:hi + 1] - off[lo])[si[valid]] + pos[valid] + base
        np.add.at(count.numpy(), c, 1)
        f = np.full(S, np.iinfo(np.int64).max)
        np.minimum.at(f, c, g)
        first.numpy()[f != np.iinfo(np.int64).max] = f[f != np.iinfo(np.int64).max]
        D.allreduce_basis_tables(count, first)
        # single-process oracle
        si0, pos0, code0, valid0 = O.window_codes(res, off, lut, len(syms), k)
        want_basis, want_tot = O.basis_codes(si0, pos0, code0, valid0, mf)
        cnt, fst = count.numpy(), first.numpy()
        keep = np.flatnonzero(cnt > mf)
        got_basis = keep[np.argsort(fst[keep], kind="stable")].astype(np.uint64)
        assert np.array_equal(got_basis, want_basis) and np.array_equal(cnt[got_basis.astype(np.int64)], want_tot)
        # ---- learn: per-rank annotation sums, then the sum over ranks --------------------------
        n_ann = 6
        ann = np.random.default_rng(1).integers(-1, n_ann, size=len(seqs))
        C = O.count_matrix(si, code, valid, hi - lo, want_basis).astype(np.int64)
        M = torch.zeros((n_ann + 1, len(want_basis)), dtype=torch.int64)
        np.add.at(M.numpy(), np.where(ann[lo:hi] < 0, n_ann, ann[lo:hi]), C)
        totals = torch.from_numpy(C.sum(axis=0))
        D.allreduce_sum_(M, totals)
        C0 = O.count_matrix(si0, code0, valid0, len(seqs), want_basis).astype(np.int64)
        want_M = np.zeros((n_ann + 1, len(want_basis)), dtype=np.int64)
        np.add.at(want_M, np.where(ann < 0, n_ann, ann), C0)
        assert np.array_equal(M.numpy(), want_M) and np.array_equal(totals.numpy(), C0.sum(axis=0))
        # ---- apply, annotation-sharded: gather per-shard top-2, merge == global top-2 -----------
        Sfull = O.cosine_scores(C0[:50], want_M[:n_ann])
        rows = D.shard_bounds(np.arange(n_ann + 1), world)[rank]
        part = Sfull[:, rows[0]:rows[1]]
        i1, i2, s1, s2 = O.top2(part) if part.shape[1] else (np.full(50, -1), np.full(50, -1), np.zeros(50), np.full(50, np.nan))
        idx, sc = D.allgather_top2(torch.from_numpy(i1.astype(np.int32)), torch.from_numpy(i2.astype(np.int32)),
                                   torch.from_numpy(s1), torch.from_numpy(s2), rows[0])
        assert tuple(idx.shape) == (world, 2, 50)
        cand_i = idx.permute(2, 0, 1).reshape(50, -1).numpy()
        cand_s = sc.permute(2, 0, 1).reshape(50, -1).numpy()
        w1, w2, ws1, ws2 = O.top2(Sfull)
        for q in range(50):
            ok = cand_i[q] >= 0
            order = sorted(zip(-cand_s[q][ok], cand_i[q][ok]))
            assert order[0][1] == w1[q] and order[1][1] == w2[q]
        # ---- sparse learn fan-in: all_to_all of sorted COO runs by annotation range ------------------
        Sp = 1000
        sel = ann[lo:hi] >= 0
        kk = (ann[lo:hi][sel].astype(np.int64)[:, None] * Sp + np.arange(3)[None, :] * 7 + (np.arange(sel.sum()) % 5)[:, None]).reshape(-1)
        uk, cnt = np.unique(kk, return_counts=True)
        bounds = [(n_ann * r // world) * Sp for r in range(world)] + [n_ann * Sp]
        rk, rv = D.alltoall_coo_by_key_range(torch.from_numpy(uk), torch.from_numpy(cnt.astype(np.int64)), bounds)
        assert rk.numel() == rv.numel() and (rk >= bounds[rank]).all() and (rk < bounds[rank + 1]).all()
        tot = torch.tensor([int(rv.sum()), int(cnt.sum())])
        dist.all_reduce(tot)
        assert tot[0].item() == tot[1].item()                 # nothing lost, nothing duplicated
        # ---- layout of the peer-memory exchange (push_plan) == the layout all_to_all_single produced -------------------
        cutk = np.searchsorted(uk, bounds)
        send = torch.from_numpy(np.diff(cutk).astype(np.int64))
        mat = torch.empty(world * world, dtype=torch.int64)
        dist.all_gather_into_tensor(mat, send)
        dst_off, runs, words = D.push_plan(mat.numpy().reshape(world, world), rank)
        assert int(runs.sum()) == rk.numel() and words >= rk.numel()
        plans = [None] * world
        dist.all_gather_object(plans, (dst_off.tolist(), send.tolist(), uk[cutk[0]:cutk[-1]].tolist()))
        buf = np.full(int(runs.sum()), -1, dtype=np.int64)              # what the pushes of all ranks would leave in MY buffer
        for s_, (offs_, send_, keys_) in enumerate(plans):
            start = int(np.sum(send_[:rank]))
            buf[offs_[rank]:offs_[rank] + send_[rank]] = keys_[start:start + send_[rank]]
        assert np.array_equal(buf, rk.numpy())
        # ---- peer buffers that cannot be set up: every rank learns it in the same collective and the path switches off ---
        os.environ["SKM_PEER_FAIL"] = "1"
        pb = D.PeerBuffers()
        assert pb.ensure(1000) is False and pb.disabled and pb.own is None and pb.ensure(10) is False
        del os.environ["SKM_PEER_FAIL"]
        # ---- the same exchange with ranges balanced by entry count (Zipf-sized annotations) ------------------------
        zrng = np.random.default_rng(7 + rank)
        wz = 1.0 / np.arange(1, 41) ** 1.3
        za = zrng.choice(40, size=3000, p=wz / wz.sum()).astype(np.int64)
        zk = np.unique(za * Sp + zrng.integers(0, Sp, size=3000))
        zb = D.balanced_annotation_bounds(torch.from_numpy(zk), Sp, 40)
        gathered = [None] * world
        dist.all_gather_object(gathered, (zb, zk))
        assert all(g[0] == zb for g in gathered) and zb[0] == 0 and zb[-1] == 40 and zb == sorted(zb)
        allk = np.concatenate([g[1] for g in gathered])
        per_range = [int(((allk // Sp >= zb[r]) & (allk // Sp < zb[r + 1])).sum()) for r in range(world)]
        biggest = np.bincount(allk // Sp, minlength=40).max()
        assert max(per_range) <= len(allk) / world + biggest, (per_range, biggest)       # within one annotation of even
        even = [int(((allk // Sp >= 40 * r // world) & (allk // Sp < 40 * (r + 1) // world)).sum()) for r in range(world)]
        assert max(per_range) <= max(even)
        # ---- wide basis fan-in: all_gather of per-rank (code, count, first) tables, merged by code -------------
        lc, li, lcnt = np.unique(code[valid], return_index=True, return_counts=True)
        lfirst = ((off[lo:hi + 1] - off[lo])[si[valid]] + pos[valid] + base)[li]
        gc, gn, gf = D.allgather_tables(torch.from_numpy(lc.astype(np.int64)), torch.from_numpy(lcnt.astype(np.int64)),
                                        torch.from_numpy(lfirst.astype(np.int64)))
        assert gc.numel() == gn.numel() == gf.numel()
        mc, inv = np.unique(gc.numpy(), return_inverse=True)
        msum = np.zeros(len(mc), np.int64)
        np.add.at(msum, inv, gn.numpy())
        mfirst = np.full(len(mc), np.iinfo(np.int64).max)
        np.minimum.at(mfirst, inv, gf.numpy())
        keep2 = msum > mf
        wide_basis = mc[keep2][np.argsort(mfirst[keep2], kind="stable")].astype(np.uint64)
        assert np.array_equal(wide_basis, want_basis)
        # ---- query-sharded outputs concatenated in rank order ---------------------------------
        mine = torch.arange(lo, hi, dtype=torch.int64).reshape(-1, 1)
        allrows = D.gather_rows(mine, dst=0)
        if rank == 0:
            assert allrows.reshape(-1).tolist() == list(range(len(seqs)))
        with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_exchange_steps(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
