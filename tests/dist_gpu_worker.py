"""Runs under torchrun (one rank per GPU, NCCL): the three exchange steps of the path on real
shards, checked against the single-process oracle.  Launched by tests/test_dist_gpu.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import skm_oracle as O  # noqa: E402
from snekmer_b200 import dist as D  # noqa: E402
from snekmer_b200 import engine as E  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank, world = D.init("nccl", torch.device("cuda", local))
    rng = np.random.default_rng(11)
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWYX"))
    seqs = ["".join(rng.choice(aa, size=int(rng.integers(0, 400)))) for _ in range(3000)]
    a, k, mf, n_ann = 2, 6, 1, 23
    ann = rng.integers(-1, n_ann, size=len(seqs)).astype(np.int32)
    res, off = O.pack(seqs)
    lo, hi = D.shard_bounds(off, world)[rank]
    batch = E.SequenceBatch.from_strings(seqs[lo:hi])
    # basis of the concatenation, identical on all ranks
    basis = E.build_basis_distributed(batch, a, k, mf)
    lut, syms = O.build_lut(a)
    si, pos, code, valid = O.window_codes(res, off, lut, len(syms), k)
    want_basis, want_tot = O.basis_codes(si, pos, code, valid, mf)
    assert np.array_equal(basis.codes_host(), want_basis), "distributed basis"
    assert np.array_equal(basis.counts.cpu().numpy(), want_tot), "distributed basis counts"
    # the same basis through the sort-based wide path (tables all-gathered, merged by one sort), and a code
    # space beyond the table limit
    wb = E.build_basis_wide_distributed(batch, a, k, mf)
    assert np.array_equal(wb.codes_host(), want_basis) and np.array_equal(wb.counts.cpu().numpy(), want_tot), "wide distributed basis"
    si2, pos2, code2, valid2 = O.window_codes(res, off, O.build_lut(None)[0], 20, 9)
    wb2 = E.build_basis_wide_distributed(batch, None, 9, 0)
    assert np.array_equal(wb2.codes_host(), O.basis_codes(si2, pos2, code2, valid2, 0)[0]), "wide distributed basis 20^9"
    # learn + NCCL sum
    M, totals = E.learn_dense_distributed(batch, a, k, basis, torch.from_numpy(ann[lo:hi]), n_ann)
    C = O.count_matrix(si, code, valid, len(seqs), want_basis).astype(np.int64)
    want_M = np.zeros((n_ann + 1, len(want_basis)), dtype=np.int64)
    np.add.at(want_M, np.where(ann < 0, n_ann, ann), C)
    assert np.array_equal(M.cpu().numpy(), want_M), "distributed learn"
    assert np.array_equal(totals.cpu().numpy(), C.sum(axis=0)), "distributed totals"
    # sparse learn: grouped 32-bit-key build on the local shard + all_to_all of COO runs by annotation range
    S = len(syms) ** k
    keys, vals = E.learn_sparse(batch, a, k, torch.from_numpy(ann[lo:hi]), n_ann)
    keys0, vals0 = keys, vals
    keys, vals, (a_lo, a_hi) = E.exchange_coo_by_annotation(keys0, vals0, S, n_ann, balance=False)
    assert (a_lo, a_hi) == (n_ann * rank // world, n_ann * (rank + 1) // world)
    kb, vb, (b_lo, b_hi) = E.exchange_coo_by_annotation(keys0, vals0, S, n_ann)          # ranges balanced by entry count
    spans = [None] * world
    torch.distributed.all_gather_object(spans, (b_lo, b_hi, int(kb.numel())))
    assert spans[0][0] == 0 and spans[-1][1] == n_ann and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1)), spans
    kbn = kb.cpu().numpy()
    gotb = np.zeros((n_ann, S), dtype=np.int64)
    gotb[kbn // S, kbn % S] = vb.cpu().numpy()
    assert ((kbn // S >= b_lo) & (kbn // S < b_hi)).all(), "balanced exchange ranges"
    full = np.zeros((n_ann, S), dtype=np.int64)
    full[:, want_basis.astype(np.int64)] = want_M[:n_ann]
    kk = keys.cpu().numpy()
    got = np.zeros((n_ann, S), dtype=np.int64)
    got[kk // S, kk % S] = vals.cpu().numpy()
    assert np.all(kk[1:] > kk[:-1]) and (kk // S >= a_lo).all() and (kk // S < a_hi).all(), "sparse learn exchange ranges"
    cols = want_basis.astype(np.int64)         # the dense matrix only holds basis columns (count > min_filter)
    assert np.array_equal(got[a_lo:a_hi][:, cols], want_M[a_lo:a_hi]), "sparse learn exchange"
    assert np.array_equal(gotb[b_lo:b_hi][:, cols], want_M[b_lo:b_hi]), "balanced sparse learn exchange"
    # the default exchange above went through NVLink peer memory (skm_coo_pack_push); the NCCL all_to_all baseline must
    # give the same lists bit for bit, and so must a run whose counts do not fit the packed word (falls back inside)
    assert E.peer_exchange_enabled(keys0), "peer exchange should be the default under NCCL"
    os.environ["SKM_EXCHANGE"] = "nccl"
    kn, vn, rn = E.exchange_coo_by_annotation(keys0, vals0, S, n_ann)
    os.environ["SKM_EXCHANGE"] = "peer"
    assert rn == (b_lo, b_hi) and torch.equal(kn, kb) and torch.equal(vn, vb), "peer exchange != NCCL exchange"
    big = vals0.clone()
    if big.numel():
        big[0] = 1 << 61
    ko, vo, _ = E.exchange_coo_by_annotation(keys0, big, S, n_ann)
    os.environ["SKM_EXCHANGE"] = "nccl"
    ko2, vo2, _ = E.exchange_coo_by_annotation(keys0, big, S, n_ann)
    os.environ["SKM_EXCHANGE"] = "peer"
    assert torch.equal(ko, ko2) and torch.equal(vo, vo2), "overflow fallback of the peer exchange"
    for _ in range(3):                                    # buffer reuse across steps (fences)
        k4, v4, _ = E.exchange_coo_by_annotation(keys0, vals0, S, n_ann)
        assert torch.equal(k4, kb) and torch.equal(v4, vb), "repeated peer exchange"
    # a node where the peer buffers cannot be set up (no peer access, CUDA IPC closed off): every rank notices in the same
    # collective, the path is disabled for the process and the NCCL exchange takes over with the same result
    D.peer_buffers().release()
    os.environ["SKM_PEER_FAIL"] = "1"
    k5, v5, _ = E.exchange_coo_by_annotation(keys0, vals0, S, n_ann)
    del os.environ["SKM_PEER_FAIL"]
    assert D.peer_buffers().disabled and not E.peer_exchange_enabled(keys0)
    assert torch.equal(k5, kb) and torch.equal(v5, vb), "fallback when the peer buffers cannot be set up"
    if rank == 0:
        sizes = [sp[2] for sp in spans]
        print("balanced exchange entries per rank:", sizes)
    # apply, annotation-sharded: all queries everywhere, my slice of annotation rows
    qbatch = E.SequenceBatch.from_strings(seqs[:500])
    Q = E.count_dense(qbatch, a, k, basis)
    rows = D.shard_bounds(np.arange(n_ann + 1), world)[rank]
    r = E.apply_dense_annotation_sharded(Q, M[rows[0]:rows[1]].contiguous(), rows[0])
    S = O.cosine_scores(C[:500], want_M[:n_ann])
    i1, i2, s1, s2 = O.top2(S)
    assert np.array_equal(r.top1.cpu().numpy(), i1) and np.array_equal(r.top2.cpu().numpy(), i2), "sharded top-2"
    assert np.max(np.abs(r.score1.cpu().numpy() - s1)) < 1e-12 and np.max(np.abs(r.score2.cpu().numpy() - s2)) < 1e-12
    # query-sharded apply with replicated M: concatenate on rank 0
    rq = E.apply_dense(E.count_dense(batch, a, k, basis), M[:n_ann].contiguous())
    allq = D.gather_rows(rq.top1.reshape(-1, 1), dst=0)
    if rank == 0:
        Sall = O.cosine_scores(C, want_M[:n_ann])
        assert np.array_equal(allq.reshape(-1).cpu().numpy(), O.top2(Sall)[0]), "query-sharded apply"
        print(f"dist_gpu_worker ok: world={world} K={basis.K}")
    D.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
