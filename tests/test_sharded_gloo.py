"""World-size-2 CPU test (gloo) of the N>1 path's host-side logic: the row-block shard planner of the library, and
the decomposition the GPU path relies on — every rank scans only ITS rows of the fixed-point matrix, the per-candidate
sums are all-reduced, and every rank takes the identical decision.  numpy stands in for the scan kernel here (test
infrastructure); the result must equal the CPU oracle's PAM on the whole matrix."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _sharded_pam(Dq, r0, r1, k):
    """The column formulation on rows [r0,r1) + all-reduce (mirrors csrc/pam.cu)."""
    n = Dq.shape[0]
    rows = Dq[r0:r1].astype(np.int64)  # this rank's shard

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    is_med = np.zeros(n, bool)
    med = []
    cap = np.full(r1 - r0, np.iinfo(np.int64).max)
    for _ in range(k):  # BUILD: argmin SP, ties -> largest index
        sp = allsum(np.minimum(rows, cap[:, None]).sum(0))
        sp[is_med] = np.iinfo(np.int64).max
        h = int(np.nonzero(sp == sp.min())[0][-1])
        med.append(h)
        is_med[h] = True
        cap = np.minimum(cap, rows[:, h])
    nsw = 0
    while True:  # SWAP
        dm = rows[:, med]
        order = np.argsort(dm, axis=1, kind="stable")
        # lowest-position medoid wins ties: stable sort over medoids sorted by index
        pos = np.argsort(np.argsort(med))
        key = dm * (k + 1) + pos[None, :]
        order = np.argsort(key, axis=1)
        near = order[:, 0]
        dn = dm[np.arange(len(dm)), near]
        ds = dm[np.arange(len(dm)), order[:, 1]] if k > 1 else dn
        td = allsum(np.array([dn.sum()]))[0]
        P = np.minimum(rows, dn[:, None])
        Q = np.minimum(rows, ds[:, None])
        sp = P.sum(0)
        aq = np.zeros((k, n), np.int64)
        for c in range(k):
            aq[c] = (Q - P)[near == c].sum(0)
        tot = allsum(np.concatenate([sp[None], aq]))
        val = tot[0][None, :] + tot[1:]  # k x n
        val[:, is_med] = np.iinfo(np.int64).max
        best = (np.iinfo(np.int64).max, -1, -1)
        vmin = val.min()
        hs = np.nonzero((val == vmin).any(axis=0))[0]
        h = int(hs[0])
        cs = [c for c in range(k) if val[c, h] == vmin]
        c = min(cs, key=lambda c: med[c])
        if k < 2 or vmin >= td:
            return med, near, nsw, td
        is_med[med[c]] = False
        is_med[h] = True
        med[c] = h
        nsw += 1


def _worker(rank, world, port, Dq, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from raceid3_stemid2_package_b200 import lib

    r0, r1 = lib.shard_rows(Dq.shape[0], rank, world)
    med, near, nsw, td = _sharded_pam(Dq, r0, r1, k)
    lab = torch.zeros(Dq.shape[0], dtype=torch.int64)
    lab[r0:r1] = torch.from_numpy(np.asarray(near) + 1)
    dist.all_reduce(lab)
    if rank == 0:
        out.put((med, lab.numpy(), nsw, int(td)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [3, 5])
def test_row_sharded_pam_equals_oracle_world2(oracle, k):
    rng = np.random.default_rng(5)
    n = 300
    cent = rng.normal(size=(5, 8)) * 2
    pts = cent[rng.integers(0, 5, n)] + rng.normal(size=(n, 8))
    D = np.sqrt(((pts[:, None] - pts[None]) ** 2).sum(-1))
    step = 2.0 ** -20
    Dq = np.rint(D / step).astype(np.int64)
    np.fill_diagonal(Dq, 0)
    ref = oracle.pam(Dq * step, k)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Dq, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    med, slot_lab, nsw, td = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(med) == sorted(ref["id_med"].tolist())
    assert nsw == ref["n_swaps"]
    assert abs(td * step / n - ref["objective"][1]) < 1e-12
    # same partition up to cluster numbering (numbering is done on the host by first appearance)
    seen, renum = {}, []
    for v in slot_lab:
        seen.setdefault(int(v), len(seen) + 1)
        renum.append(seen[int(v)])
    assert np.array_equal(np.asarray(renum), ref["clustering"])
