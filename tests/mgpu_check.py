"""Run under torchrun (one process per GPU): the row-sharded multi-GPU path must give results identical to the
single-GPU path (integer sums => independent of the number of ranks).  Exit code 0 + "MGPU OK" on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import raceid3_stemid2_package_b200 as rid  # noqa: E402
from raceid3_stemid2_package_b200.synth import synth_features_numpy  # noqa: E402


def main():
    n_cells = int(os.environ.get("MGPU_CELLS", "3000"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = rid.init_distributed(device=local)
    x, _ = synth_features_numpy(n_cells, 300, 6, 11)
    x[17] = x[4]
    samples = rid.bootstrap_samples(n_cells, 3, 17000)
    # bootmethod "subset" resamples (clustexp(samp=...)): sparse subsets, whose symmetric compaction reaches below the
    # part of the peers' blocks a replica pulls first (the library then has to fetch the rest)
    sparse_samples = rid.bootstrap_samples(n_cells, 2, 17000, samp=max(600, n_cells // 12))
    sub = np.arange(n_cells - 1, -1, -2).astype(np.int32)
    # the sparse route (rid_dist_csc): every rank uploads the non-zeros of its slice of the cells, slices go round over NVLink
    from scipy.sparse import csr_matrix
    from raceid3_stemid2_package_b200.synth import synth_counts_numpy
    cnt, _ = synth_counts_numpy(n_cells, 300, 6, 11)
    csr = csr_matrix(cnt)  # CSR of cells x genes == CSC of genes x cells
    tot = cnt.sum(axis=1)

    def run(c):
        out = {}
        dmc = c.dist_csc(csr.indices, csr.indptr, csr.data, 300, np.arange(300, dtype=np.int32), tot, "pearson")
        out["csc_blk"] = dmc.sub(rows=np.array([0, 7, n_cells // 2 + 1, n_cells - 1], dtype=np.int32))
        r = dmc.pam(3)
        out["csc_pam3"] = (r["id_med"], r["clustering"])
        dmc.free()
        for metric in ("pearson", "euclidean"):
            dm = c.dist_cellmajor(x, metric)
            r0, r1 = dm.row0, dm.row1
            out[metric + "_rows"] = (r0, r1)
            out[metric + "_blk"] = dm.sub(rows=np.array([0, 5, n_cells // 2, n_cells - 1], dtype=np.int32))
            for k in (1, 4, 7):
                r = dm.pam(k, want_sil=True)
                out[f"{metric}_pam{k}"] = (r["id_med"], r["clustering"], r["objective"], r["n_swaps"], r["avg_sil"])
            r = dm.pam(5, subset=sub)
            out[metric + "_pamsub"] = (r["id_med"], r["clustering"], r["objective"], r["n_swaps"])
            out[metric + "_logW"] = dm.gap_sweep(8)
            cb = dm.clusterboot(4, samples)
            out[metric + "_boot"] = (cb["partition"], cb["id_med"], cb["bootresult"])
            cs = dm.clusterboot(3, sparse_samples)
            out[metric + "_boot_subset"] = (cs["partition"], cs["id_med"], cs["bootresult"])
            med, ids = dm.medoids(cb["partition"])
            out[metric + "_med"] = (med, ids, dm.refresh(med))
            out[metric + "_sil"] = dm.silhouette(cb["partition"])
            # floating-point sums in an order that depends on the launch geometry: equal to the solver tolerance only
            out[metric + "_mds"] = dm.cmdscale(2)
            dm.free()
        return out

    multi = run(ctx)
    ok = True
    if rank == 0:
        single = run(rid.Context(device=local))
        for key, v in single.items():
            if key.endswith("_rows"):
                continue
            a, b = v, multi[key]
            a = a if isinstance(a, tuple) else (a,)
            b = b if isinstance(b, tuple) else (b,)
            for u, w in zip(a, b):
                atol = 1e-6 * float(np.max(np.abs(u))) if key.endswith("_mds") else 1e-12
                same = np.array_equal(np.asarray(u), np.asarray(w)) or (
                    np.asarray(u).dtype.kind == "f" and np.allclose(u, w, rtol=0, atol=atol))
                if not same:
                    ok = False
                    print("MISMATCH", key, np.asarray(u).ravel()[:8], np.asarray(w).ravel()[:8], flush=True)
        print(f"world={world} rows(rank0)={multi['pearson_rows']} " + ("MGPU OK" if ok else "MGPU FAIL"), flush=True)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
