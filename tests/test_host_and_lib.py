"""CPU-side checks: the C-ABI library loads and exports every symbol include/raceid_b200.h declares, fails loudly
without a GPU, the host API validates like the reference (R/RaceID.R:753-756,895-904), R's RNG stream, the RData
fixture reader and the filterdata restatement."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from raceid3_stemid2_package_b200 import lib

    hdr = open(os.path.join(ROOT, "include", "raceid_b200.h")).read()
    declared = set(re.findall(r"\b(rid_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rid_ctx", "rid_dmat"}
    L = lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.SYMBOLS)
    assert L.rid_version() >= 100


def test_no_gpu_fails_loudly_no_cpu_fallback():
    import torch

    from raceid3_stemid2_package_b200 import lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.RaceIDError) as e:
        lib.Context(device=0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_path_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, load or link it."""
    pkg = os.path.join(ROOT, "raceid3_stemid2_package_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".R", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "orc_"):
                    assert needle not in txt, (f, needle)


def test_shard_rows_cover_and_align():
    from raceid3_stemid2_package_b200 import lib

    for n in (5, 127, 128, 129, 1000, 100000, 250000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                a, b = lib.shard_rows(n, r, world)
                assert a == prev and b >= a
                assert a % 128 == 0 or a == n
                prev = b
            assert prev == n


def test_rrng_known_answers():
    """R >= 3.6 known-answer values (Mersenne-Twister + rejection sampling)."""
    from raceid3_stemid2_package_b200.rrng import RRng, r_unique

    assert np.allclose(RRng(42).unif_rand(3), [0.9148060435, 0.9370754133, 0.2861395348], atol=1e-9)
    assert np.allclose(RRng(1).unif_rand(3), [0.2655086631, 0.3721238996, 0.5728533633], atol=1e-9)
    assert list(RRng(42).sample_int(10)) == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    assert list(RRng(123).sample_int(10)) == [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]
    assert list(RRng(1).sample_int(10)) == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    # bulk and one-at-a-time draws consume the stream identically
    a = RRng(7).sample_int(1000, 1000, replace=True)
    r = RRng(7)
    b = np.concatenate([r.sample_int(1000, 1, replace=True) for _ in range(1000)])
    assert np.array_equal(a, b)
    assert list(r_unique(np.array([3, 1, 3, 2, 1]))) == [3, 1, 2]


def test_api_validation_matches_reference_messages():
    import raceid3_stemid2_package_b200 as rid

    sc = rid.SCseq(np.ones((3, 3)))
    with pytest.raises(ValueError, match="metric has to be one of"):
        rid.compdist(sc, metric="cosine")
    with pytest.raises(ValueError, match="Run compdist before clustexp"):
        rid.clustexp(sc)
    sc.distances = np.zeros((3, 3))
    for kw, msg in [(dict(clustnr=0), "clustnr has to be a positive integer"),
                    (dict(bootnr=1.5), "bootnr has to be a positive integer"),
                    (dict(cln=-1), "cln has to be a non-negative integer"),
                    (dict(sat=False), "cln has to be a positive integer or sat has to be TRUE"),
                    (dict(samp=1), "samp needs to be an integer >1"),
                    (dict(FUNcluster="dbscan"), "FUNcluster needs to be one of")]:
        with pytest.raises(ValueError, match=msg):
            rid.clustexp(sc, **kw)
    # the consumers of the matrix after clustexp: plotsilhouette (R/RaceID.R:1276-1277) and comptsne's start configuration
    sc2 = rid.SCseq(np.ones((3, 3)))
    with pytest.raises(ValueError, match="run compdist before plotsilhouette"):
        rid.silhouette(sc2)
    with pytest.raises(ValueError, match="run compdist before comptsne"):
        rid.tsne_initial_config(sc2)
    sc2.distances = np.zeros((3, 3))
    with pytest.raises(ValueError, match="run clustexp before plotsilhouette"):
        rid.silhouette(sc2)
    sc2.cluster["kpart"] = np.ones(3, dtype=np.int32)
    with pytest.raises(ValueError, match="only a single cluster: no silhouette plot"):
        rid.silhouette(sc2)
    with pytest.raises(ValueError, match="more than one row"):
        rid.SCseq(np.ones((1, 5)))
    with pytest.raises(ValueError, match="negative values"):
        rid.SCseq(-np.ones((3, 3)))


def test_rdata_reader_and_filterdata_on_bundled_fixture():
    """Only runs where the reference checkout exists (this container); the GPU box uses tests/golden instead."""
    path = "/root/reference/data/intestinalDataSmall.RData"
    if not os.path.exists(path):
        pytest.skip("reference data not present")
    import raceid3_stemid2_package_b200 as rid
    from raceid3_stemid2_package_b200.rdata import load_dgcmatrix

    M, genes, cells = load_dgcmatrix(path)
    assert M.shape == (213, 103) and int((M != 0).sum()) == 20801
    sc = rid.filterdata(rid.SCseq(M, genes_all=genes, cells_all=cells))
    x = rid.getfdata(sc, g=sc.cluster["features"])
    g = load_golden("cfg1_intestinalDataSmall.npz")
    assert x.shape == g["x"].shape and np.array_equal(x, g["x"])
    assert abs(x.min() - 0.1) < 1e-12


def test_c_sampler_equals_python_mirror_of_r_rng():
    """rid_r_bootstrap_samples (host-only C restatement of R's Mersenne-Twister + rejection sampling + do_sample) against
    the Python mirror that carries R's known answers, for both bootmethods and the awkward sizes (1, 2, 2^16, 2^16+1)."""
    from raceid3_stemid2_package_b200 import lib
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples, bootstrap_samples_py

    for n, B, seed, samp in [(1000, 4, 17000, None), (70001, 2, 42, None), (5000, 3, 7, 1200), (10, 3, 1, None), (2, 3, 1, None),
                             (300, 2, 5, 1), (65536, 2, 9, None), (65537, 2, 9, None), (1, 2, 3, None), (50, 2, 8, 80)]:
        a = bootstrap_samples(n, B, seed, samp)
        b = bootstrap_samples_py(n, B, seed, samp)
        assert len(a) == len(b) == B
        for u, v in zip(a, b):
            assert np.array_equal(u, v), (n, B, seed, samp)
    cat, lens = bootstrap_samples(1000, 4, 17000, packed=True)
    assert lens.dtype == np.int32 and cat.dtype == np.int32 and int(lens.sum()) == len(cat)
    assert lib.r_bootstrap_samples(10, 0, 1)[1].size == 0


def test_bench_fp64_rows_restatement_matches_oracle(oracle):
    """bench.py's parity_spot compares sampled rows of the device matrix with an FP64 numpy restatement of cor()/dist();
    that restatement is held to the oracle here (all four metrics, ties, a duplicated cell)."""
    import bench
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    x, _ = synth_features_numpy(90, 70, 3, 2)
    x[11] = x[3]
    for metric in ("pearson", "spearman", "logpearson", "euclidean"):
        Do, _ = oracle.dist(np.ascontiguousarray(x.T), metric)
        Z = bench.transform_rows_fp64(x, metric)
        if metric == "euclidean":
            ref = np.sqrt(((Z[:, None, :] - Z[None, :, :]) ** 2).sum(-1))
        else:
            ref = 1.0 - np.clip(Z @ Z.T, -1.0, 1.0)
            np.fill_diagonal(ref, 0.0)
        assert np.max(np.abs(ref - Do)) < 1e-12, metric
    out = {"partition": np.arange(5), "id_med": np.array([1, 2]), "bootresult": np.ones((2, 3)), "logW": None}
    d1 = bench.digest_of(out)
    out["partition"] = out["partition"].astype(np.int64)
    assert bench.digest_of(out) == d1 and len(d1) == 40
    out["bootresult"][0, 0] = np.nextafter(1.0, 2.0)
    assert bench.digest_of(out) != d1


def test_sparse_scseq_filterdata_equals_dense():
    """SCseq accepts a scipy sparse count matrix (upstream: dgCMatrix) and keeps it sparse; filterdata / getfdata give
    the dense route's genes, features and values."""
    import raceid3_stemid2_package_b200 as rid
    from scipy.sparse import csc_matrix

    rng = np.random.default_rng(15)
    counts = rng.poisson(0.5, size=(300, 220)).astype(np.float64)
    counts[:, counts.sum(0) == 0] = 1.0
    a = rid.filterdata(rid.SCseq(counts), mintotal=150, minexpr=1, minnumber=3)
    b = rid.filterdata(rid.SCseq(csc_matrix(counts)), mintotal=150, minexpr=1, minnumber=3)
    assert 0 < len(a.cell_names) < 220 and a.cell_names == b.cell_names
    assert a.genes == b.genes and a.cluster["features"] == b.cluster["features"]
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(rid.getfdata(a), rid.getfdata(b))
    assert np.array_equal(rid.getfdata(a, g=a.cluster["features"][:5], n=a.cell_names[:7]),
                          rid.getfdata(b, g=a.cluster["features"][:5], n=a.cell_names[:7]))
    c = rid.SCseq.from_filtered_counts(csc_matrix(counts))
    tot = counts.sum(0)
    assert np.array_equal(rid.getfdata(c), counts / tot * tot.min() + 0.1)
    with pytest.raises(ValueError):
        rid.SCseq(csc_matrix(-counts))


def test_findoutliers_merge_validation_messages():
    import raceid3_stemid2_package_b200 as rid

    sc = rid.SCseq(np.ones((3, 4)))
    with pytest.raises(ValueError, match="outdistquant has to be a number between 0 and 1"):
        rid.findoutliers_merge(sc, [], outdistquant=2)
    with pytest.raises(ValueError, match="outdistquant has to be a number between 0 and 1"):
        rid.findoutliers_merge(sc, [], outdistquant="a")
    with pytest.raises(ValueError, match="run clustexp before findoutliers"):
        rid.findoutliers_merge(sc, [])
