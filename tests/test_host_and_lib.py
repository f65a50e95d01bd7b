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
