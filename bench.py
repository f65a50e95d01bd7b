#!/usr/bin/env python
"""bench.py — compdist + clustexp(kmedoids) wall seconds on synthetic UMI matrices (BASELINE.json's metric).

A "step" is one pass of the hot path over one synthetic matrix: compdist(metric) -> clustexp(sat, cln, clustnr, bootnr).
The default workload is BASELINE.json configs[3]: 100 000 cells x 3 000 genes, pearson, `clustexp(sc, cln=30)` — i.e.
R's defaults sat=TRUE (the 30-k clusGapExt sweep runs although cln is given, R/RaceID.R:894,904), clustnr=30, bootnr=50.

  value : seconds per step with the feature matrix already resident in HBM, through the C ABI (rid_dist ->
          rid_gap_sweep -> rid_clusterboot); CUDA events on the library's stream, barrier + synchronize on both sides,
          max over ranks
  e2e   : seconds per step through the reference-facing API, `rid.compdist(sc)` -> `rid.clustexp(sc, ...)` on an SCseq
          object that holds the count matrix the way the R package does (dgCMatrix = CSC, here in page-locked host
          memory): H2D copy of the CSC triplet inside the timed region (getfdata is fused on the device, rid_dist_csc),
          R-compatible bootstrap index vectors drawn inside it, partition / jaccard / medoids / logW read back
  roofline : the dominant kernel family (k_scan, the PAM streaming pass of BUILD and SWAP, HBM-bound): algorithmic
          bytes / CUDA-event duration of its launches inside the timed region, against MEASURED_PEAKS.json's hbm_gbs;
          detail.roofline_all holds the same object for the compaction, incremental-SWAP and distance kernels
  parity_spot / result_digest : checked OUTSIDE the timed region at the benchmarked size — sampled rows of D against
          FP64 numpy, a subset PAM and a bootstrap replicate against the CPU oracle; sha1 of the step's results
          (identical for every GPU count) and the number of accepted swaps
  detail.sat_false / swap_heavy / cfg5 / cooperative : the lighter sat=FALSE variant, a ring-structured synthetic on
          which SWAP really iterates, BASELINE configs[4] (250k cells, logpearson, needs >= 2 GPUs) and the all-reduce
          schedule (RID_NO_REPLICA=1) of the headline workload
  cpu_baseline / --impl reference : the CPU oracle (restatement of the reference's R/C path; R itself is not installed
          in this image) on a bounded sample of the same workload, extrapolated ~ N^2

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); the distance matrix is row-block sharded and the
library runs its own NCCL collectives and NVLink peer copies.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "compdist+clustexp wall s at 100k cells; dist TFLOP/s; PAM swap GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=3000)
    ap.add_argument("--metric", default="pearson")
    ap.add_argument("--cln", type=int, default=30)
    ap.add_argument("--sat", type=int, default=1, help="1: clustexp's default sat=TRUE (the k sweep runs); 0: sat=FALSE")
    ap.add_argument("--clustnr", type=int, default=30)
    ap.add_argument("--bootnr", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1004)
    ap.add_argument("--structure", default="clusters", choices=["clusters", "ring"])
    ap.add_argument("--clusters", type=int, default=0, help="true clusters / profiles of the synthetic data (0: cln)")
    ap.add_argument("--cpu-cells", type=int, default=0,
                    help="cells in the CPU sample (0: 2500 for the 1-thread leg, ~10 s; 6000 for the all-threads reference arm)")
    ap.add_argument("--cpu-bootnr", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the detail legs (sat_false, swap_heavy, cfg5, cooperative)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--timed-kinds", default="gram,gram_i8,scan_swap,scan_build,scan_delta,scan_group,compact",
                    help="kernel kinds bracketed with CUDA events inside the timed region")
    return ap.parse_args()


def workload_name(a, cells=None, metric=None, sat=None, bootnr=None, structure=None):
    return (f"synthetic {cells or a.cells} cells x {a.genes} genes ({structure or a.structure}), metric={metric or a.metric}, "
            f"clustexp(sat={bool(a.sat if sat is None else sat)}, cln={a.cln}, clustnr={a.clustnr}, "
            f"bootnr={a.bootnr if bootnr is None else bootnr}, FUNcluster=kmedoids)")


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (oracle): bounded sample, extrapolated ~ N^2 (distance: N^2 G; PAM: k N^2 per iteration, iteration counts
# assumed equal to the sample's) and linearly in bootnr
# ---------------------------------------------------------------------------------------------------------------
def cpu_sample(a, omp: bool):
    import numpy as np

    from oracle import oracle as orc
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    L = orc.lib(omp)
    if omp:
        # launchers (torchrun) export OMP_NUM_THREADS=1: this arm is "the reference's CPU path with all the host threads
        # it can use", so ask for every core explicitly
        L.orc_set_threads(os.cpu_count() or 1)
    cores = L.orc_num_threads()
    ns = min(a.cpu_cells or (6000 if omp else 2500), a.cells)
    x, _ = synth_features_numpy(ns, a.genes, a.cln, a.seed, a.structure)
    xt = np.ascontiguousarray(x.T)
    t0 = time.perf_counter()
    D, _ = orc.dist(xt, a.metric, omp=omp)
    t_dist = time.perf_counter() - t0
    Dc = np.ascontiguousarray(D.T)
    k = min(a.cln, ns - 1)
    t_sweep = 0.0
    nsw = min(ns, 3000 if omp else 1000)  # the 30 PAM runs of the sweep get a smaller sample of their own
    if a.sat:
        sub = np.ascontiguousarray(np.arange(nsw), dtype=np.int32)
        t0 = time.perf_counter()
        out = np.zeros(min(a.clustnr, nsw - 1))
        L.orc_gap_sweep(Dc, ns, sub.ctypes.data, nsw, len(out), out)
        t_sweep = time.perf_counter() - t0
    bs = min(a.cpu_bootnr, a.bootnr)
    samples = bootstrap_samples(ns, bs, 17000)
    cat = np.ascontiguousarray(np.concatenate(samples), dtype=np.int32)
    lens = np.ascontiguousarray([len(s) for s in samples], dtype=np.int32)
    part = np.zeros(ns, dtype=np.int32)
    med = np.zeros(k, dtype=np.int32)
    br = np.zeros((bs, k))
    bm = np.zeros(k)
    t0 = time.perf_counter()
    L.orc_clusterboot(Dc, ns, k, cat, lens, 0, part, med, br, bm)  # c1 only
    t_c1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    L.orc_clusterboot(Dc, ns, k, cat, lens, bs, part, med, br, bm)
    t_boot = max(time.perf_counter() - t0 - t_c1, 0.0)
    s2 = (a.cells / ns) ** 2
    s2w = (a.cells / nsw) ** 2
    full = (t_dist + t_c1) * s2 + t_sweep * s2w + t_boot * s2 * (a.bootnr / max(bs, 1))
    sample = (f"oracle port ({'OpenMP ' + str(cores) + ' threads' if omp else '1 thread'}) on {ns} cells x {a.genes} genes, "
              f"k={k}, {bs} of {a.bootnr} bootstrap replicates: dist {t_dist:.2f}s, sweep {t_sweep:.2f}s"
              f"{' (on the first ' + str(nsw) + ' cells, x' + format(s2w, '.0f') + ')' if a.sat else ''}, pam(c1) {t_c1:.2f}s, "
              f"boot {t_boot:.2f}s; extrapolated x(N/n)^2={s2:.0f} (distance ~N^2 G, PAM ~k N^2 per iteration with the "
              f"sample's iteration counts) and linearly in bootnr")
    return {"value": full, "unit": "s", "cores": cores, "kind": "port", "sample": sample,
            "sample_seconds": t_dist + t_sweep + t_c1 + t_boot}


def run_reference(a):
    """The reference's CPU path (oracle port, all host threads) on a bounded sample.  The sample is deterministic, so
    it is timed once (twice if it is short: the faster run counts) instead of warmup+steps times; rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    info = cpu_sample(a, omp=True)
    runs = 1
    if info["sample_seconds"] < 30.0:
        again = cpu_sample(a, omp=True)
        runs = 2
        if again["value"] < info["value"]:
            info = again
    v = info["value"]
    info["sample"] += f"; sample timed {runs}x (deterministic input), best run reported"
    line = {"metric": METRIC, "value": v, "unit": "s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": {"workload": workload_name(a)},
            "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons DURING the timed region.  In-process NVML polling (one cheap call per 500 ms); a
    polling `nvidia-smi -lms` process proved intrusive (it stretched driver calls of the timed step on some boxes) and is
    only the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._thr = None
        self.how = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.device]) if vis and vis.split(",")[0].isdigit() else self.device
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mx.append(float(mx))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in self.BITS.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self._stop.wait(0.5)

            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "1000", "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(1.0)
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        for r in self.rows:
            p = [t.strip() for t in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": self.how}


# ---------------------------------------------------------------------------------------------------------------
# results of one step -> digest (identical for every GPU count: all sums are exact integers)
# ---------------------------------------------------------------------------------------------------------------
def digest_of(out) -> str:
    import numpy as np

    h = hashlib.sha1()
    h.update(np.ascontiguousarray(out["partition"], dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(out["id_med"], dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(out["bootresult"], dtype=np.float64).tobytes())
    if out.get("logW") is not None:
        h.update(np.ascontiguousarray(out["logW"], dtype=np.float64).tobytes())
    return h.hexdigest()


def transform_rows_fp64(x_ng, metric):
    """FP64 numpy restatement of what cor()/dist() see for each cell (rows of x): centred unit rows for the correlation
    metrics (average-tie ranks for spearman, log2 for logpearson), raw rows for euclidean."""
    import numpy as np

    x = np.asarray(x_ng, dtype=np.float64)
    if metric == "euclidean":
        return x
    if metric == "logpearson":
        x = np.log2(x)
    elif metric == "spearman":
        from scipy.stats import rankdata

        x = rankdata(x, method="average", axis=1)
    x = x - x.mean(axis=1, keepdims=True)
    x = x - x.mean(axis=1, keepdims=True)  # second pass like cov.c
    n = np.sqrt((x * x).sum(axis=1, keepdims=True))
    return x / n


def parity_spot(ctx, dm, x_host, a, metric, rank, rows=64, sub_cells=4000):
    """Value-level checks at the benchmarked size (VERDICT r1 #1a), outside any timed region.  Collective: every rank
    calls the library; rank 0 compares."""
    import numpy as np

    from oracle import oracle as orc
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples

    N = dm.N
    rng = np.random.default_rng(12345)
    ridx = np.sort(rng.choice(N, size=min(rows, N), replace=False)).astype(np.int32)
    got = dm.sub(ridx, None)  # rows x N
    out = {"rows": int(len(ridx))}
    t0 = time.perf_counter()
    if rank == 0:
        Z = transform_rows_fp64(x_host, metric)
        if metric == "euclidean":
            ref = np.stack([np.sqrt(((Z - Z[r]) ** 2).sum(1)) for r in ridx])
        else:
            ref = 1.0 - np.clip(Z[ridx] @ Z.T, -1.0, 1.0)
            ref[np.arange(len(ridx)), ridx] = 0.0
        out["max_abs_D"] = float(np.max(np.abs(got - ref)))
        out["tolerance"] = 1e-6
        del Z
    # a PAM run on a subset of the resident matrix, and c1 + one bootstrap replicate on that sub-matrix, against the oracle
    S = rng.permutation(N)[:min(sub_cells, N)].astype(np.int32)  # order significant (positions drive the tie rules)
    k = min(a.cln, len(S) - 1)
    r = dm.pam(k, subset=S)
    Dsub = dm.sub(S, S)
    dm2 = ctx.from_matrix(Dsub)
    smp = bootstrap_samples(len(S), 1, 17000)
    rb = dm2.clusterboot(k, smp)
    dm2.free()
    if rank == 0:
        ro = orc.pam(Dsub, k)
        out["pam_identical"] = bool(np.array_equal(r["id_med"], ro["id_med"]) and np.array_equal(r["clustering"], ro["clustering"])
                                    and r["n_swaps"] == ro["n_swaps"])
        out["pam_subset_cells"] = int(len(S))
        out["pam_swaps"] = int(ro["n_swaps"])
        bo = orc.clusterboot(Dsub, k, smp)
        out["boot_identical"] = bool(np.array_equal(rb["partition"], bo["partition"]) and
                                     np.array_equal(rb["bootresult"], np.asarray(bo["bootresult"])))
        out["seconds"] = time.perf_counter() - t0
    return out


# ---------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    import numpy as np
    import torch

    import raceid3_stemid2_package_b200 as rid
    from raceid3_stemid2_package_b200.synth import synth_features_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: libraceid_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = rid.init_distributed(device=local)
    else:
        ctx = rid.Context(device=local)
        rid.set_context(ctx)

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t]

    N, G = a.cells, a.genes
    dev = torch.device("cuda", local)
    KINDS = ("gram", "gram_i8", "scan_swap", "scan_delta", "scan_build", "scan_group", "prep", "compact")

    # ceilings measured on this GPU right now (outside the timed region)
    dmma_peak = ctx.measure_dmma_peak()
    read_peak = ctx.measure_read_peak(8.0)
    i8_peak = ctx.measure_i8_peak()

    class Workload:
        """One synthetic matrix resident in HBM (identical on every rank) + the step over it through the C ABI."""

        def __init__(self, cells, metric, sat, bootnr, seed, structure, clusters=None):
            self.N, self.metric, self.sat, self.bootnr = cells, metric, sat, bootnr
            self.x_dev, _ = synth_features_torch(cells, G, clusters or a.cln, seed, dev, structure)
            torch.cuda.synchronize()
            # R-compatible index vectors; for the C-ABI leg they are an input of rid_clusterboot, drawn once
            self.samples = rid.bootstrap_samples(cells, bootnr, 17000, packed=True)
            self.state = {}

        def step(self, keep=False):
            dm = ctx.dist((self.x_dev.data_ptr(), G, self.N), self.metric)
            out = {"t_dist": ctx.last_ms}
            if self.sat:
                out["logW"] = dm.gap_sweep(min(a.clustnr, self.N - 1))
                out["t_sweep"] = ctx.last_ms
            r = dm.clusterboot(a.cln, self.samples)
            out["t_boot"] = ctx.last_ms
            out.update(partition=r["partition"], bootmean=r["bootmean"], bootresult=r["bootresult"], id_med=r["id_med"])
            self.state = out
            if keep:
                return dm
            dm.free()
            return None

        def timed(self, steps, warmup, kinds=None):
            for _ in range(warmup):
                self.step()
            barrier()
            if kinds is not None:
                # the CUDA events that bracket the kernels are created here; their first use is slow (the first step after
                # this call measured ~0.3 s longer), so one untimed step records on all of them before they are recycled
                ctx.profile(kinds)
                self.step()
                ctx.profile(kinds)
            l0, s0 = ctx.launches, ctx.swaps
            barrier()
            ctx.event_record(0)
            t0 = time.perf_counter()
            walls, step_stats = [], []
            for _ in range(steps):
                ts = time.perf_counter()
                self.step()
                walls.append((time.perf_counter() - ts) * 1e3)
                st_ = ctx.stats()
                step_stats.append([round(self.state.get("t_dist", 0.0), 1), round(self.state.get("t_sweep", 0.0), 1),
                                   round(st_["c1_ms"], 1), round(st_["first_picks_and_replica_wait_ms"], 1),
                                   round(st_["replicates_ms"], 1), round(st_["replicates_host_enqueue_ms"], 1)])
            ctx.event_record(1)
            barrier()
            wall = time.perf_counter() - t0
            dev_ms = ctx.event_elapsed_ms(0, 1)
            prof = {k: ctx.profile_get(k) for k in KINDS} if kinds is not None else None
            if kinds is not None:
                ctx.profile(False)
            ms, wl = allmax([dev_ms, wall * 1e3])
            swaps = allsum([ctx.swaps - s0])[0]
            st_last = ctx.stats()
            mx = allmax([st_last["c1_ms"], st_last["first_picks_and_replica_wait_ms"], st_last["replicates_ms"], st_last["replica_pull_ms"]])
            mn = allmax([-st_last["replicates_ms"], -st_last["replica_pull_ms"]])
            stats_ranks = {"max_c1_ms": mx[0], "max_first_picks_and_replica_wait_ms": mx[1], "max_replicates_ms": mx[2],
                           "max_replica_pull_ms": mx[3], "min_replicates_ms": -mn[0], "min_replica_pull_ms": -mn[1]}
            return {"ms_per_step": ms / steps, "wall_ms_per_step": wl / steps, "launches": ctx.launches - l0,
                    "n_swaps_per_step": swaps / steps, "prof": prof, "lib_stats_rank0_last_step": st_last,
                    "lib_stats_over_ranks_last_step": stats_ranks,
                    "step_wall_ms_rank0": walls,
                    "step_phases_rank0_dist_sweep_c1_first_replicates_hostenqueue_ms": step_stats}

    # ---- headline: resident leg, W warm-up + K timed steps ---------------------------------------------------
    W = Workload(N, a.metric, bool(a.sat), a.bootnr, a.seed, a.structure, a.clusters or None)
    for _ in range(a.warmup):
        W.step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_NVML"):  # one sampler per job: eight concurrent pollers slow every rank's driver calls down
        clocks.start()
    # inside the timed region the kernels of the roofline objects are bracketed with CUDA events (pre-created); the
    # per-kind breakdown of everything comes from one extra, untimed step below
    res = W.timed(a.steps, 0, [k for k in a.timed_kinds.split(",") if k])
    clk = clocks.stop()
    ms_per_step, wall_ms_per_step, launches, prof = res["ms_per_step"], res["wall_ms_per_step"], res["launches"], res["prof"]
    state = W.state
    digest = digest_of(state)
    d2h = sum(int(np.asarray(state[k]).nbytes) for k in ("partition", "bootresult", "id_med")) + \
        (int(state["logW"].nbytes) if state.get("logW") is not None else 0)

    # ---- untimed breakdown step: every kernel kind bracketed (its events slow the step down a little; shares only) ----
    ctx.profile(True)
    ctx.event_record(2)
    W.step()
    ctx.event_record(3)
    barrier()
    breakdown_ms = ctx.event_elapsed_ms(2, 3)
    prof_all = {k: ctx.profile_get(k) for k in KINDS}
    ctx.profile(False)

    # ---- parity at the benchmarked size (outside the timed regions) -------------------------------------------
    spot = None
    x_host_dense = None
    if not a.no_parity:
        x_host_dense = W.x_dev.cpu().numpy() if rank == 0 else None
        dm = W.step(keep=True)
        spot = parity_spot(ctx, dm, x_host_dense, a, a.metric, rank)
        dm.free()
        x_host_dense = None

    # ---- e2e leg: the reference-facing API on an SCseq object with the count matrix in CSC, host outputs ---------------
    e2e = None
    e2e_detail = None
    if not a.no_e2e:
        from scipy.sparse import csc_matrix

        from raceid3_stemid2_package_b200.synth import synth_features_torch as _gen

        cnt, _ = _gen(N, G, a.cln, a.seed, dev, a.structure, want_counts=True)  # N x G counts; CSR of it == CSC of genes x cells
        sp = cnt.to_sparse_csr()
        del cnt

        def pinned(t, dtype):
            h = torch.empty(t.shape, dtype=dtype, pin_memory=True)
            h.copy_(t.to(dtype))
            return h.numpy()

        indptr, indices, data = pinned(sp.crow_indices(), torch.int32), pinned(sp.col_indices(), torch.int32), pinned(sp.values(), torch.float64)
        del sp
        torch.cuda.synchronize()
        counts_csc = csc_matrix((data, indices, indptr), shape=(G, N), copy=False)
        sc = rid.SCseq.from_filtered_counts(counts_csc)
        # (several ranks: each uploads the non-zeros of its own slice of the cells; the small tables go up on every rank)
        h2d = int(data.nbytes + indices.nbytes + world * (indptr.nbytes + N * 8 + G * 4))

        def api_step():
            rid.compdist(sc, metric=a.metric)
            rid.clustexp(sc, sat=bool(a.sat), cln=a.cln, clustnr=a.clustnr, bootnr=a.bootnr, rseed=17000, verbose=False)
            return sc

        api_step()  # warm-up of the host path
        barrier()
        nst = max(1, min(a.steps, a.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(nst):
            api_step()
        barrier()
        e = allmax([(time.perf_counter() - t0) / nst])[0]
        clb = sc.cluster["clb"]
        api_out = {"partition": sc.cluster["kpart"], "id_med": clb["result"]["result"]["pamobject"]["id.med"] - 1,
                   "bootresult": clb["bootresult"], "logW": sc.cluster["gap"]["Tab"][:, 0] if sc.cluster.get("gap") else None}
        e2e = {"value": e, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        e2e_detail = {"steps": nst, "api": "rid.compdist(sc) -> rid.clustexp(sc) on SCseq.from_filtered_counts(CSC, page-locked)",
                      "digest": digest_of(api_out), "digest_equals_c_abi_leg": digest_of(api_out) == digest,
                      "csc_nnz": int(len(data))}
        sc.distances.free()
        del sc, counts_csc, data, indices, indptr

    # ---- rooflines -------------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = ("measured copy bandwidth, MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks
               else "fallback 6650 GB/s (B200_PROFILING.md)")
    try:
        ratios = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_ratios.json")))
    except Exception:
        ratios = {}

    def traffic_of(kernel, alg_bytes_per_launch):
        r = ratios.get(kernel, {}).get("dram_per_algorithmic")
        return alg_bytes_per_launch * r if r else None

    def rooflines(prof, step_ms_total, Nw, worldw):
        roofs = []
        # k_scan<MODE>: one streaming body, three instantiations (0: BUILD step 0, 1: SWAP / grouped sums, 2: incremental BUILD)
        sc_ = {f: sum(prof[k][f] for k in ("scan_swap", "scan_build", "scan_group")) for f in ("launches", "ms", "work", "traffic")}
        if sc_["launches"] and sc_["ms"] > 0:
            ach = sc_["work"] / (sc_["ms"] * 1e-3) / 1e9
            roofs.append({"bound": "hbm", "kernel": "k_scan<MODE> / k_colsum_mma (PAM streaming passes: BUILD steps incl. the batched first picks, SWAP scan, grouped column sums of the sweep's W.k)",
                          "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                          "traffic": traffic_of("k_scan", sc_["work"] / sc_["launches"]), "peak_source": hbm_src,
                          "read_only_stream_gbs_measured_here": read_peak, "launches": sc_["launches"],
                          "avg_launch_ms": sc_["ms"] / sc_["launches"],
                          "algorithmic_bytes_per_launch": sc_["work"] / sc_["launches"],
                          "requested_bytes_per_launch": sc_["traffic"] / sc_["launches"],
                          "by_mode": {k: {"launches": prof[k]["launches"], "ms": prof[k]["ms"],
                                          "gbs": prof[k]["work"] / (prof[k]["ms"] * 1e-3) / 1e9 if prof[k]["ms"] else None}
                                      for k in ("scan_swap", "scan_build", "scan_group")},
                          "note": "bytes = rows scanned x subset size x 4, every distance read once per pass; the kernel only "
                                  "reads, so the measured copy bandwidth (read + write) understates its ceiling: "
                                  "read_only_stream_gbs_measured_here is the read-only streaming figure of this run; traffic = "
                                  "algorithmic bytes x the DRAM/algorithmic ratio of the ncu capture",
                          "share_of_step": sc_["ms"] / step_ms_total})
        sd = prof["scan_delta"]
        if sd["launches"] and sd["ms"] > 0:
            ach = sd["work"] / (sd["ms"] * 1e-3) / 1e9
            roofs.append({"bound": "hbm", "kernel": "k_scan_delta (incremental SWAP: rows whose nearest / second medoid changed)",
                          "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                          "peak_source": hbm_src, "launches": sd["launches"], "avg_launch_ms": sd["ms"] / sd["launches"],
                          "algorithmic_bytes_per_launch": sd["work"] / sd["launches"],
                          "requested_bytes_per_launch": sd["traffic"] / sd["launches"],
                          "note": "launches enqueued ahead of the local optimum are no-ops and count with ~2 us each",
                          "share_of_step": sd["ms"] / step_ms_total})
        cp = prof["compact"]
        if cp["launches"] and cp["ms"] > 0:
            ach = cp["work"] / (cp["ms"] * 1e-3) / 1e9
            roofs.append({"bound": "hbm", "kernel": "k_compact_sym (D[bsamp,bsamp] copy of a bootstrap replicate, fused with BUILD step 0)",
                          "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                          "traffic": traffic_of("k_compact_sym", cp["work"] / cp["launches"]),
                          "peak_source": hbm_src, "launches": cp["launches"], "avg_launch_ms": cp["ms"] / cp["launches"],
                          "algorithmic_bytes_per_launch": cp["work"] / cp["launches"],
                          "requested_bytes_per_launch": cp["traffic"] / cp["launches"],
                          "note": "one GPU / full replica: algorithmic bytes = upper triangle of D[bsamp,bsamp] gathered once + the "
                                  "square written (1.5 m^2 x 4); requested = the 32-byte sectors those gathers touch (x N/m) + "
                                  "the square.  Row shards use k_compact (whole rows read: rows x (N + m) x 4)",
                          "share_of_step": cp["ms"] / step_ms_total})
        gm = prof["gram"]
        if gm["launches"] and gm["ms"] > 0:
            ach = gm["work"] / (gm["ms"] * 1e-3) / 1e12
            roofs.append({"bound": "tensor", "kernel": "k_gram_dmma (FP64 DMMA distance contraction)", "achieved": ach,
                          "peak": dmma_peak, "unit": "TFLOP/s", "frac": ach / dmma_peak if dmma_peak else None,
                          "traffic": None,
                          "peak_source": "FP64 DMMA issue-rate ceiling measured in this run (rid_measure_dmma_peak); "
                                         "MEASURED_PEAKS.json holds no FP64 figure (nominal B200 FP64 tensor: 40 TFLOP/s)",
                          "launches": gm["launches"], "avg_launch_ms": gm["ms"] / gm["launches"],
                          "issued_flops_per_launch": gm["work"] / gm["launches"],
                          "algorithmic_flops_per_launch": 2.0 * Nw * Nw * G / worldw,
                          "note": "work = flops actually issued (upper-triangle tiles only; the mirror tile is a store)",
                          "share_of_step": gm["ms"] / step_ms_total})
        g8 = prof["gram_i8"]
        if g8["launches"] and g8["ms"] > 0:
            ach = g8["work"] / (g8["ms"] * 1e-3) / 1e12
            bf16 = float(peaks.get("bf16_tflops", 1590.0))
            roofs.append({"bound": "tensor", "kernel": "k_gram_i8_persist (tcgen05.mma kind::i8, exact Ozaki-split distance contraction)",
                          "achieved": ach, "peak": i8_peak, "unit": "TOP/s", "frac": ach / i8_peak if i8_peak else None,
                          "traffic": (ratios.get("k_gram_i8_persist", {}).get("dram_bytes_per_launch_cfg4")
                                      if (Nw, G, worldw) == (100000, 3000, 1) else None),
                          "peak_source": "int8 tcgen05 issue-rate ceiling measured in this run (rid_measure_i8_peak: M=128 N=192 "
                                         "K=32 kind::i8 from shared memory, one CTA per SM, no loads)",
                          "frac_of_2x_measured_bf16_burst": ach / (2.0 * bf16),
                          "launches": g8["launches"], "avg_launch_ms": g8["ms"] / g8["launches"],
                          "issued_int8_ops_per_launch": g8["work"] / g8["launches"],
                          "fp64_equivalent_tflops_2N2G": 2.0 * Nw * Nw * G * g8["launches"] / (g8["ms"] * 1e-3) / 1e12,
                          "note": "work = digit products issued x 2*128*96*Kp per computed tile (tiles touching the upper triangle)",
                          "share_of_step": g8["ms"] / step_ms_total})
        roofs.sort(key=lambda r: -r["share_of_step"])
        return roofs

    roofs = rooflines(prof, ms_per_step * a.steps, N, world)
    roof = roofs[0] if roofs else None
    sw = prof["scan_swap"]
    detail = {"breakdown_step_ms": breakdown_ms}
    for k, v in prof_all.items():
        if not v["launches"]:
            continue
        # shares against the TIMED step: the breakdown step itself is stretched by the events that bracket every kernel
        d = {"launches": v["launches"], "ms_total": v["ms"], "share_of_step": v["ms"] / ms_per_step}
        if k == "gram_i8":
            d["int8_tops_issued"] = v["work"] / (v["ms"] * 1e-3) / 1e12
            d["tflops_algorithmic_2N2G"] = 2.0 * N * N * G / world * v["launches"] / (v["ms"] * 1e-3) / 1e12
            d["tflops_algorithmic_2N2G_timed"] = (2.0 * N * N * G / world * prof[k]["launches"] / (prof[k]["ms"] * 1e-3) / 1e12
                                                  if prof[k]["launches"] and prof[k]["ms"] > 0 else None)
        elif k == "gram":
            d["tflops_issued"] = v["work"] / (v["ms"] * 1e-3) / 1e12
            d["tflops_algorithmic_2N2G"] = 2.0 * N * N * G / world * v["launches"] / (v["ms"] * 1e-3) / 1e12
        elif v["ms"] > 0:
            d["algorithmic_gbs"] = v["work"] / (v["ms"] * 1e-3) / 1e9
            d["requested_gbs"] = v["traffic"] / (v["ms"] * 1e-3) / 1e9
        detail[k] = d
    detail["phase_ms_last_step"] = {k: state.get(k) for k in ("t_dist", "t_sweep", "t_boot") if k in state}
    detail["wall_ms_per_step"] = wall_ms_per_step
    detail["clusters_found"] = int(state["partition"].max())
    detail["bootmean_min"] = float(np.min(state["bootmean"]))
    detail["roofline_all"] = roofs
    detail["lib_stats_rank0_last_step"] = res["lib_stats_rank0_last_step"]
    detail["lib_stats_over_ranks_last_step"] = res["lib_stats_over_ranks_last_step"]
    detail["step_wall_ms_rank0"] = res["step_wall_ms_rank0"]
    detail["step_phases_rank0_dist_sweep_c1_first_replicates_hostenqueue_ms"] = res["step_phases_rank0_dist_sweep_c1_first_replicates_hostenqueue_ms"]
    detail["ceilings_measured_here"] = {"int8_tops": i8_peak, "fp64_dmma_tflops": dmma_peak, "read_only_stream_gbs": read_peak}
    if e2e_detail:
        detail["e2e"] = e2e_detail
    del W
    ctx.trim()
    torch.cuda.empty_cache()

    # ---- detail legs -------------------------------------------------------------------------------------------------
    def leg(cells, metric, sat, bootnr, seed, structure, steps=2, warmup=1, clusters=None, env=None, want_spot=False):
        old = {}
        for k_, v_ in (env or {}).items():
            old[k_] = os.environ.get(k_)
            os.environ[k_] = v_
        try:
            Wl = Workload(cells, metric, sat, bootnr, seed, structure, clusters)
            r = Wl.timed(steps, warmup, ["gram", "gram_i8", "scan_swap", "scan_build", "scan_delta", "scan_group", "compact"])
            out = {"workload": workload_name(a, cells, metric, sat, bootnr, structure), "value": r["ms_per_step"] / 1e3, "unit": "s",
                   "steps": steps, "warmup": warmup, "result_digest": digest_of(Wl.state), "n_swaps_per_step": r["n_swaps_per_step"],
                   "gpu_launches_per_step": r["launches"] / steps, "bootmean_min": float(np.min(Wl.state["bootmean"])),
                   "lib_stats_rank0_last_step": r["lib_stats_rank0_last_step"], "step_wall_ms_rank0": r["step_wall_ms_rank0"],
                   "lib_stats_over_ranks_last_step": r["lib_stats_over_ranks_last_step"],
                   "step_phases_rank0_dist_sweep_c1_first_replicates_hostenqueue_ms": r["step_phases_rank0_dist_sweep_c1_first_replicates_hostenqueue_ms"],
                   "phase_ms_last_step": {k: Wl.state.get(k) for k in ("t_dist", "t_sweep", "t_boot") if k in Wl.state}}
            rf = rooflines(r["prof"], r["ms_per_step"] * steps, cells, world)
            out["roofline"] = rf[0] if rf else None
            out["kernels"] = {x["kernel"].split(" ")[0]: {"achieved": x["achieved"], "unit": x["unit"], "frac": x["frac"],
                                                         "share_of_step": x["share_of_step"], "launches": x["launches"]} for x in rf}
            if want_spot:
                xh = Wl.x_dev.cpu().numpy() if rank == 0 else None
                dmk = Wl.step(keep=True)
                out["parity_spot"] = parity_spot(ctx, dmk, xh, a, metric, rank, rows=32, sub_cells=3000)
                dmk.free()
            del Wl
            return out
        finally:
            for k_, v_ in old.items():
                if v_ is None:
                    os.environ.pop(k_, None)
                else:
                    os.environ[k_] = v_
            ctx.trim()
            torch.cuda.empty_cache()

    if not a.no_extra:
        if a.sat:
            detail["sat_false"] = leg(N, a.metric, False, a.bootnr, a.seed, a.structure, steps=3, warmup=1)
        # ring-structured data (10 profiles mixed along a closed trajectory): BUILD is not the optimum, SWAP iterates
        detail["swap_heavy"] = leg(N, a.metric, False, min(a.bootnr, 10), a.seed + 100, "ring", steps=2, warmup=1, clusters=10)
        if world > 1:
            # the north_star's all-reduce schedule on the headline workload: every PAM pass row-sharded, SP / AQ combined
            # by ncclAllReduce, no replica
            detail["cooperative"] = leg(N, a.metric, bool(a.sat), a.bootnr, a.seed, a.structure, steps=2, warmup=1,
                                        env={"RID_NO_REPLICA": "1"})
            # BASELINE configs[4]: 250k cells x 3k genes, logpearson, row-sharded 250 GB matrix (cannot replicate)
            torch.cuda.empty_cache()
            free_b, _tot = torch.cuda.mem_get_info()
            need = 250000.0 * 250112.0 * 4.0 / world + 40e9
            fits_all = -allmax([-(1.0 if free_b > need else 0.0)])[0] > 0.5
            if fits_all:
                detail["cfg5"] = leg(250000, "logpearson", False, min(a.bootnr, 10), 1005, "clusters", steps=2, warmup=1,
                                     want_spot=not a.no_parity)
            else:
                detail["cfg5"] = {"skipped": f"250k x 250k x 4 B / {world} GPUs does not fit next to the work buffers ({free_b / 1e9:.0f} GB free)"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu = cpu_sample(a, omp=False)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step / 1e3, "unit": "s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "s8 x s8 -> s32 tensor cores, exact (distance; f64 DMMA for euclidean) + u32/u64 fixed point (PAM)", "data": "synthetic",
                "config": {"workload": workload_name(a), "seed": a.seed, "parallelism": f"row-block shard x{world}",
                           "l2": "inputs (distance matrix %.1f GB) larger than L2" % (N * N * 4 / 1e9),
                           "c1": ("clusterboot's c1 = pam(D, cln) is the sweep's PAM(cln) of the SAME step (identical call on the same "
                                  "matrix, computed once inside the timed region; RID_NO_MEMO=1 repeats it: +30 ms at cfg4)"
                                  if a.sat and a.cln <= a.clustnr else "computed by rid_clusterboot")},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "result_digest": digest, "n_swaps": res["n_swaps_per_step"], "parity_spot": spot,
                "dist_tflops": (detail.get("gram_i8") or detail.get("gram") or {}).get("tflops_algorithmic_2N2G"),
                "pam_swap_gbs": (sw["work"] / (sw["ms"] * 1e-3) / 1e9) if sw["launches"] and sw["ms"] > 0 else None, "detail": detail}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
