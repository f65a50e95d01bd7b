#!/usr/bin/env python
"""bench.py — compdist + clustexp(kmedoids) wall seconds on synthetic UMI matrices (BASELINE.json's metric).

A "step" is one pass of the hot path over one synthetic matrix: compdist(metric) -> clustexp(sat, cln, bootnr) through
the reference-facing API (raceid3_stemid2_package_b200.compdist / clustexp -> C ABI -> CUDA kernels).

  value : seconds per step with the feature matrix already resident in HBM (CUDA events on the library's stream,
          barrier + synchronize on both sides, max over ranks)
  e2e   : seconds per step starting from a HOST (pinned) feature matrix: H2D copy inside the timed region, partition
          / jaccard / medoids read back to the host
  roofline : the dominant kernel (k_scan, the PAM streaming pass of BUILD and SWAP, HBM-bound): algorithmic bytes /
          CUDA-event duration of its launches inside the timed region, against MEASURED_PEAKS.json's hbm_gbs;
          detail.roofline_all holds the same object for the compaction, incremental-SWAP and distance kernels
  cpu_baseline / --impl reference : the CPU oracle (restatement of the reference's R/C path; R itself is not installed
          in this image) on a bounded sample of the same workload, extrapolated ~ N^2

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); the distance matrix is row-block sharded and the
library runs its own NCCL all-reduces.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--sat", type=int, default=0)
    ap.add_argument("--clustnr", type=int, default=30)
    ap.add_argument("--bootnr", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1004)
    ap.add_argument("--cpu-cells", type=int, default=0,
                    help="cells in the CPU sample (0: 2500 for the 1-thread leg, ~10 s; 6000 for the all-threads reference arm)")
    ap.add_argument("--cpu-bootnr", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--timed-kinds", default="gram,gram_i8,scan_swap,scan_build,scan_delta,compact",
                    help="kernel kinds bracketed with CUDA events inside the timed region")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic {a.cells} cells x {a.genes} genes, metric={a.metric}, clustexp(sat={bool(a.sat)}, cln={a.cln}, "
            f"clustnr={a.clustnr}, bootnr={a.bootnr}, FUNcluster=kmedoids)")


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
    cores = L.orc_num_threads()
    ns = min(a.cpu_cells or (6000 if omp else 2500), a.cells)
    x, _ = synth_features_numpy(ns, a.genes, a.cln, a.seed)
    xt = np.ascontiguousarray(x.T)
    t0 = time.perf_counter()
    D, _ = orc.dist(xt, a.metric, omp=omp)
    t_dist = time.perf_counter() - t0
    Dc = np.ascontiguousarray(D.T)
    k = min(a.cln, ns - 1)
    t_sweep = 0.0
    if a.sat:
        t0 = time.perf_counter()
        out = np.zeros(min(a.clustnr, ns - 1))
        L.orc_gap_sweep(Dc, ns, None, ns, len(out), out)
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
    full = (t_dist + t_sweep + t_c1) * s2 + t_boot * s2 * (a.bootnr / max(bs, 1))
    sample = (f"oracle port ({'OpenMP ' + str(cores) + ' threads' if omp else '1 thread'}) on {ns} cells x {a.genes} genes, "
              f"k={k}, {bs} of {a.bootnr} bootstrap replicates: dist {t_dist:.2f}s, sweep {t_sweep:.2f}s, pam(c1) {t_c1:.2f}s, "
              f"boot {t_boot:.2f}s; extrapolated x(N/n)^2={s2:.0f} (distance ~N^2 G, PAM ~k N^2 per iteration with the "
              f"sample's iteration counts) and linearly in bootnr")
    return {"value": full, "unit": "s", "cores": cores, "kind": "port", "sample": sample,
            "sample_seconds": t_dist + t_sweep + t_c1 + t_boot}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    info = None
    for i in range(a.warmup + a.steps):
        info = cpu_sample(a, omp=True)
        if i >= a.warmup:
            vals.append(info["value"])
        if i == 0 and info["sample_seconds"] * (a.warmup + a.steps) > 240:
            # keep the whole run within a few minutes: one warm-up only
            a.warmup = min(a.warmup, 1)
    v = statistics.mean(vals)
    info["value"] = v
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
    """SM clock / throttle reasons DURING the timed region.  In-process NVML polling (one cheap call per 250 ms); a
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

    N, G = a.cells, a.genes
    dev = torch.device("cuda", local)
    x_dev, _ = synth_features_torch(N, G, a.cln, a.seed, dev)  # N x G, cell-major; identical on every rank
    torch.cuda.synchronize()
    x_host = None
    if not a.no_e2e:
        x_host = torch.empty((N, G), dtype=torch.float64, pin_memory=True)
        x_host.copy_(x_dev)
        torch.cuda.synchronize()
    h2d = N * G * 8

    samples = rid.bootstrap_samples(N, a.bootnr, 17000)  # R-compatible index vectors; drawn once, outside the timing
    # ceilings measured on this GPU right now (outside the timed region)
    dmma_peak = ctx.measure_dmma_peak()
    read_peak = ctx.measure_read_peak(8.0)

    state = {}

    def step(resident: bool):
        if resident:
            dm = ctx.dist((x_dev.data_ptr(), G, N), a.metric)
        else:
            dm = ctx.dist_cellmajor(x_host.numpy(), a.metric)
        t_dist = ctx.last_ms
        out = {}
        if a.sat:
            out["logW"] = dm.gap_sweep(min(a.clustnr, N - 1))
            out["t_sweep"] = ctx.last_ms
        r = dm.clusterboot(a.cln, samples)
        out["t_boot"] = ctx.last_ms
        out["t_dist"] = t_dist
        out["partition"] = r["partition"]
        out["bootmean"] = r["bootmean"]
        out["d2h"] = r["partition"].nbytes + r["bootresult"].nbytes + r["id_med"].nbytes
        dm.free()
        state.update(out)

    # ---- resident leg: W warm-up + K timed steps -------------------------------------------------------------
    for _ in range(a.warmup):
        step(True)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:  # one sampler per job: eight concurrent nvidia-smi pollers slow every rank's driver calls down
        clocks.start()
    # inside the timed region the kernels of the roofline objects are bracketed with CUDA events (~1800 pairs per step at
    # the defaults, pre-created); the per-kind breakdown of everything comes from one extra, untimed step below
    ctx.profile([k for k in a.timed_kinds.split(",") if k])
    l0 = ctx.launches
    barrier()
    ctx.event_record(0)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step(True)
    ctx.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.event_elapsed_ms(0, 1)
    launches = ctx.launches - l0
    KINDS = ("gram", "gram_i8", "scan_swap", "scan_delta", "scan_build", "scan_group", "prep", "compact")
    prof = {k: ctx.profile_get(k) for k in KINDS}
    ctx.profile(False)
    clk = clocks.stop()
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0]) / a.steps
    wall_ms_per_step = float(t[1]) / a.steps

    # ---- e2e leg: host input, host outputs -------------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        step(False)  # warm-up of the host path
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step(False)
        barrier()
        e = torch.tensor([(time.perf_counter() - t0) / a.steps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
        e2e = {"value": float(e[0]), "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(state["d2h"])}

    # ---- untimed breakdown step: every kernel kind bracketed (its events slow the step down a little; shares only) ----
    ctx.profile(True)
    ctx.event_record(2)
    step(True)
    ctx.event_record(3)
    barrier()
    breakdown_ms = ctx.event_elapsed_ms(2, 3)
    prof_all = {k: ctx.profile_get(k) for k in KINDS}
    ctx.profile(False)

    # ---- rooflines: the two hot kernels; `roofline` is the one with the larger share of the step --------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = ("measured copy bandwidth, MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks
               else "fallback 6650 GB/s (B200_PROFILING.md)")
    step_ms_total = ms_per_step * a.steps
    roofs = []
    # DRAM bytes per algorithmic byte of the captured launches (ncu --set full; profiles/ncu_traffic_ratios.json)
    try:
        ratios = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_ratios.json")))
    except Exception:
        ratios = {}

    def traffic_of(kernel, alg_bytes_per_launch):
        r = ratios.get(kernel, {}).get("dram_per_algorithmic")
        return alg_bytes_per_launch * r if r else None

    sw = prof["scan_swap"]
    # k_scan<MODE>: one streaming body, three instantiations (0: BUILD step 0, 1: SWAP / grouped sums, 2: incremental BUILD)
    sc = {f: sum(prof[k][f] for k in ("scan_swap", "scan_build", "scan_group")) for f in ("launches", "ms", "work", "traffic")}
    if sc["launches"]:
        ach = sc["work"] / (sc["ms"] * 1e-3) / 1e9
        roofs.append({"bound": "hbm", "kernel": "k_scan<MODE> (PAM streaming pass: BUILD steps, SWAP scan, grouped column sums)",
                      "achieved": ach, "peak": hbm_peak,
                      "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic_of("k_scan", sc["work"] / sc["launches"]),
                      "peak_source": hbm_src,
                      "read_only_stream_gbs_measured_here": read_peak, "launches": sc["launches"],
                      "avg_launch_ms": sc["ms"] / sc["launches"],
                      "algorithmic_bytes_per_launch": sc["work"] / sc["launches"],
                      "requested_bytes_per_launch": sc["traffic"] / sc["launches"],
                      "by_mode": {k: {"launches": prof[k]["launches"], "ms": prof[k]["ms"],
                                      "gbs": prof[k]["work"] / (prof[k]["ms"] * 1e-3) / 1e9 if prof[k]["ms"] else None}
                                  for k in ("scan_swap", "scan_build", "scan_group")},
                      "note": "bytes = rows scanned x subset size x 4, every distance read once per pass; the kernel only "
                              "reads, so the measured copy bandwidth (read + write) understates its ceiling: "
                              "read_only_stream_gbs_measured_here is the read-only streaming figure of this run; traffic = "
                              "algorithmic bytes x the DRAM/algorithmic ratio of the ncu capture",
                      "share_of_step": sc["ms"] / step_ms_total})
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
    if cp["launches"]:
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
    if gm["launches"]:
        ach = gm["work"] / (gm["ms"] * 1e-3) / 1e12
        roofs.append({"bound": "tensor", "kernel": "k_gram_dmma (FP64 DMMA distance contraction)", "achieved": ach,
                      "peak": dmma_peak, "unit": "TFLOP/s", "frac": ach / dmma_peak if dmma_peak else None,
                      "traffic": None,
                      "peak_source": "FP64 DMMA issue-rate ceiling measured in this run (rid_measure_dmma_peak); "
                                     "MEASURED_PEAKS.json holds no FP64 figure (nominal B200 FP64 tensor: 40 TFLOP/s)",
                      "launches": gm["launches"], "avg_launch_ms": gm["ms"] / gm["launches"],
                      "issued_flops_per_launch": gm["work"] / gm["launches"],
                      "algorithmic_flops_per_launch": 2.0 * N * N * G / world,
                      "note": "work = flops actually issued (upper-triangle tiles only; the mirror tile is a store)",
                      "share_of_step": gm["ms"] / step_ms_total})
    g8 = prof["gram_i8"]
    if g8["launches"]:
        ach = g8["work"] / (g8["ms"] * 1e-3) / 1e12
        bf16 = float(peaks.get("bf16_tflops", 1590.0))
        i8_peak = 2.0 * bf16  # int8 tcgen05 issues at twice the bf16 rate; the bf16 figure is the measured cuBLAS burst
        roofs.append({"bound": "tensor", "kernel": "k_gram_i8_persist (tcgen05.mma kind::i8, exact Ozaki-split distance contraction)",
                      "achieved": ach, "peak": i8_peak, "unit": "TOP/s", "frac": ach / i8_peak,
                      "traffic": (ratios.get("k_gram_i8_persist", {}).get("dram_bytes_per_launch_cfg4")
                                  if (N, G, world) == (100000, 3000, 1) else None),
                      "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (measured cuBLAS bf16 burst; int8 dense is 2x bf16 "
                                     "on sm_100: nominal 4500 TOP/s)" if "bf16_tflops" in peaks else
                                     "2 x fallback 1590 TFLOP/s bf16 (B200_PROFILING.md)",
                      "launches": g8["launches"], "avg_launch_ms": g8["ms"] / g8["launches"],
                      "issued_int8_ops_per_launch": g8["work"] / g8["launches"],
                      "fp64_equivalent_tflops_2N2G": 2.0 * N * N * G * g8["launches"] / (g8["ms"] * 1e-3) / 1e12,
                      "note": "work = 13 digit products x 2*128*96*Kp per computed tile (tiles touching the upper triangle)",
                      "share_of_step": g8["ms"] / step_ms_total})
    roofs.sort(key=lambda r: -r["share_of_step"])
    roof = roofs[0] if roofs else None
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
                                                  if prof[k]["launches"] else None)
        elif k == "gram":
            d["tflops_issued"] = v["work"] / (v["ms"] * 1e-3) / 1e12
            d["tflops_algorithmic_2N2G"] = 2.0 * N * N * G / world * v["launches"] / (v["ms"] * 1e-3) / 1e12
        else:
            d["algorithmic_gbs"] = v["work"] / (v["ms"] * 1e-3) / 1e9
            d["requested_gbs"] = v["traffic"] / (v["ms"] * 1e-3) / 1e9
        detail[k] = d
    detail["phase_ms_last_step"] = {k: state.get(k) for k in ("t_dist", "t_sweep", "t_boot") if k in state}
    detail["wall_ms_per_step"] = wall_ms_per_step
    detail["clusters_found"] = int(state["partition"].max())
    detail["bootmean_min"] = float(np.min(state["bootmean"]))
    detail["roofline_all"] = roofs

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu = cpu_sample(a, omp=False)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step / 1e3, "unit": "s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "s8 x s8 -> s32 tensor cores, exact (distance; f64 DMMA for euclidean) + u32/u64 fixed point (PAM)", "data": "synthetic",
                "config": {"workload": workload_name(a), "seed": a.seed, "parallelism": f"row-block shard x{world}",
                           "l2": "inputs (distance matrix %.1f GB) larger than L2" % (N * N * 4 / 1e9)},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "dist_tflops": (detail.get("gram_i8") or detail.get("gram") or {}).get("tflops_algorithmic_2N2G"),
                "pam_swap_gbs": (sw["work"] / (sw["ms"] * 1e-3) / 1e9) if sw["launches"] else None, "detail": detail}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
