"""Diagnostic: where does the time of one compdist+clusterboot step go?  Prints wall vs device ms per call together
with the SM / memory clocks and board power sampled (NVML, 20 ms) while the step ran."""
import os, sys, time, subprocess, threading, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch

print(subprocess.run("lscpu | egrep 'Model name|^CPU\\(s\\)'; cat /proc/loadavg; nvidia-smi --query-gpu=power.limit,power.default_limit,clocks.max.sm,clocks.max.mem --format=csv",
                     shell=True, capture_output=True, text=True).stdout)
N, G, K, B = 100000, 3000, 30, 50
ctx = rid.Context(device=0); rid.set_context(ctx)
dev = torch.device("cuda", 0)
x, _ = synth_features_torch(N, G, K, 1004, dev)
torch.cuda.synchronize()
samples = rid.bootstrap_samples(N, B, 17000)

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
log = []
stop = threading.Event()
def loop():
    while not stop.is_set():
        try:
            log.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        except Exception as e:
            log.append((time.perf_counter(), -1, -1, -1, 0))
        stop.wait(0.02)
thr = threading.Thread(target=loop, daemon=True); thr.start()

def step(tag, kinds=None):
    if kinds is not None: ctx.profile(kinds)
    w0 = time.perf_counter()
    dm = ctx.dist((x.data_ptr(), G, N), "pearson"); w1 = time.perf_counter(); d_ms = ctx.last_ms
    r = dm.clusterboot(K, samples); w2 = time.perf_counter(); b_ms = ctx.last_ms
    dm.free(); ctx.sync(); w3 = time.perf_counter()
    extra = ""
    if kinds is not None:
        tot = 0.0
        for k in ("gram", "gram_i8", "scan_swap", "scan_delta", "scan_build", "scan_group", "prep", "compact"):
            p = ctx.profile_get(k); tot += p["ms"]
            extra += f" {k}={p['ms']:.0f}"
        extra = f" | kernels {tot:.0f}:" + extra
        ctx.profile(False)
    s = [r_ for r_ in log if w1 <= r_[0] <= w2]
    sm = [r_[1] for r_ in s]; mem = [r_[2] for r_ in s]; pw = [r_[3] for r_ in s]
    rs = 0
    for r_ in s: rs |= r_[4]
    clk = f"sm min/med {min(sm)}/{statistics.median(sm)} mem min {min(mem)} W mean/max {statistics.mean(pw):.0f}/{max(pw):.0f} reasons {rs:#x} n={len(s)}" if s else ""
    print(f"{tag}: dist {1e3*(w1-w0):.0f}/{d_ms:.0f} boot {1e3*(w2-w1):.0f}/{b_ms:.0f} total {1e3*(w3-w0):.0f} | {clk}{extra}", flush=True)

for i in range(3): step("warm")
for i in range(6): step("plain")
for i in range(3): step("prof", True)
stop.set()
