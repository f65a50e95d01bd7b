"""One compdist at cfg4, ONE bootstrap-sized PAM run (k = 4), a 2-replicate clusterboot and a 3-k sweep: every hot kernel
appears within a few dozen launches (k_gram_i8_persist, k_compact_sym<0|1>, k_colsum_batch, k_scan<0|1|2>, k_scan_delta,
k_wk_batch), which is what the ncu capture needs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G = 100000, 3000
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, 30, 1004, torch.device("cuda", 0)); torch.cuda.synchronize()
dm = ctx.dist((x.data_ptr(), G, N), "pearson")
samples = rid.bootstrap_samples(N, 2, 17000)
r = dm.pam(4, subset=samples[0])
print("dist ms", ctx.last_ms, "pam medoids", r["id_med"], "swaps", r.get("n_swaps"))
cb = dm.clusterboot(4, samples)
print("clusterboot ms", ctx.last_ms, cb["bootmean"])
lw = dm.gap_sweep(3)
print("sweep ms", ctx.last_ms, lw)
dm.free()
