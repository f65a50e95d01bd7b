"""One compdist at cfg4 plus ONE bootstrap-sized PAM run (k = 4): every hot kernel appears within a handful of launches
(k_gram_i8_persist, k_compact_sym, k_scan<2> BUILD, k_scan<1> SWAP, k_scan_delta), which is what the ncu capture needs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G = 100000, 3000
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, 30, 1004, torch.device("cuda", 0)); torch.cuda.synchronize()
dm = ctx.dist((x.data_ptr(), G, N), "pearson")
sub = rid.bootstrap_samples(N, 1, 17000)[0]
r = dm.pam(4, subset=sub)
print("dist ms", ctx.last_ms, "pam medoids", r["id_med"], "swaps", r.get("n_swaps"))
dm.free()
