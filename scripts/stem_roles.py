"""Timing experiment: the tcgen05 stem with individual warp roles idled (results are garbage, only the time matters)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bnn_b200
from bnn_b200 import functional as BF
dev = torch.device("cuda:0"); torch.manual_seed(0)
x = torch.randn(256, 3, 224, 224, device=dev); w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
g, h = 0.5 + torch.rand(64, device=dev), 0.2 * torch.randn(64, device=dev)
wops = BF.stem_tc_weights(w)
def timed(fn, reps=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
res = {}
names = {0: "all roles", 1: "no accumulator work", 2: "no output work", 3: "no acc + no out", 4: "no converter work", 8: "no MMAs",
         7: "MMA issue only", 11: "converter only", 14: "accumulators only", 13: "output only", 12: "acc + out only", 9: "conv+out, no mma/acc", 15: "nothing (sync skeleton)"}
for d, nm in names.items():
    res[nm] = timed(lambda: BF.stem_tc(x, wops, (g, h), flags=d << 8))
print(json.dumps(res, indent=1))
