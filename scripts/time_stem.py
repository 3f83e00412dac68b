"""Time the two stem kernels (bs 256, 224x224) with CUDA events and compare them against a float64 stem."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bnn_b200  # noqa: E402
from bnn_b200 import functional as BF  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
bs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
x = torch.randn(bs, 3, 224, 224, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
g, h = 0.5 + torch.rand(64, device=dev), 0.2 * torch.randn(64, device=dev)
w_t = BF.stem_weight_layout(w)
wfrag = BF.stem_mma_weights(w)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


wops = BF.stem_tc_weights(w)
amax_buf = BF.amax(x)
res = {"bs": bs, "fma_ms": timed(lambda: BF.stem(x, w_t, (g, h))), "mma_ms": timed(lambda: BF.stem_mma(x, wfrag, (g, h))),
       "mma_chain_ms": timed(lambda: BF.stem_mma(x, wfrag, (g, h), flags=16)),
       "tc_ms": timed(lambda: BF.stem_tc(x, wops, (g, h))), "tc_guarded_ms": timed(lambda: BF.stem_tc(x, wops, (g, h), guard=True)),
       "tc_amax_given_ms": timed(lambda: BF.stem_tc(x, wops, (g, h), guard=amax_buf)),
       "amax_ms": timed(lambda: BF.amax(x, out=amax_buf))}
if "--tc-only" in sys.argv:
    print(json.dumps(res)); sys.exit(0)
xs = x[:4]
y64 = torch.nn.functional.conv2d(xs.double(), w.double(), stride=2, padding=3)
y64 = torch.nn.functional.max_pool2d(torch.relu(y64 * g.double().view(1, -1, 1, 1) + h.double().view(1, -1, 1, 1)), 3, 2, 1)
a, abits = BF.stem(xs, w_t, (g, h))
b, bbits = BF.stem_mma(xs, wfrag, (g, h))
sc = float(y64.abs().max())
c, _ = BF.stem_mma(xs, wfrag, (g, h), flags=16)
res["mma_chain_err"] = float((c.double() - y64).abs().max()) / sc
res["mma_chain_rms"] = float((c.double() - y64).pow(2).mean().sqrt()) / sc
res["fma_err"] = float((a.double() - y64).abs().max()) / sc
res["mma_err"] = float((b.double() - y64).abs().max()) / sc
res["fma_rms"] = float((a.double() - y64).pow(2).mean().sqrt()) / sc
res["mma_rms"] = float((b.double() - y64).pow(2).mean().sqrt()) / sc
t, tbits = BF.stem_tc(xs, wops, (g, h), guard=True)
res["tc_err"] = float((t.double() - y64).abs().max()) / sc
res["tc_rms"] = float((t.double() - y64).pow(2).mean().sqrt()) / sc
res["plane_bit_mismatches"] = int((abits.bits ^ bbits.bits).to(torch.int64).bitwise_and(0xffffffff).ne(0).sum())
res["plane_words"] = int(abits.bits.numel())
print(json.dumps(res))
