"""Timeline of CTA 0 of the tcgen05 stem (clock64 stamps per role and row), for the pipeline analysis in DESIGN.md."""
import ctypes, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bnn_b200
from bnn_b200 import functional as BF, native
dev = torch.device("cuda:0"); torch.manual_seed(0)
x = torch.randn(256, 3, 224, 224, device=dev); w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
g, h = 0.5 + torch.rand(64, device=dev), 0.2 * torch.randn(64, device=dev)
ops, wls = BF.stem_tc_weights(w)
out = torch.empty((256, 64, 56, 56), device=dev).contiguous(memory_format=torch.channels_last)
bits = torch.empty((256, 1, 56, 56, 4), dtype=torch.int32, device=dev)
for dbg in (0, 15, 7):
    tl = torch.zeros((4, 512, 4), dtype=torch.int64, device=dev)
    p = native.StemTcParams()
    p.x, p.n, p.h, p.w, p.w_ops, p.w_log2_scale, p.x_log2_scale = x.data_ptr(), 256, 224, 224, ops.data_ptr(), wls, 7
    p.bn_scale, p.bn_shift, p.pool, p.out, p.out_bits, p.out_bits2 = g.data_ptr(), h.data_ptr(), 1, out.data_ptr(), bits.data_ptr(), tl.data_ptr()
    for _ in range(3):
        native.check(native.lib().bnn_stem_tc_run(ctypes.byref(p), (dbg | 16) << 8, torch.cuda.current_stream().cuda_stream), "run")
    torch.cuda.synchronize()
    t = tl.cpu().numpy()
    base = t[t > 0].min()
    conv, mma, acc, outw = t[0], t[1], t[2], t[3]
    rows = slice(40, 60)
    def d(a): return np.diff(a).astype(int).tolist()
    print(json.dumps({"dbg": dbg,
        "mma_row_period": d(mma[rows, 2]), "mma_wait_cycles": (mma[rows, 1] - mma[rows, 0]).astype(int).tolist(),
        "mma_issue_cycles": (mma[rows, 3] - mma[rows, 2]).astype(int).tolist(),
        "acc_wait_cycles": (acc[rows, 1] - acc[rows, 0]).astype(int).tolist(), "acc_work_cycles": (acc[rows, 2] - acc[rows, 1]).astype(int).tolist(),
        "acc_wake_after_mma_issue_end": (acc[rows, 1] - mma[rows, 3]).astype(int).tolist(),
        "conv_wait_cycles": (conv[rows, 1] - conv[rows, 0]).astype(int).tolist(), "conv_work_cycles": (conv[rows, 2] - conv[rows, 1]).astype(int).tolist(),
        "conv_period": d(conv[rows, 2]),
        "out_wait": (outw[20:30, 1] - outw[20:30, 0]).astype(int).tolist(), "out_work": (outw[20:30, 2] - outw[20:30, 1]).astype(int).tolist(),
        "total_cycles": int(t.max() - base)}))
