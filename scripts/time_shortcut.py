"""Time the one-kernel shortcut against the two-launch form on the three ResNet-18 shapes (bs 256) and, with
--r50, the four ResNet-50 shapes (bs 128)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bnn_b200  # noqa: E402
from bnn_b200 import functional as BF  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timed(fn, reps=200):
    """back-to-back launches between two events (a single sub-50-us launch is shorter than its host-side launch cost)"""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [(256, 64, 56, 128, 2), (256, 128, 28, 256, 2), (256, 256, 14, 512, 2)]
if "--r50" in sys.argv:
    SHAPES = [(128, 64, 56, 256, 1), (128, 256, 56, 512, 2), (128, 512, 28, 1024, 2), (128, 1024, 14, 2048, 2)]
for (bs, c, hw, co, pool) in SHAPES:
    x = torch.randn(bs, c, hw, hw, device=dev).contiguous(memory_format=torch.channels_last)
    wts = BF.pack_weights(torch.randn(co, c, 1, 1, device=dev), True, True)
    bn = (0.5 + torch.rand(co, device=dev), torch.randn(co, device=dev))

    def two():
        p = BF.pack_activations(x, pool=pool, ceil_mode=True)
        return BF.bconv2d_fused(p, wts, bn=bn, channels_last=True)[0]

    def one():
        return BF.shortcut(x, wts, pool, True, bn=bn)

    assert torch.equal(one(), two())
    byts = 4 * x.numel() + 4 * bs * co * (hw // pool) ** 2
    t1, t2 = timed(one), timed(two)
    print(json.dumps({"c_in": c, "hw": hw, "c_out": co, "one_kernel_ms": t1, "two_launch_ms": t2,
                      "one_kernel_gb_s": byts / t1 * 1e-6, "algorithmic_mb": byts / 1e6}))
