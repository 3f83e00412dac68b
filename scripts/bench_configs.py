"""Throughput of the other BASELINE configs on the fused engine (CUDA graph, input resident): ResNet-18 pre-activation /
PReLU (the examples/imagenet.py variant), ResNet-50 XNOR-Net++ bs128 (configs[2]), Hierarchical-Block harness bs64 at
256x256 (configs[3]).  One JSON object per workload; these are not bench.py lines (BASELINE.json quotes configs[1])."""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnn_b200 as bnn  # noqa: E402
from bnn_b200 import fuse, native, workloads  # noqa: E402
from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def cfg(post=bnn.Identity):
    return bnn.BConfig(BasicInputBinarizer, post, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))


def build(name):
    torch.manual_seed(0)
    if name == "resnet18_pre_prelu":
        m, c, bs, res = workloads.resnet18(workloads.PreBasicBlock, nn.PReLU), cfg(), 256, 224
    elif name == "resnet50_xnorpp":
        m, c, bs, res = workloads.resnet50(), cfg(BasicScaleBinarizer), 128, 224
    elif name == "hblock_net":
        m, c, bs, res = workloads.HBlockNet(), cfg(), 64, 256
    else:
        raise SystemExit(name)
    m = bnn.prepare_binary_model(m, c, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    if name == "resnet50_xnorpp":
        g = torch.Generator().manual_seed(2)
        for mod in m.modules():
            if isinstance(mod, bnn.layers.Conv2d) and hasattr(mod.activation_post_process, "alpha"):
                mod.activation_post_process.alpha.data = 0.5 + torch.rand(mod.activation_post_process.alpha.shape, generator=g)
    return m.eval().to(dev), bs, res


for name in (sys.argv[1:] or ["resnet18_pre_prelu", "resnet50_xnorpp", "hblock_net"]):
    model, bs, res = build(name)
    x = torch.randn(bs, 3, res, res, device=dev)
    out = {"workload": name, "batch": bs, "res": res}
    for label, eng in (("fused", fuse.optimize(model)), ("per_layer", model)):
        with torch.no_grad():
            for _ in range(3):
                eng(x)
            torch.cuda.synchronize()
            l0 = native.launch_count(); eng(x); launches = native.launch_count() - l0
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    y = eng(x)
            torch.cuda.current_stream().wait_stream(side)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 20
            e0.record()
            for _ in range(steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        out[label] = {"ms_per_step": ms, "images_s": bs / ms * 1e3, "native_launches": launches,
                      "engine": type(eng).__name__}
        del g
    print(json.dumps(out), flush=True)
