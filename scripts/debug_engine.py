"""Diagnostics: where does a fused engine diverge?  stem kernels against each other, engines with each stem, per image."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bnn_b200 as bnn
from bnn_b200 import functional as BF, fuse, workloads
from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
from oracle import floatsim as fs
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"

def build(post):
    torch.manual_seed(0)
    m = workloads.resnet18()
    cfg = bnn.BConfig(BasicInputBinarizer, post, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    m = bnn.prepare_binary_model(m, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(m, seed=1)
    return m.eval()

def rel(a, b): return float((a - b).abs().max() / b.abs().max())

for n, res in ((4, 224), (2, 224), (8, 224), (4, 96)):
    m = build(BasicScaleBinarizer)
    twin = fs.mirror_model(m)
    x = torch.randn(n, 3, res, res, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = twin(x)
        md = m.to(DEV); xd = x.to(DEV)
        w = md.conv1.weight
        bn = fuse._FoldedBN(md.bn1).get()
        s_mma, b_mma = BF.stem_mma(xd, BF.stem_mma_weights(w), bn)
        s_tc, b_tc = BF.stem_tc(xd, BF.stem_tc_weights(w), bn, guard=True)
        s_tc2, _ = BF.stem_tc(xd, BF.stem_tc_weights(w), bn, guard=False)
        d = (s_tc - s_mma).abs().amax(dim=(1, 2, 3)) / s_mma.abs().max()
        print(f"n={n} res={res}: stem tc(guard) vs mma per image {d.tolist()}  tc(no guard) vs mma {rel(s_tc2, s_mma):.2e} bits differ {int((b_tc.bits != b_mma.bits).sum())}")
        for stem in ("tc", "mma", "fma"):
            for rep in range(2):
                y = fuse.optimize(md, stem=stem)(xd).cpu()
                per = ((y - want).abs().amax(dim=1) / want.abs().max()).tolist()
                print(f"   engine stem={stem} rep{rep}: rel {rel(y, want):.2e} per image {['%.1e' % v for v in per]}")
        y = md(xd).cpu()
        print(f"   per-layer: rel {rel(y, want):.2e}")
