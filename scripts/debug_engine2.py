"""Where does engine(stem=tc) diverge from engine(stem=mma)?  Runs the block chain from both stems' outputs."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bnn_b200 as bnn
from bnn_b200 import functional as BF, fuse, workloads
from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
torch.manual_seed(0)
m = workloads.resnet18()
cfg = bnn.BConfig(BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
m = bnn.prepare_binary_model(m, cfg, ignore_layers_name=["_first_", "_last_"])
workloads.randomize_batchnorm(m, seed=1)
m = m.eval().to(DEV)
x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(2)).to(DEV)

def chain(eng, x0, bits0, tag):
    outs = []
    xx, bb = x0, bits0
    for i, plan in enumerate(eng.plans):
        nxt = eng.plans[i + 1] if i + 1 < len(eng.plans) else None
        xx, bb = eng._run_block(plan, xx, bb, nxt)
        outs.append(xx.clone())
    return outs

with torch.no_grad():
    eng = fuse.optimize(m, stem="tc")
    w = m.conv1.weight
    bn = eng.stem.bn.get()
    x_mma, b_mma = BF.stem_mma(x, BF.stem_mma_weights(w), bn)
    x_tc, b_tc = BF.stem_tc(x, eng.stem.tc_weight(), bn, guard=True)
    torch.cuda.synchronize()
    print("stem out diff per image", ((x_tc - x_mma).abs().amax(dim=(1, 2, 3))).tolist(), "bits differ", int((b_tc.bits != b_mma.bits).sum()))
    o_mma = chain(eng, x_mma, b_mma, "mma")
    o_tc = chain(eng, x_tc, b_tc, "tc")
    o_tcc = chain(eng, x_tc.clone(), BF.PackedActivations(b_tc.bits.clone(), b_tc.n, b_tc.c, b_tc.h, b_tc.w), "tc-clone")
    for i, (a, b, c) in enumerate(zip(o_mma, o_tc, o_tcc)):
        print(f"block {i}: tc vs mma per image {[('%.1e' % v) for v in (b - a).abs().amax(dim=(1, 2, 3)).tolist()]}  "
              f"tc-clone vs mma {[('%.1e' % v) for v in (c - a).abs().amax(dim=(1, 2, 3)).tolist()]}  max|a| {float(a.abs().max()):.2f}")
    y_eng = eng(x)
    y_chain = m.fc(torch.flatten(m.avgpool(o_tc[-1]), 1))
    y_mma = m.fc(torch.flatten(m.avgpool(o_mma[-1]), 1))
    print("engine vs chain(tc)", float((y_eng - y_chain).abs().max()), "engine vs chain(mma)", float((y_eng - y_mma).abs().max()),
          "chain tc vs mma", float((y_chain - y_mma).abs().max()))
    # stem out inside the engine: rerun and capture
    x2, b2 = BF.stem_tc(x.contiguous(), eng.stem.tc_weight(), eng.stem.bn.get(), nx=eng._entry_affine(eng.plans[0]), guard=True, x_log2_scale=eng._x_log2_scale)
    print("second stem call equal to first:", torch.equal(x2, x_tc), torch.equal(b2.bits, b_tc.bits))
    # where do the stem outputs differ most (image 0)?
    d = (x_tc[0] - x_mma[0]).abs()
    idx = torch.nonzero(d > 1e-4)
    print("image0 positions with |diff| > 1e-4:", idx.shape[0], idx[:10].tolist())
