"""How often does the chaotic sign() cascade make an image's logits disagree (> 1e-3) with the reference's CPU forward?
Per stem kernel, over N images at 224x224 (the reference itself is this sensitive: see DESIGN.md)."""
import json, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from bnn_b200 import fuse
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
CFG = sys.argv[2] if len(sys.argv) > 2 else "resnet18"
RES = bench.CONFIGS[CFG]["res"]
m = bench.build_model(CFG, "basic_relu")
twin, kind, _ = bench.cpu_twin(CFG, "basic_relu", m)
x = torch.randn(N, 3, RES, RES, generator=torch.Generator().manual_seed(4321))
with torch.no_grad():
    want = twin(x)
    md = m.to("cuda:0"); xd = x.to("cuda:0")
    res = {"images": N, "reference": kind, "config": CFG}
    for stem in (("tc", "mma", "fma") if CFG != "hblock" else ("tc", "torch")):
        eng = fuse.optimize(md, stem=stem) if stem != "torch" else fuse.optimize(md, fuse_stem=False)
        y = torch.cat([eng(xd[i:i + 16]) for i in range(0, N, 16)]).cpu()
        per = (y - want).abs().amax(1) / want.abs().max()
        res[stem] = {"images_off_by_more_than_1e-3": int((per > 1e-3).sum()), "median_rel_err": float(per.median()), "max_rel_err": float(per.max())}
    y = torch.cat([md(xd[i:i + 16]) for i in range(0, N, 16)]).cpu()
    per = (y - want).abs().amax(1) / want.abs().max()
    res["per_layer_torch_stem"] = {"images_off_by_more_than_1e-3": int((per > 1e-3).sum()), "median_rel_err": float(per.median()), "max_rel_err": float(per.max())}
print(json.dumps(res))
