"""One steady-state fused forward of a BASELINE config between cudaProfilerStart/Stop (for ncu launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv, name = sys.argv[:1], sys.argv[1]
import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("bench_configs_mod", os.path.join(os.path.dirname(__file__), "bench_configs.py"))
src = open(spec.origin).read().split("\nfor name in (sys.argv[1:]")[0]
ns = {"__name__": "bench_configs_mod", "__file__": spec.origin}
exec(compile(src, spec.origin, "exec"), ns)
model, bs, res = ns["build"](name)
engine = ns["fuse"].optimize(model)
x = torch.randn(bs, 3, res, res, device="cuda")
with torch.no_grad():
    for _ in range(3):
        engine(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    engine(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
