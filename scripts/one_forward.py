"""One steady-state forward of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum ...` (launch list of exactly one step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bnn_b200  # noqa: E402
from bnn_b200 import fuse  # noqa: E402

fused = "--no-fuse" not in sys.argv
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
config = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--config=")), "resnet18")
model = bench.build_model(config, "basic_relu").cuda()
engine = fuse.optimize(model) if fused else model
x = torch.randn(bench.CONFIGS[config]["batch"], 3, bench.CONFIGS[config]["res"], bench.CONFIGS[config]["res"], device="cuda")
with torch.no_grad():
    for _ in range(3):
        engine(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    engine(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
