import sys, copy, time
import torch
sys.path.insert(0, ".")
import bench
torch.set_grad_enabled(False)
m = bench.build_reference_model("resnet18", "basic_relu")
m64 = copy.deepcopy(m).double()
N = 24
x = torch.randn(N, 3, 224, 224, generator=torch.Generator().manual_seed(1000))
t = time.time()
y32 = m(x)
y64 = m64(x.double())
per = ((y32.double() - y64).abs().amax(1) / y64.abs().max())
print("fp32 reference vs its own fp64 evaluation, per image rel err:", ["%.1e" % v for v in per.tolist()], time.time() - t)
# perturb the input by 1 ulp-ish noise: reference fp32 on x*(1+1e-7)
y32b = m(x * (1 + 1e-7))
per2 = ((y32b - y32).abs().amax(1) / y32.abs().max())
print("fp32 reference on x vs x*(1+1e-7):", ["%.1e" % v for v in per2.tolist()])
