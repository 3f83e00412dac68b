import sys, copy, time
import torch
sys.path.insert(0, ".")
import bench
torch.set_grad_enabled(False)
m = bench.build_reference_model("resnet18", "basic_relu")
N = 32
x = torch.randn(N, 3, 224, 224, generator=torch.Generator().manual_seed(1000))
x0 = m.maxpool(m.relu(m.bn1(m.conv1(x))))
def rest(t):
    t = m.layer4(m.layer3(m.layer2(m.layer1(t))))
    return m.fc(torch.flatten(m.avgpool(t), 1))
y = rest(x0)
scale = float(x0.abs().max())
g = torch.Generator().manual_seed(5)
for delta in (1e-8, 1e-7, 3e-7, 1e-6, 3e-6):
    noise = torch.randn(x0.shape, generator=g) * (delta * scale)
    xp = torch.where(x0 > 0, x0 + noise, x0)          # keep exact ReLU zeros (a kernel reproduces those exactly)
    yp = rest(xp)
    per = ((yp - y).abs().amax(1) / y.abs().max())
    print(f"stem-output noise {delta:.0e} x max|x| (rms): images with logits change > 1e-3: {int((per > 1e-3).sum())} / {N}; max {float(per.max()):.1e}")
