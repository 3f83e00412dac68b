"""Workload definitions for the benchmark / parity configurations of BASELINE.json.

These are the *float* architectures the reference converts (reference bnn/models/resnet.py,
bnn/models/layers/res_block.py, hierarchical_block.py) re-stated as table-driven builders with
the same module names, construction order and initialisation, so that (a) ``state_dict`` keys are
interchangeable and (b) ``torch.manual_seed(s)`` yields bit-identical parameters to the
reference's constructors -- golden logits generated from the real reference in the build
container (tests/golden/make_golden.py) can then be checked on a box where the reference is
not installed.  They are inputs to ``prepare_binary_model``; they contain no binarized code.
"""
from typing import Callable, List, Optional

import torch
import torch.nn as nn


def _conv3x3(cin: int, cout: int, stride: int = 1) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


def _conv1x1(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, 1, bias=False)


def _act(activation, channels: int) -> nn.Module:
    return activation(inplace=True) if activation is nn.ReLU else activation(num_parameters=channels)


class BasicBlock(nn.Module):
    """conv-bn-act, conv-bn, (+ identity), act  (reference res_block.py:8-56)."""
    expansion = 1
    pre_activation = False

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=nn.BatchNorm2d, activation=nn.ReLU):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride)
        self.bn1 = norm(inplanes if self.pre_activation else planes)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = norm(planes)
        self.act1 = _act(activation, planes)
        self.act2 = _act(activation, planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.act1(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y += shortcut
        return self.act2(y)


class PreBasicBlock(BasicBlock):
    """bn-conv-act twice, then + identity  (reference res_block.py:121-167)."""
    pre_activation = True

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.act1(self.conv1(self.bn1(x)))
        y = self.act2(self.conv2(self.bn2(y)))
        y += shortcut
        return y


class Bottleneck(nn.Module):
    """1x1 - 3x3(stride) - 1x1 x4  (reference res_block.py:59-118)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=nn.BatchNorm2d, activation=nn.ReLU):
        super().__init__()
        width = planes
        self.conv1 = _conv1x1(inplanes, width)
        self.bn1 = norm(width)
        self.conv2 = _conv3x3(width, width, stride)
        self.bn2 = norm(width)
        self.conv3 = _conv1x1(width, planes * 4)
        self.bn3 = norm(planes * 4)
        self.act1 = _act(activation, width)
        self.act2 = _act(activation, width)
        self.act3 = _act(activation, planes * 4)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.act1(self.bn1(self.conv1(x)))
        y = self.act2(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        y += shortcut
        return self.act3(y)


class PreBottleneck(Bottleneck):
    """bn-conv-act three times, then + identity (reference res_block.py:170-229)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm=nn.BatchNorm2d, activation=nn.ReLU):
        super().__init__(inplanes, planes, stride, downsample, norm, activation)
        self.bn1 = norm(inplanes)
        self.bn3 = norm(planes)

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.act1(self.conv1(self.bn1(x)))
        y = self.act2(self.conv2(self.bn2(y)))
        y = self.act3(self.conv3(self.bn3(y)))
        y += shortcut
        return y


class HBlock(nn.Module):
    """Hierarchical block: three bn-act-conv stages whose outputs are concatenated
    (reference hierarchical_block.py:8-60)."""

    def __init__(self, inplanes, planes, downsample=None, norm=nn.BatchNorm2d, activation=nn.ReLU):
        super().__init__()
        half, quarter = planes // 2, planes // 4
        self.bn1 = norm(inplanes)
        self.conv1 = _conv3x3(inplanes, half)
        self.bn2 = norm(half)
        self.conv2 = _conv3x3(half, quarter)
        self.bn3 = norm(quarter)
        self.conv3 = _conv3x3(quarter, quarter)
        self.act1 = _act(activation, half)
        self.act2 = _act(activation, half)
        self.act3 = _act(activation, quarter)
        self.downsample = downsample

    def forward(self, x):
        o1 = self.conv1(self.act1(self.bn1(x)))
        o2 = self.conv2(self.act2(self.bn2(o1)))
        o3 = self.conv3(self.act3(self.bn3(o2)))
        y = torch.cat((o1, o2, o3), 1)
        y += x if self.downsample is None else self.downsample(x)
        return y


def _reference_init(model: nn.Module) -> None:
    # reference resnet.py:103-108
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class ResNet(nn.Module):
    """ImageNet ResNet skeleton (reference resnet.py:50-167): 7x7/2 stem, four stages, avg-pool
    + 1x1 shortcut convs, global pool, fc.  ``fc_in`` exists because upstream wires ``fc`` to the
    last stage's *planes* (resnet.py:101,143) which is only right for expansion-1 blocks."""

    def __init__(self, block, depths: List[int], num_classes: int = 1000, activation=None,
                 norm: Callable[..., nn.Module] = None, fc_in: Optional[int] = None):
        super().__init__()
        norm = norm or nn.BatchNorm2d
        activation = activation or nn.ReLU
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        stages = []
        for i, (planes, depth) in enumerate(zip((64, 128, 256, 512), depths)):
            stages.append(self._stage(block, planes, depth, 1 if i == 0 else 2, norm, activation))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, num_classes)
        _reference_init(self)
        if fc_in is not None and fc_in != 512:      # the harness patch, applied in the harness's order
            self.fc = nn.Linear(fc_in, num_classes)

    def _stage(self, block, planes, depth, stride, norm, activation) -> nn.Sequential:
        shortcut = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            shortcut = nn.Sequential(
                nn.AvgPool2d(kernel_size=stride, stride=stride, ceil_mode=True, count_include_pad=False),
                _conv1x1(self.inplanes, planes * block.expansion),
                norm(planes * block.expansion))
        blocks = [block(self.inplanes, planes, stride, shortcut, norm=norm, activation=activation)]
        self.inplanes = planes * block.expansion
        blocks += [block(self.inplanes, planes, norm=norm, activation=activation) for _ in range(1, depth)]
        return nn.Sequential(*blocks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(block=None, activation=None, **kw) -> ResNet:
    """reference resnet.py:180-189; ``block=PreBasicBlock, activation=nn.PReLU`` is the variant
    trained by examples/imagenet.py:153."""
    return ResNet(block or BasicBlock, [2, 2, 2, 2], activation=activation, **kw)


def resnet34(block=None, activation=None, **kw) -> ResNet:
    return ResNet(block or BasicBlock, [3, 4, 6, 3], activation=activation, **kw)


def resnet50(activation=None, **kw) -> ResNet:
    """Bottleneck ResNet-50 with ``fc`` on 2048 features.  Upstream's ``resnet50()`` builds
    ``fc = Linear(512, ...)`` and cannot run forward; harnesses patch
    ``model.fc = nn.Linear(2048, 1000)`` after construction; ``fc_in=2048`` applies the same patch at
    the same point of the RNG stream, so seeded parameters match the patched reference model."""
    return ResNet(Bottleneck, [3, 4, 6, 3], activation=activation, fc_in=2048, **kw)


class HBlockNet(nn.Module):
    """BASELINE config 4 harness (no such top-level model upstream):
    fp stem to 64 ch at /2, HBlock(64->256) with a BN-ReLU-1x1 shortcut, 2x2 avg-pool,
    4 x HBlock(256->256), global pool, fc."""

    def __init__(self, num_classes: int = 1000, depth: int = 4, hblock: Optional[Callable[..., nn.Module]] = None):
        """``hblock(inplanes, planes, downsample)`` builds one block; the default is this module's ``HBlock``.  The
        golden generator and bench.py's reference arm pass the reference's own class here
        (``bnn.models.layers.HBlock(i, p, downsample=d, norm_layer=nn.BatchNorm2d)``): module names, construction
        order and therefore seeded parameters are the same either way."""
        super().__init__()
        hblock = hblock or (lambda i, p, d: HBlock(i, p, downsample=d))
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        shortcut = nn.Sequential(nn.BatchNorm2d(64), nn.ReLU(inplace=True), _conv1x1(64, 256))
        self.block0 = hblock(64, 256, shortcut)
        self.pool = nn.AvgPool2d(2)
        self.blocks = nn.Sequential(*[hblock(256, 256, None) for _ in range(depth)])
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(256, num_classes)
        _reference_init(self)

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.blocks(self.pool(self.block0(x)))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def randomize_batchnorm(model: nn.Module, seed: int = 1) -> nn.Module:
    """Non-degenerate BN statistics/affine: with default BN the reference
    disagrees with its own fp64 evaluation by ~20 % because sign() amplifies rounding residue of
    exact-zero dot products; with these it agrees to ~2e-7.  Also randomises learned post scales."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(0.2 * torch.randn(n, generator=g))
            m.running_var.copy_(0.5 + 1.5 * torch.rand(n, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(n, generator=g))
            m.bias.data.copy_(0.2 * torch.randn(n, generator=g))
    for m in model.modules():
        post = getattr(m, "activation_post_process", None)
        if post is not None and hasattr(post, "alpha"):
            post.alpha.data.copy_(0.5 + torch.rand(post.alpha.shape, generator=g))
    return model
