"""Cross-module fusion for inference.

``optimize(model)`` wraps a prepared ResNet-style model (``bnn.models.resnet`` layout: ``conv1,
bn1, relu, maxpool, layer1..4, avgpool, fc``; blocks of the reference's ``BasicBlock`` /
``PreBasicBlock`` / ``Bottleneck`` / ``PreBottleneck`` shape, bnn/models/layers/res_block.py:8-229) in an
engine that runs each residual block as one kernel launch per binarized conv (plus one for a shortcut):

* the eval-mode BatchNorm after (or before) a binarized conv, the ReLU / PReLU, the residual add
  and the *next* layer's sign() are folded into the conv kernel's epilogue (``bnn_bconv2d_fused_fwd``);
* a conv whose output only feeds another binarized conv never writes fp32 at all -- it emits the
  next layer's sign/mask planes directly (lanes <-> channels makes that one ``ballot`` per word);
* the residual stream stays fp32 and is read once / written once per block; between fused blocks it is
  kept in torch's channels_last (NHWC) memory format, which matches the kernel's lanes <-> channels mapping;
* the shortcut (AvgPool2d -> sign -> conv1x1 -> BatchNorm) is one kernel (``bnn_shortcut_fwd``): the pooled
  planes never leave shared memory (other shortcut shapes: ``bnn_avgpool_pack_f32`` + the fused conv).

Per-channel BatchNorm constants are folded once (``g = weight / sqrt(var + eps)``, ``h = bias -
mean * g``) and re-folded when any of the BatchNorm tensors changes (version counters).
Blocks the engine does not recognise run through their own ``forward`` (per-layer kernels), so
the engine is always a drop-in for ``model`` in eval mode.  Results agree with the unfused path
to fp32 rounding of the BatchNorm fold (tests/test_gpu_fused.py).  The fp32 stem (conv7x7/2 + BN + ReLU +
max-pool, reference resnet.py:85-92) is one kernel too (``bnn_stem_mma_fwd`` by default, ``bnn_stem_fwd`` with
``stem="fma"``) when it has the reference's shape;
global pooling and the classifier stay torch ops.
"""
import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import functional as BF
from . import native, runtime
from .layers import Conv2d
from .layers.binary_modules import NotLowerable, _pair


class _FoldedBN:
    """(scale, shift) of an eval-mode BatchNorm2d, cached against in-place updates."""

    def __init__(self, bn: nn.BatchNorm2d) -> None:
        self.bn, self.key, self.pair = bn, None, None

    def get(self) -> Tuple[torch.Tensor, torch.Tensor]:
        bn = self.bn
        tensors = (bn.running_mean, bn.running_var, bn.weight, bn.bias)
        key = tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors) + (float(bn.eps),)
        if key != self.key:
            with torch.no_grad():
                g = torch.rsqrt(bn.running_var + bn.eps)
                if bn.weight is not None:
                    g = g * bn.weight
                h = -bn.running_mean * g
                if bn.bias is not None:
                    h = h + bn.bias
                self.pair = (g.float().contiguous(), h.float().contiguous())
            self.key = key
        return self.pair


def _activation_spec(mod: nn.Module):
    if isinstance(mod, nn.ReLU):
        return native.ACT_RELU, None
    if isinstance(mod, nn.PReLU):
        return native.ACT_PRELU, mod
    if isinstance(mod, nn.Identity):
        return native.ACT_NONE, None
    return None


def _fusable_conv(conv) -> bool:
    if not isinstance(conv, Conv2d) or conv._is_float_layer():
        return False
    if conv.groups != 1:
        # grouped convs (ResNeXt-style ``groups`` kwarg, BATS cells) run as per-group launches of the per-layer path
        # (Conv2d._forward_grouped); a block holding one is not fused and keeps its own forward
        return False
    try:
        low = conv._lowering()
    except NotLowerable:
        return False
    if conv.weight.is_cuda and BF.needs_general_path(conv._packed_weights(low), _pair(conv.stride), _pair(conv.dilation)):
        return False           # ternary weights / split-K reductions run the per-layer general path
    return (not low.has_post) or low.fused_post


_BLOCK_LAYOUTS = {
    # block class name -> (pre-activation?, ((conv, bn, act), ...)) in execution order
    "BasicBlock": (False, (("conv1", "bn1", "act1"), ("conv2", "bn2", "act2"))),
    "PreBasicBlock": (True, (("conv1", "bn1", "act1"), ("conv2", "bn2", "act2"))),
    "Bottleneck": (False, (("conv1", "bn1", "act1"), ("conv2", "bn2", "act2"), ("conv3", "bn3", "act3"))),
    "PreBottleneck": (True, (("conv1", "bn1", "act1"), ("conv2", "bn2", "act2"), ("conv3", "bn3", "act3"))),
}


class _BlockPlan:
    """One residual block of the reference's zoo (bnn/models/layers/res_block.py): a chain of binarized convs,
    each with a BatchNorm behind it (post-activation blocks: conv-bn-act ... conv-bn, +identity, act) or in front
    of it (pre-activation blocks: bn-conv-act ... , +identity), and an optional AvgPool-conv1x1-BN shortcut."""

    def __init__(self, block: nn.Module) -> None:
        self.block = block
        self.kind = None
        layout = _BLOCK_LAYOUTS.get(type(block).__name__)
        if layout is None:
            return
        pre, stages = layout
        if not all(hasattr(block, k) for stage in stages for k in stage):
            return
        ok = True
        for conv, bn, act in stages:
            b = getattr(block, bn)
            ok = ok and _fusable_conv(getattr(block, conv)) and isinstance(b, nn.BatchNorm2d) and b.track_running_stats
            ok = ok and _activation_spec(getattr(block, act)) is not None
        if ok and pre:
            # a pre-activation block normalises the INPUT of conv i with bn i: channel counts must line up
            ok = all(getattr(block, bn).num_features == getattr(block, conv).in_channels for conv, bn, _ in stages)
        ds = getattr(block, "downsample", None)
        self.shortcut = None
        if ok and ds is not None:
            # reference resnet.py:129-133: AvgPool2d(k=stride) -> binarized conv1x1 -> BatchNorm
            good = (isinstance(ds, nn.Sequential) and len(ds) == 3 and isinstance(ds[0], nn.AvgPool2d)
                    and _fusable_conv(ds[1]) and isinstance(ds[2], nn.BatchNorm2d))
            if good:
                pool = ds[0]
                k, st = _pair(pool.kernel_size), _pair(pool.stride)
                good = (k[0] == k[1] == st[0] == st[1] and _pair(pool.padding) == (0, 0)
                        and not pool.count_include_pad and pool.divisor_override is None)
            ok = good
            if good:
                self.shortcut = (ds[0], ds[1], _FoldedBN(ds[2]))
        if ok:
            self.kind = "pre" if pre else "basic"
            self.stages = [(getattr(block, c), _FoldedBN(getattr(block, b)), getattr(block, a)) for c, b, a in stages]
            self.bn1 = self.stages[0][1]

    @property
    def fused(self) -> bool:
        return self.kind is not None


def _slope(act: nn.Module, channels: int, dev) -> Optional[torch.Tensor]:
    if not isinstance(act, nn.PReLU):
        return None
    w = act.weight.detach()
    return w if w.numel() == channels else w.expand(channels).contiguous()


def _conv_args(conv: Conv2d):
    low = conv._lowering()
    return dict(bias=conv._bias(), post=conv._post_scale(low), stride=_pair(conv.stride),
                padding=conv._resolved_padding(), dilation=_pair(conv.dilation), use_alpha=low.compute_alpha,
                flags=runtime.kernel_flags()), conv._packed_weights(low)


class _StemPlan:
    """conv 7x7/2/3 (3->64, fp32, no bias) + BatchNorm + ReLU + MaxPool 3/2/1 (reference resnet.py:85-92)."""

    def __init__(self, model: nn.Module, need_pool: bool = True) -> None:
        self.ok = False
        conv, bn, pool = getattr(model, "conv1", None), getattr(model, "bn1", None), getattr(model, "maxpool", None)
        if type(conv) is not nn.Conv2d or not isinstance(bn, nn.BatchNorm2d):
            return
        if need_pool and not isinstance(pool, nn.MaxPool2d):
            return
        if not isinstance(getattr(model, "relu", None), nn.ReLU) or getattr(model, "stem_type", "basic") != "basic":
            return
        good = (conv.in_channels == 3 and conv.out_channels == 64 and conv.kernel_size == (7, 7) and conv.stride == (2, 2)
                and conv.padding == (3, 3) and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None
                and conv.padding_mode == "zeros" and bn.track_running_stats)
        if need_pool:
            good = good and (_pair(pool.kernel_size) == (3, 3) and _pair(pool.stride) == (2, 2) and _pair(pool.padding) == (1, 1)
                             and _pair(pool.dilation) == (1, 1) and not pool.ceil_mode)
        if good:
            self.ok, self.conv, self.bn = True, conv, _FoldedBN(bn)
            self.key, self.w_t = None, None
            self.mma_key, self.mma_w = None, None
            self.tc_key, self.tc_w = None, None

    def weight(self) -> torch.Tensor:
        w = self.conv.weight
        key = (w.data_ptr(), w._version)
        if key != self.key:
            # [3,7,7,32,2]: lane l holds channels (l, l+32) as one 8-byte pair for the packed-fma stem kernel
            self.w_t = w.detach().permute(1, 2, 3, 0).reshape(3, 7, 7, 2, 32).transpose(3, 4).contiguous()
            self.key = key
        return self.w_t

    def mma_weight(self):
        """(fragment buffer, log2 scale) for the mma.sync stem kernel, cached per weight version."""
        w = self.conv.weight
        key = (w.data_ptr(), w._version)
        if key != self.mma_key:
            self.mma_w = BF.stem_mma_weights(w)
            self.mma_key = key
        return self.mma_w

    def tc_weight(self):
        """(B-operand image, log2 scale) for the tcgen05 stem kernel, cached per weight version."""
        w = self.conv.weight
        key = (w.data_ptr(), w._version)
        if key != self.tc_key:
            self.tc_w = BF.stem_tc_weights(w)
            self.tc_key = key
        return self.tc_w


class FusedResNet(nn.Module):
    """Inference engine over a prepared ResNet; same call signature as the wrapped model."""

    def __init__(self, model: nn.Module, fuse_stem: bool = True, stem: str = "auto", input_range: float = None,
                 overlap_shortcuts: bool = True) -> None:
        super().__init__()
        self.stem_kernel_used = "torch"
        # the down-sampling shortcut of a block depends only on the block input: it is launched on a second stream and
        # joined before the conv that adds it, so its (HBM / latency bound) CTAs fill the SM slots that the (POPC bound)
        # conv1 / conv2 launches of the same block leave idle in their ramps and tails.  Inside a CUDA-graph capture the
        # fork / join become graph edges.
        self.overlap_shortcuts = bool(overlap_shortcuts)
        self._side_streams = {}
        self._shapes_seen = set()       # the first forward of an input shape autotunes tile plans: no concurrency then
        self._overlap_now = False
        if stem not in ("auto", "tc", "mma", "fma"):
            raise ValueError("stem must be 'auto', 'tc' (tcgen05, split fp16), 'mma' (mma.sync, split fp16) or 'fma' "
                             f"(fp32 fma chain), got {stem!r}")
        self.model = model
        self.stem_kernel = "tc" if stem == "auto" else stem
        # split-fp16 stems: None = guarded (max|x| is measured on the device every forward, any magnitude is safe);
        # a number = the caller's bound on |x| (e.g. 3.0 for normalised images): fixed scale, no measuring pass
        if input_range is not None and not (input_range > 0 and math.isfinite(input_range)):
            raise ValueError(f"input_range must be a positive finite bound on |x|, got {input_range!r}")
        self.input_range = input_range
        self._x_log2_scale = (BF.STEM_X_LOG2_SCALE if input_range is None
                              else max(-60, min(60, 15 - math.frexp(float(input_range))[1])))
        blocks: List[nn.Module] = []
        for name in ("layer1", "layer2", "layer3", "layer4"):
            blocks += list(getattr(model, name))
        self.plans = [_BlockPlan(b) for b in blocks]
        self.stem = _StemPlan(model) if fuse_stem else None

    @property
    def fused_blocks(self) -> int:
        return sum(p.fused for p in self.plans)

    def invalidate_caches(self) -> None:
        """Forget folded BatchNorm constants and stem operands (see ``bnn_b200.invalidate``)."""
        for p in self.plans:
            if p.fused:
                for _, bn, _ in p.stages:
                    bn.key = None
                if p.shortcut is not None:
                    p.shortcut[2].key = None
        if self.stem is not None and self.stem.ok:
            self.stem.bn.key = self.stem.key = self.stem.mma_key = None
            self.stem.tc_key = None

    # what the NEXT block needs in front of its first sign(): its bn1 if it is pre-activation
    @staticmethod
    def _entry_affine(plan: Optional[_BlockPlan]):
        if plan is None or not plan.fused:
            return None
        return plan.bn1.get() if plan.kind == "pre" else None

    def _run_block(self, plan: _BlockPlan, x: torch.Tensor, xbits, nxt: Optional[_BlockPlan]):
        blk = plan.block
        want_next_bits = nxt is not None and nxt.fused
        nx = self._entry_affine(nxt)
        if not plan.fused:
            y = blk(x)
            bits = BF.pack_activations(y, pre=nx) if want_next_bits else None
            return y, bits
        if xbits is None:
            xbits = BF.pack_activations(x, pre=self._entry_affine(plan))
        join = None
        # shortcut branch
        if plan.shortcut is not None:
            pool, conv_d, bn_d = plan.shortcut
            kw, wts = _conv_args(conv_d)
            one_kernel = (wts.kh == 1 and wts.kw == 1 and kw["stride"] == (1, 1) and kw["padding"] == (0, 0)
                          and conv_d.groups == 1 and x.stride(1) == 1 and wts.c_in <= min(1024, runtime.shortcut_max_cin())
                          and wts.c_out <= 4096)
            if one_kernel:
                # pool + sign + conv1x1 + BN without the planes ever reaching HBM
                bn_pair = bn_d.get()
                if self.overlap_shortcuts and self._overlap_now and x.is_cuda:
                    main = torch.cuda.current_stream(x.device)
                    side = self._side_streams.get(x.device)
                    if side is None:
                        side = self._side_streams[x.device] = torch.cuda.Stream(x.device)
                    k = _pair(pool.kernel_size)[0]
                    out = torch.empty(BF.shortcut_out_shape(x, wts, k, pool.ceil_mode), dtype=torch.float32, device=x.device,
                                      memory_format=torch.channels_last)         # allocated on the main stream
                    side.wait_stream(main)                                         # fork: the block input is complete
                    with torch.cuda.stream(side):
                        shortcut = BF.shortcut(x, wts, k, pool.ceil_mode, bias=kw["bias"], post=kw["post"], bn=bn_pair,
                                               use_alpha=kw["use_alpha"], out=out)
                    join = (main, side)
                else:
                    shortcut = BF.shortcut(x, wts, _pair(pool.kernel_size)[0], pool.ceil_mode, bias=kw["bias"],
                                           post=kw["post"], bn=bn_pair, use_alpha=kw["use_alpha"])
            else:
                pooled = BF.pack_activations(x, pool=_pair(pool.kernel_size)[0], ceil_mode=pool.ceil_mode)
                shortcut, _ = BF.bconv2d_fused(pooled, wts, bn=bn_d.get(), channels_last=True, **kw)
        else:
            shortcut = x
        dev = x.device
        bits = xbits
        last = len(plan.stages) - 1
        for i, (conv, bn, act_mod) in enumerate(plan.stages):
            code, prelu = _activation_spec(act_mod)
            kw, wts = _conv_args(conv)
            slope = _slope(prelu, conv.out_channels, dev)
            if i == last and join is not None:
                join[0].wait_stream(join[1])            # join: the launch below adds the shortcut
            if plan.kind == "basic":
                # conv -> bn -> act -> [sign of the next conv]; the last conv adds the shortcut before its activation
                if i < last:
                    _, bits = BF.bconv2d_fused(bits, wts, bn=bn.get(), activation=code, act_slope=slope,
                                               want_out=False, want_bits=True, **kw)
                else:
                    y, bits = BF.bconv2d_fused(bits, wts, bn=bn.get(), residual=shortcut, activation=code,
                                               act_slope=slope, want_out=True, want_bits=want_next_bits, nx=nx,
                                               channels_last=True, **kw)
            else:
                # [bn -> sign] -> conv -> act; the planes handed on go through the NEXT conv's BatchNorm;
                # the shortcut is added after the last activation
                if i < last:
                    _, bits = BF.bconv2d_fused(bits, wts, activation=code, act_slope=slope, want_out=False,
                                               want_bits=True, nx=plan.stages[i + 1][1].get(), **kw)
                else:
                    y, bits = BF.bconv2d_fused(bits, wts, residual=shortcut, residual_after_act=True, activation=code,
                                               act_slope=slope, want_out=True, want_bits=want_next_bits, nx=nx,
                                               channels_last=True, **kw)
        return y, bits

    def set_uint8_input(self, mean, std) -> "FusedResNet":
        """Accept decoded images as uint8 [n,h,w,3] tensors: the tcgen05 stem normalises them while it stages the input
        window, ``(x.float() - mean) * (1 / std)`` per channel (mean / std on the 0..255 scale, e.g. ImageNet's
        (123.675, 116.28, 103.53) / (58.395, 57.12, 57.375)) -- a quarter of the host-to-device bytes of an fp32 batch."""
        mean = [float(v) for v in mean]
        istd = [float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(v), dtype=torch.float32)) for v in std]
        if len(mean) != 3 or len(istd) != 3:
            raise ValueError("mean and std must have three entries")
        self.u8_norm = (mean, istd)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = self.model
        if m.training or (torch.is_grad_enabled() and any(p.requires_grad for p in m.parameters()) and x.requires_grad):
            raise native.NativeError("FusedResNet is an inference engine: call model.eval() and use torch.no_grad()")
        with torch.no_grad():
            bits = None
            first = self.plans[0] if self.plans else None
            shape_key = (tuple(x.shape), x.dtype)
            self._overlap_now = shape_key in self._shapes_seen
            self._shapes_seen.add(shape_key)
            if x.dtype == torch.uint8:
                if getattr(self, "u8_norm", None) is None:
                    raise native.NativeError("uint8 input: call engine.set_uint8_input(mean, std) first")
                if not (self.stem is not None and self.stem.ok and first is not None and first.fused and x.is_cuda and x.dim() == 4
                        and x.shape[3] == 3 and self.stem_kernel == "tc"):
                    raise native.NativeError("uint8 [n,h,w,3] input needs the tcgen05 stem kernel and a fused first block")
                self.stem_kernel_used = "bnn_stem_tc_fwd(uint8)"
                x, bits = BF.stem_tc(x.contiguous(), self.stem.tc_weight(), self.stem.bn.get(), nx=self._entry_affine(first),
                                     u8_norm=self.u8_norm)
            elif (self.stem is not None and self.stem.ok and first is not None and first.fused and x.is_cuda
                    and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3 and min(x.shape[2:]) >= 7):
                # fp32 stem in one kernel: NHWC residual stream + the first binarized conv's planes
                self.stem_kernel_used = {"mma": "bnn_stem_mma_fwd", "fma": "bnn_stem_fwd", "tc": "bnn_stem_tc_fwd"}[self.stem_kernel]
                if self.stem_kernel == "tc":
                    x, bits = BF.stem_tc(x.contiguous(), self.stem.tc_weight(), self.stem.bn.get(),
                                         nx=self._entry_affine(first), guard=self.input_range is None,
                                         x_log2_scale=self._x_log2_scale)
                elif self.stem_kernel == "mma":
                    x, bits = BF.stem_mma(x.contiguous(), self.stem.mma_weight(), self.stem.bn.get(),
                                          nx=self._entry_affine(first), guard=self.input_range is None,
                                          x_log2_scale=self._x_log2_scale)
                else:
                    x, bits = BF.stem(x.contiguous(), self.stem.weight(), self.stem.bn.get(),
                                      nx=self._entry_affine(first), flags=runtime.kernel_flags())
            else:
                x = m.conv1(x)
                if getattr(m, "stem_type", "basic") == "basic" and hasattr(m, "bn1"):
                    x = m.maxpool(m.relu(m.bn1(x)))
                # the residual stream lives in NHWC between fused blocks (lanes <-> channels: every load/store of
                # the conv epilogue is a full 128-byte line)
                if first is not None and first.fused:
                    x = x.contiguous(memory_format=torch.channels_last)
            for i, plan in enumerate(self.plans):
                nxt = self.plans[i + 1] if i + 1 < len(self.plans) else None
                x, bits = self._run_block(plan, x, bits, nxt)
            x = torch.flatten(m.avgpool(x), 1)
            return m.fc(x)


class _HBlockPlan:
    """Hierarchical block (reference bnn/models/layers/hierarchical_block.py:8-60): three bn-ReLU-conv stages whose
    outputs are concatenated, plus the shortcut.  Every conv input is relu(bn(.)), i.e. its planes are ``bn(.) > 0``."""

    def __init__(self, block: nn.Module) -> None:
        self.block, self.ok = block, False
        names = ("bn1", "conv1", "act1", "bn2", "conv2", "act2", "bn3", "conv3", "act3")
        if type(block).__name__ != "HBlock" or not all(hasattr(block, k) for k in names):
            return
        convs = [block.conv1, block.conv2, block.conv3]
        bns = [block.bn1, block.bn2, block.bn3]
        ok = all(_fusable_conv(c) for c in convs) and all(isinstance(b, nn.BatchNorm2d) and b.track_running_stats for b in bns)
        ok = ok and all(isinstance(a, nn.ReLU) for a in (block.act1, block.act2, block.act3))
        ok = ok and all(b.num_features == c.in_channels for b, c in zip(bns, convs))
        ds = getattr(block, "downsample", None)
        self.shortcut = None
        if ok and ds is not None:
            # the harness shortcut: BatchNorm -> ReLU -> binarized conv1x1
            good = (isinstance(ds, nn.Sequential) and len(ds) == 3 and isinstance(ds[0], nn.BatchNorm2d)
                    and isinstance(ds[1], nn.ReLU) and _fusable_conv(ds[2]) and ds[0].track_running_stats)
            ok = good
            if good:
                self.shortcut = (_FoldedBN(ds[0]), ds[2])
        if ok:
            self.ok = True
            self.convs, self.bns = convs, [_FoldedBN(b) for b in bns]


def run_hblock(plan: _HBlockPlan, x: torch.Tensor, bits=None, shortcut_bits=None) -> torch.Tensor:
    """One fused HBlock on a channels_last fp32 tensor: 3 (+2 with a shortcut conv) launches after the input pack.
    ``bits`` / ``shortcut_bits``: the planes of relu(bn1(x)) / relu(bn_shortcut(x)) when the producer already emitted them."""
    if bits is None:
        bits = BF.pack_activations(x, pre=plan.bns[0].get(), pre_relu=True)
    if plan.shortcut is not None:
        bn_d, conv_d = plan.shortcut
        kw, wts = _conv_args(conv_d)
        if shortcut_bits is None:
            shortcut_bits = BF.pack_activations(x, pre=bn_d.get(), pre_relu=True)
        res, _ = BF.bconv2d_fused(shortcut_bits, wts, channels_last=True, **kw)
    else:
        res = x
    n, _, h, w = x.shape
    planes = sum(c.out_channels for c in plan.convs)
    y = torch.empty((n, planes, h, w), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    c0 = 0
    for i, conv in enumerate(plan.convs):
        kw, wts = _conv_args(conv)
        sl = slice(c0, c0 + conv.out_channels)
        nxt = plan.bns[i + 1].get() if i + 1 < len(plan.convs) else None
        # block output slice = conv + shortcut slice; the next stage sees relu(bn(conv)) -> planes before the add
        _, bits = BF.bconv2d_fused(bits, wts, residual=res[:, sl], residual_after_act=True, out=y[:, sl],
                                   want_bits=nxt is not None, nx=nxt, nx_relu=True, bits_before_residual=True, **kw)
        c0 += conv.out_channels
    return y


class FusedHBlockNet(nn.Module):
    """Engine for the Hierarchical-Block harness of BASELINE configs[3] (``workloads.HBlockNet`` layout:
    conv1/bn1/relu, block0, pool, blocks, avgpool, fc): fp32 stem and pooling stay torch ops, every HBlock is fused."""

    def __init__(self, model: nn.Module, fuse_stem: bool = True, input_range: float = None) -> None:
        super().__init__()
        self.model = model
        self.plans = [_HBlockPlan(model.block0)] + [_HBlockPlan(b) for b in model.blocks]
        self.stem = _StemPlan(model, need_pool=False) if fuse_stem else None
        self.stem_kernel_used = "torch"
        self.input_range = input_range
        self._x_log2_scale = (BF.STEM_X_LOG2_SCALE if input_range is None
                              else max(-60, min(60, 15 - math.frexp(float(input_range))[1])))

    @property
    def fused_blocks(self) -> int:
        return sum(p.ok for p in self.plans)

    def invalidate_caches(self) -> None:
        for p in self.plans:
            if p.ok:
                for bn in p.bns:
                    bn.key = None
                if p.shortcut is not None:
                    p.shortcut[0].key = None
        if self.stem is not None and self.stem.ok:
            self.stem.bn.key = self.stem.key = self.stem.mma_key = self.stem.tc_key = None

    def set_uint8_input(self, mean, std) -> "FusedHBlockNet":
        """Accept decoded images as uint8 [n,h,w,3] tensors (see ``FusedResNet.set_uint8_input``)."""
        mean = [float(v) for v in mean]
        istd = [float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(v), dtype=torch.float32)) for v in std]
        if len(mean) != 3 or len(istd) != 3:
            raise ValueError("mean and std must have three entries")
        self.u8_norm = (mean, istd)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m = self.model
        if m.training:
            raise native.NativeError("FusedHBlockNet is an inference engine: call model.eval()")
        with torch.no_grad():
            p0 = self.plans[0]
            if x.dtype == torch.uint8:
                if getattr(self, "u8_norm", None) is None:
                    raise native.NativeError("uint8 input: call engine.set_uint8_input(mean, std) first")
                if not (self.stem is not None and self.stem.ok and p0.ok and x.is_cuda and x.dim() == 4 and x.shape[3] == 3):
                    raise native.NativeError("uint8 [n,h,w,3] input needs the tcgen05 stem kernel and a fused first block")
                self.stem_kernel_used = "bnn_stem_tc_fwd(no pool, uint8)"
                nx2 = p0.shortcut[0].get() if p0.shortcut is not None else None
                res = BF.stem_tc(x.contiguous(), self.stem.tc_weight(), self.stem.bn.get(), nx=p0.bns[0].get(), nx_relu=True,
                                 nx2=nx2, nx2_relu=True, pool=False, u8_norm=self.u8_norm)
                x = run_hblock(p0, res[0], res[1], res[2] if nx2 is not None else None)
            elif (self.stem is not None and self.stem.ok and p0.ok and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
                    and x.shape[1] == 3 and min(x.shape[2:]) >= 7):
                # conv7x7/2 + BN + ReLU on the tcgen05 stem kernel (no max-pool): NHWC fp32 + the planes of block0's
                # conv1 (bn1-ReLU-sign) and of its shortcut conv (its own BatchNorm) in one launch
                self.stem_kernel_used = "bnn_stem_tc_fwd(no pool)"
                nx2 = p0.shortcut[0].get() if p0.shortcut is not None else None
                res = BF.stem_tc(x.contiguous(), self.stem.tc_weight(), self.stem.bn.get(), nx=p0.bns[0].get(), nx_relu=True,
                                 nx2=nx2, nx2_relu=True, pool=False, guard=self.input_range is None,
                                 x_log2_scale=self._x_log2_scale)
                x = run_hblock(p0, res[0], res[1], res[2] if nx2 is not None else None)
            else:
                x = m.relu(m.bn1(m.conv1(x))).contiguous(memory_format=torch.channels_last)
                x = run_hblock(p0, x) if p0.ok else m.block0(x)
            bits = None
            nxt = self.plans[1] if len(self.plans) > 1 else None
            pool = m.pool
            fusable_pool = (isinstance(pool, nn.AvgPool2d) and _pair(pool.kernel_size) == (2, 2) and _pair(pool.stride) == (2, 2)
                            and _pair(pool.padding) == (0, 0) and not pool.ceil_mode and pool.divisor_override is None
                            and nxt is not None and nxt.ok and x.shape[1] in (64, 128, 256, 512)
                            and x.is_contiguous(memory_format=torch.channels_last))
            if fusable_pool:
                # pool + the next block's bn1-ReLU-sign in one pass over the 1 GB tensor
                x, bits = BF.avgpool2_pack(x, pre=nxt.bns[0].get(), pre_relu=True)
            else:
                x = pool(x)
            for i, plan in enumerate(self.plans[1:]):
                x = run_hblock(plan, x, bits if i == 0 else None) if plan.ok else plan.block(x)
            return m.fc(torch.flatten(m.avgpool(x), 1))


def optimize(model: nn.Module, fuse_stem: bool = True, stem: str = "auto", input_range: float = None,
             overlap_shortcuts: bool = True) -> nn.Module:
    """Return the fused inference engine for ``model`` if its layout is recognised, else ``model``.
    ``stem``: "tc" (= "auto") the stem kernel on the tcgen05 tensor cores, "mma" the same arithmetic on mma.sync (both:
    split-fp16 operands, fp32-level accuracy), "fma" the fp32 fma-chain stem kernel (bit-identical to the oracle's
    summation order).  ``input_range``: a bound on |x| the caller guarantees (e.g. 3.0 for normalised images) -- the
    split-fp16 stems then use a fixed input scale; by default (None) they measure max|x| on the device every forward
    (one extra pass over the input, graph-capturable), so inputs of any magnitude are handled.  ``overlap_shortcuts``:
    launch each down-sampling shortcut on a second stream, concurrently with the first convs of its block."""
    needed = ("conv1", "layer1", "layer2", "layer3", "layer4", "avgpool", "fc")
    if all(hasattr(model, k) for k in needed):
        engine = FusedResNet(model, fuse_stem=fuse_stem, stem=stem, input_range=input_range,
                             overlap_shortcuts=overlap_shortcuts)
        if engine.fused_blocks:
            return engine
    if all(hasattr(model, k) for k in ("conv1", "bn1", "relu", "block0", "pool", "blocks", "avgpool", "fc")):
        engine = FusedHBlockNet(model, fuse_stem=fuse_stem, input_range=input_range)
        if engine.fused_blocks:
            return engine
    return model
