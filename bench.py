#!/usr/bin/env python
"""bench.py -- images/sec of a binarized-network forward (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine, BASELINE configs[1]
    python bench.py --config resnet50|hblock ...                    # BASELINE configs[2] / configs[3], same line
    python bench.py --impl reference ...                            # the reference's CPU float-sim arm
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward pass of the prepared model over one synthetic batch per GPU (resnet18: 256 images of
224x224, BASELINE configs[1]; resnet50: 128 of 224x224; hblock: 64 of 256x256), random-init weights, randomised
BatchNorm statistics.  One JSON line is printed by rank 0.

  value      whole-job images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public module API with PINNED HOST input every step
             (H2D of the batch and D2H of the logits inside the timed region); h2d_ceiling = what the box's
             host->device path delivers for the same buffers with all ranks copying at once
  roofline   the binarized layers' launches as binary MAC/s against the POPC-pipe peak measured on this GPU by
             bnn_ubench (these layers are POPC-bound); roofline.hbm is the same launches as
             algorithmic bytes/s against MEASURED_PEAKS.json hbm_gbs
  tighter_roofline  sum over layers of max(bytes / HBM rate, bMAC / POPC rate) against the launches and the step
  parity     rows of the TIMED graph's output (every rank's block at N > 1) against the reference's CPU forward
  dropin     the same model without fuse.optimize (prepare_binary_model only: per-layer kernels + torch glue)
  cpu_baseline  the reference's CPU float simulation on this box's host cores, bounded sample
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "resnet18": dict(metric="images/sec ResNet-18 XNOR fwd @bs256", batch=256, res=224,
                     what="resnet18 XNOR-Net ({variant}, 19 binarized convs, first/last fp32)"),
    # configs[2]: ResNet-50, XNOR-Net++ (learned per-channel post scale), fc patched to 2048 inputs
    "resnet50": dict(metric="images/sec ResNet-50 XNOR++ fwd @bs128", batch=128, res=224,
                     what="resnet50 XNOR-Net++ (Bottleneck, 52 binarized convs, learned alpha, first/last fp32)"),
    # configs[3]: Hierarchical-Block harness
    "hblock": dict(metric="images/sec HBlock-net XNOR fwd @bs64", batch=64, res=256,
                   what="HBlock harness (stem, HBlock(64,256)+BN-ReLU-1x1 shortcut, avgpool2, 4xHBlock(256,256); 16 binarized convs)"),
}


def workload_config(args, world):
    """`config` of the JSON line: identical for the B200 arm and the reference arm of the same invocation."""
    c = CONFIGS[args.config]
    return {"workload": f"{c['what'].format(variant=args.variant)} {c['res']}x{c['res']} bs{args.batch}/GPU",
            "global_batch": args.batch * world, "parallelism": f"dp{world}"}


def build_model(config: str, variant: str):
    """The workload prepared with THIS repo's package (bnn_b200)."""
    import bnn_b200 as bnn
    from bnn_b200 import workloads
    from bnn_b200.ops import BasicInputBinarizer, BasicScaleBinarizer, XNORWeightBinarizer
    torch.manual_seed(0)
    post = bnn.Identity
    if config == "resnet18":
        model = workloads.resnet18(workloads.PreBasicBlock, nn.PReLU) if variant == "pre_prelu" else workloads.resnet18()
    elif config == "resnet50":
        model, post = workloads.resnet50(), BasicScaleBinarizer
    else:
        model = workloads.HBlockNet()
    cfg = bnn.BConfig(BasicInputBinarizer, post, XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    model = bnn.prepare_binary_model(model, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(model, seed=1)
    return model.eval()


def build_reference_model(config: str, variant: str, state_dict=None):
    """The same workload built and prepared by the UNMODIFIED reference (oracle/_ref, byte-compiled from
    /root/reference by oracle/build.py).  Returns None when oracle/_ref is absent.  With `state_dict`, the
    parameters of a bnn_b200-prepared model are loaded (the keys interchange by design)."""
    from oracle import build as oracle_build
    ref = oracle_build.load_ref()
    if ref is None:
        return None
    from bnn_b200 import workloads                       # randomize_batchnorm + the harness skeleton only
    rops = importlib.import_module("bnn_ref.ops")
    rresnet = importlib.import_module("bnn_ref.models.resnet")
    rlayers = importlib.import_module("bnn_ref.models.layers")
    torch.manual_seed(0)
    post = ref.Identity
    if config == "resnet18":
        model = (rresnet.resnet18(block_type=rlayers.PreBasicBlock, activation=nn.PReLU) if variant == "pre_prelu"
                 else rresnet.resnet18())
    elif config == "resnet50":
        model = rresnet.resnet50()
        model.fc = nn.Linear(2048, 1000)                 # upstream wires fc to 512 features (resnet.py:101,143)
        post = rops.BasicScaleBinarizer
    else:
        model = workloads.HBlockNet(hblock=lambda i, p, d: rlayers.HBlock(i, p, downsample=d, norm_layer=nn.BatchNorm2d))
    cfg = ref.BConfig(activation_pre_process=rops.BasicInputBinarizer, activation_post_process=post,
                      weight_pre_process=rops.XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    model = ref.prepare_binary_model(model, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(model, seed=1)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model.eval()


def layer_algorithmics(model, batch, res):
    """Algorithmic bytes / binary MACs per binarized layer at the drop-in contract."""
    import copy
    import bnn_b200 as bnn
    shapes = {}
    twin = copy.deepcopy(model).cpu()
    hooks = []
    for name, m in twin.named_modules():
        if isinstance(m, bnn.layers.Conv2d):
            hooks.append(m.register_forward_hook(
                lambda mod, inp, out, name=name: shapes.__setitem__(name, (tuple(inp[0].shape), tuple(out.shape)))))
    with torch.no_grad(), bnn.runtime.floatsim_enabled():
        twin(torch.zeros(1, 3, res, res))
    table = {}
    for name, m in twin.named_modules():
        if name in shapes:
            (_, ci, h, w), (_, co, ho, wo) = shapes[name]
            k = ci * m.kernel_size[0] * m.kernel_size[1]
            table[name] = dict(bytes=batch * 4 * (ci * h * w + co * ho * wo) + co * k // 8 + 8 * co,
                               bmac=batch * co * ho * wo * k)
    return table


def fused_layer_order(model):
    """Names of the binarized convs in the order the fused engines launch them (shortcut first)."""
    names = []
    if hasattr(model, "layer1"):
        for lname in ("layer1", "layer2", "layer3", "layer4"):
            for bi, blk in enumerate(getattr(model, lname)):
                if getattr(blk, "downsample", None) is not None:
                    names.append(f"{lname}.{bi}.downsample.1")
                names += [f"{lname}.{bi}.{c}" for c in ("conv1", "conv2", "conv3") if hasattr(blk, c)]
    elif hasattr(model, "block0"):
        names += ["block0.downsample.2", "block0.conv1", "block0.conv2", "block0.conv3"]
        for bi in range(len(model.blocks)):
            names += [f"blocks.{bi}.{c}" for c in ("conv1", "conv2", "conv3")]
    return names


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """keep only the samples taken while the timed region ran (50 ms slack either side)"""
        self.t0, self.t1 = t0 - 0.05, t1 + 0.05

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", -1e30), getattr(self, "t1", 1e30)
        for stamp, line in self.lines:
            if not (t0 <= stamp <= t1):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def measured_traffic(config):
    """DRAM bytes of the binarized-path launches of one step, from the committed ncu capture (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if config == "resnet18" and os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def bind_host_to_gpu_numa(dev_index: int):
    """Multi-GPU runs: run this rank's host threads (and therefore first-touch its pinned upload buffers) on the NUMA
    node the GPU hangs off, so eight ranks do not pull their batches across the socket interconnect.  Best
    effort: any missing sysfs entry or permission leaves the affinity as it was.  Returns the node or None."""
    try:
        props = torch.cuda.get_device_properties(dev_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                                   # noqa: BLE001 -- strictly optional
        return None


def pick_cpu_threads(twin, res):
    """The float simulation is many small torch ops; on a many-core host the default (all cores) is not always
    the fastest setting.  Give the CPU arm its best case: try a few intra-op thread counts on a tiny batch."""
    ncpu = os.cpu_count() or 1
    cands = sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 16)}, reverse=True)
    x = torch.randn(8, 3, res, res, generator=torch.Generator().manual_seed(0))
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for t in cands:
            torch.set_num_threads(t)
            twin(x)
            t0 = time.perf_counter()
            twin(x)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = t, dt
    torch.set_num_threads(best)
    return best


def cpu_twin(config, variant, model_cpu=None):
    """The CPU arm's model: the unmodified reference (oracle/_ref) when it is there, else the oracle's port of it.
    Returns (module, kind, description)."""
    sd = None if model_cpu is None else model_cpu.state_dict()
    ref_model = build_reference_model(config, variant, sd)
    if ref_model is not None:
        return ref_model, "reference", "unmodified reference bnn 0.1.2 (oracle/_ref/bnn_ref, byte-compiled from /root/reference)"
    from oracle import floatsim
    return floatsim.mirror_model(model_cpu if model_cpu is not None else build_model(config, variant)), "port", \
        "oracle/floatsim.py (torch CPU restatement of the reference's forward)"


def cpu_floatsim_rate(config, variant, model_cpu, sample_batch, iters, res):
    """images/s of the reference's CPU float simulation on host cores (bounded sample)."""
    twin, kind, what = cpu_twin(config, variant, model_cpu)
    threads = pick_cpu_threads(twin, res)
    x = torch.randn(sample_batch, 3, res, res, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        twin(x)                                           # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            twin(x)
        dt = time.perf_counter() - t0
    return sample_batch * iters / dt, threads, kind, what


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU float-sim forward on all host cores; every step is one forward over a
    bounded sample (--ref-batch images, default = the full per-GPU batch of the configuration)."""
    if rank != 0:
        return
    c = CONFIGS[args.config]
    twin, kind, what = cpu_twin(args.config, args.variant)
    threads = pick_cpu_threads(twin, c["res"])
    sample = args.ref_batch or args.batch
    x = torch.randn(sample, 3, c["res"], c["res"], generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        t0 = time.perf_counter()
        twin(x)
        t1 = time.perf_counter() - t0
        # every step is one forward over the full per-GPU batch; only if --steps is so large that the run would not end
        # within a few minutes is the per-step sample cut (the rate in images/s is what is reported either way)
        budget = 240.0
        if not args.ref_batch and t1 * (args.steps + args.warmup) > budget:
            sample = max(16, int(sample * budget / (t1 * (args.steps + args.warmup))) // 8 * 8)
            x = x[:sample].contiguous()
        for _ in range(max(0, args.warmup - 1)):
            twin(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            twin(x)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": c["metric"], "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} forwards of {sample} images, {what}, torch CPU fp32, "
                                   f"{threads} of {os.cpu_count()} host threads (fastest of a small sweep)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def percentile(v, q):
    s = sorted(v)
    if not s:
        return None
    i = (len(s) - 1) * q
    lo, hi = int(i), min(int(i) + 1, len(s) - 1)
    return s[lo] + (s[hi] - s[lo]) * (i - lo)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="resnet18", choices=sorted(CONFIGS))
    ap.add_argument("--variant", default="basic_relu", choices=["basic_relu", "pre_prelu"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the configuration's)")
    ap.add_argument("--ref-batch", type=int, default=0, help="images per CPU step of the reference arm (default: --batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph")
    ap.add_argument("--no-fuse", action="store_true", help="per-layer kernels + torch glue (no cross-module fusion)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the extra timing of the unfused drop-in path")
    ap.add_argument("--stem", default="auto", choices=["auto", "tc", "mma", "fma"],
                    help="fused engine's stem kernel: tcgen05 (tc), mma.sync split-fp16 (mma) or the fp32 fma chain")
    ap.add_argument("--layers-out", default=None, help="write the per-layer table to this JSON file")
    ap.add_argument("--shortcut-max-cin", type=int, default=-1, help="experiment knob: bnn_b200.runtime.shortcut_max_cin")
    ap.add_argument("--no-overlap", action="store_true", help="A/B: shortcuts on the main stream instead of a second one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfgd = CONFIGS[args.config]
    args.batch = args.batch or cfgd["batch"]
    RES = cfgd["res"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 engine has no CPU fallback (use --impl reference)")
    if args.shortcut_max_cin >= 0:
        from bnn_b200 import runtime as _rt
        _rt.shortcut_max_cin(args.shortcut_max_cin)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_host_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    # fp32 glue (stem conv, fc) in true fp32: the reference's CPU float-sim is the parity target
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    import bnn_b200 as bnn
    from bnn_b200 import functional as BF
    from bnn_b200 import native, sharded

    model_cpu = build_model(args.config, args.variant)
    algo = layer_algorithmics(model_cpu, args.batch, RES)
    import copy
    model = copy.deepcopy(model_cpu).to(dev)
    from bnn_b200 import fuse
    engine = model if args.no_fuse else fuse.optimize(model, stem=args.stem, overlap_shortcuts=not args.no_overlap)   # public API: bnn_b200.fuse.optimize
    B = args.batch
    # every rank gets its own images (seed 1000 + rank): rank 0 can regenerate any rank's rows for the parity check
    x_host = torch.randn(B, 3, RES, RES, generator=torch.Generator().manual_seed(1000 + rank)).pin_memory()
    x_dev = x_host.to(dev)                                 # 154 MB at bs256 > 126 MB L2

    def step_resident():
        y = engine(x_dev)
        if world > 1:
            y = sharded.gather_logits(y)
        return y

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize; returns (total ms, max over ranks; per-step ms list)."""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        sync_all()
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        sync_all()
        ms = torch.tensor([evs[0].elapsed_time(evs[-1])], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]

    def capture(eng):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eng(x_dev)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):          # the forward only; the collective stays outside
                out = eng(x_dev)
        torch.cuda.current_stream().wait_stream(side)
        return g, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # started before warm-up: nvidia-smi needs a moment to deliver samples
    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        sync_all()

        graph = None
        if not args.no_graph:
            # the whole forward as one CUDA graph: the launches of a step stop costing host time
            graph, graph_out = capture(engine)

            def step_graph():
                graph.replay()
                return sharded.gather_logits(graph_out) if world > 1 else graph_out

            for _ in range(2):
                step_graph()
            sync_all()

        launches0 = native.launch_count()
        t_region0 = time.perf_counter()
        launches_per_step = None
        if graph is not None:
            ms, per_step = timed(step_graph, args.steps)
            timed_out = step_graph().clone()
            # a replay re-issues every captured launch; count them from one eager step
            l0 = native.launch_count(); step_resident(); launches_per_step = native.launch_count() - l0
        else:
            ms, per_step = timed(step_resident, args.steps)
            timed_out = step_resident().clone()
            launches_per_step = (native.launch_count() - launches0) // (args.steps + 1)
        if rank == 0:
            sampler.window(t_region0, time.perf_counter())
        clocks = sampler.stop() if rank == 0 else None

        # ---- parity of what was just timed (outside the timed region): rows of every rank's block of the output of the
        # timed graph against the reference's CPU forward on the same rows
        parity = None
        if rank == 0:
            rows_per_rank = max(2, -(-8 // world))
            twin, twin_kind, _ = cpu_twin(args.config, args.variant, model_cpu)
            torch.set_num_threads(os.cpu_count() or 1)

            def cpu_reference_rows(rows):                      # the CHECKER: reference CPU forward of a few rows
                return twin(rows)

            got, want = [], []
            for r in range(world):
                xr = x_host if r == 0 else torch.randn(B, 3, RES, RES, generator=torch.Generator().manual_seed(1000 + r))
                want.append(twin(xr[:rows_per_rank].clone()))
                got.append(timed_out[r * B: r * B + rows_per_rank].float().cpu())
            got, want = torch.cat(got), torch.cat(want)
            per_row = ((got - want).abs().amax(1) / want.abs().max())
            err = float(per_row.max())
            within = int((per_row <= 1e-3).sum())
            parity = {"max_rel_err": err, "median_rel_err": float(per_row.median()), "rows": int(got.shape[0]),
                      "rows_within_tolerance": within, "ranks": world,
                      "argmax_equal": bool((got.argmax(1) == want.argmax(1)).all()), "tolerance": 1e-3,
                      "against": f"{twin_kind} CPU fp32 forward of the same parameters, rows 0..{rows_per_rank - 1} of "
                                 "every rank's block of the timed (graph-replayed, all-gathered) logits"}
            # BASELINE configs[1] (the metric's configuration): every checked row within 1e-3.  The Bottleneck / HBlock
            # networks put a sign() behind continuous thresholds (bn -> relu -> sign; 1x1 convs on pooled maps): an
            # activation within fp32 rounding noise of its threshold flips and cascades, in the reference itself as much
            # as here -- its own logits move by up to 7e-2 when the input is perturbed by one ulp (DESIGN.md section 6,
            # profiles/r02_chaos_reference.txt), and the per-layer path with torch's cuDNN stem shows the same rows off.
            # There the check is: median row within 1e-3, at least three quarters of the rows within 1e-3, every arg-max equal.
            strict = args.config == "resnet18"
            ok = err <= 1e-3 if strict else (parity["median_rel_err"] <= 1e-3 and 4 * within >= 3 * parity["rows"]
                                             and parity["argmax_equal"])
            parity["criterion"] = "all rows" if strict else "median row + 3/4 of rows + arg-max (sign-flip cascades, see DESIGN.md)"
            if not ok:
                raise SystemExit(f"bench.py: parity FAILED on the timed output: {json.dumps(parity)}")

        # ---- host -> device ceiling of this box for the same pinned buffers, every rank copying at once
        h2d_buf = torch.empty_like(x_dev)
        cs = torch.cuda.Stream()
        with torch.cuda.stream(cs):
            for _ in range(2):
                h2d_buf.copy_(x_host, non_blocking=True)
        sync_all()
        ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ncopies = 10
        with torch.cuda.stream(cs):
            ce0.record(cs)
            for _ in range(ncopies):
                h2d_buf.copy_(x_host, non_blocking=True)
            ce1.record(cs)
        sync_all()
        h2d_ms = torch.tensor([ce0.elapsed_time(ce1)], device=dev)
        if world > 1:
            dist.all_reduce(h2d_ms, op=dist.ReduceOp.MAX)
        h2d_bytes = x_host.numel() * 4
        h2d_ceiling_gbs = world * ncopies * h2d_bytes / (float(h2d_ms.item()) * 1e-3) * 1e-9
        del h2d_buf

        # ---- end to end through the public API (bnn_b200.pipeline.HostPipeline): every step uploads the batch from
        # pinned host memory and downloads the logits; upload of batch i+1 overlaps the forward of batch i
        from bnn_b200.pipeline import HostPipeline
        pipe = HostPipeline(engine, x_host, dev, use_graphs=(graph is not None),
                            post=(sharded.gather_logits if world > 1 else None))
        for _ in range(2):
            pipe.submit(x_host)
        pipe.drain()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.compute_stream)
        for _ in range(args.steps):
            pipe.submit(x_host)
        e1.record(pipe.d2h_stream)            # behind the last logits download (which waits for the last forward)
        last_logits = pipe.drain()
        sync_all()
        ms_e2e_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
        ms_e2e = float(ms_e2e_t.item())
        e2e_diff = float((last_logits[:B].float() - timed_out[:B].float().cpu()).abs().max())
        assert torch.isfinite(last_logits).all()
        del pipe

        # ---- the same end-to-end loop fed with DECODED IMAGES (uint8 NHWC, a quarter of the bytes): the tcgen05 stem
        # normalises while it stages the window.  Reported beside `e2e` (fp32 tensors, the reference's input contract).
        e2e_u8 = None
        if hasattr(engine, "set_uint8_input") and not args.no_fuse and args.stem in ("auto", "tc"):
            mean, std = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
            eng8 = fuse.optimize(model, stem="tc", overlap_shortcuts=not args.no_overlap).set_uint8_input(mean, std)
            xu_host = torch.randint(0, 256, (B, RES, RES, 3), dtype=torch.uint8,
                                    generator=torch.Generator().manual_seed(2000 + rank)).pin_memory()
            pipe8 = HostPipeline(eng8, xu_host, dev, use_graphs=(graph is not None),
                                 post=(sharded.gather_logits if world > 1 else None))
            for _ in range(2):
                pipe8.submit(xu_host)
            pipe8.drain()
            sync_all()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record(pipe8.compute_stream)
            for _ in range(args.steps):
                pipe8.submit(xu_host)
            u1.record(pipe8.d2h_stream)
            logits8 = pipe8.drain()
            sync_all()
            ms8_t = torch.tensor([u0.elapsed_time(u1)], device=dev)
            if world > 1:
                dist.all_reduce(ms8_t, op=dist.ReduceOp.MAX)
            ms8 = float(ms8_t.item())
            e2e_u8 = {"value": B * world / (ms8 / args.steps * 1e-3), "unit": "images/s", "ms_per_step": ms8 / args.steps,
                      "h2d_bytes_per_step": xu_host.numel() * world, "d2h_bytes_per_step": B * world * 1000 * 4 * world,
                      "input": "uint8 [n,h,w,3] pinned host batch; (x - mean) * (1/std) folded into bnn_stem_tc_run"}
            if rank == 0:
                k = 4
                istd = torch.tensor([float(torch.tensor(1.0) / torch.tensor(v)) for v in std])
                xn = ((xu_host[:k].float() - torch.tensor(mean)) * istd).permute(0, 3, 1, 2).contiguous()
                want8 = cpu_reference_rows(xn)
                got8 = logits8[:k].float()
                rows8 = (got8 - want8).abs().amax(1) / want8.abs().max()
                e2e_u8["parity_max_rel_err"] = float(rows8.max())
                e2e_u8["parity_rows_within_tolerance"] = int((rows8 <= 1e-3).sum())
                # all rows for ResNet-18; the deeper configs by 3 of 4 rows (sign-flip cascades)
                need = k if args.config == "resnet18" else k - 1
                if e2e_u8["parity_rows_within_tolerance"] < need:
                    raise SystemExit(f"bench.py: uint8 end-to-end parity FAILED: {json.dumps(e2e_u8)}")
            del pipe8, eng8

        # ---- the literal drop-in (prepare_binary_model only, no fuse.optimize): per-layer kernels + torch glue
        dropin = None
        if not args.no_fuse and not args.no_dropin:
            for _ in range(3):
                model(x_dev)
            sync_all()
            dsteps = max(3, min(10, args.steps))
            try:
                g2, _ = capture(model)
                for _ in range(2):
                    g2.replay()
                dms, _ = timed(g2.replay, dsteps)
                del g2
            except RuntimeError:                         # not capturable on this torch build: eager timing
                dms, _ = timed(lambda: model(x_dev), dsteps)
            dropin = {"value": B * world / (dms / dsteps * 1e-3), "unit": "images/s", "ms_per_step": dms / dsteps,
                      "steps": dsteps, "what": "prepare_binary_model only: bit-pack + bnn_bconv2d_fwd per layer, fp32 NCHW "
                                               "between layers, torch stem / BN / act / add (input resident, CUDA graph)"}

        # ---- per-launch CUDA-event timing of the binarized path (same data, same stream, warm)
        per_layer = {}
        if rank == 0:
            records = []
            orig_pack, orig_conv, orig_fused, orig_short = BF.pack_activations, BF.bconv2d, BF.bconv2d_fused, BF.shortcut

            def ev():
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                return e

            def pack_t(x, *a, **k):
                e0 = ev(); r = orig_pack(x, *a, **k); records.append(("pack", e0, ev())); return r

            def conv_t(*a, **k):
                e0 = ev(); r = orig_conv(*a, **k); records.append(("conv", e0, ev())); return r

            def fused_t(act, wts, *a, **k):
                e0 = ev(); r = orig_fused(act, wts, *a, **k)
                records.append(("conv", e0, ev(), (wts.c_in, wts.c_out, wts.kh, act.h, act.w)))
                return r

            def short_t(x, wts, *a, **k):        # pool + sign + conv1x1 + BN of a down-sampling shortcut: one launch
                e0 = ev(); r = orig_short(x, wts, *a, **k); records.append(("conv", e0, ev())); return r

            BF.pack_activations, BF.bconv2d, BF.bconv2d_fused, BF.shortcut = pack_t, conv_t, fused_t, short_t
            order = []
            hooks = [m.register_forward_hook(lambda mod, i, o, n=n: order.append(n))
                     for n, m in model.named_modules() if isinstance(m, bnn.layers.Conv2d)]
            reps = max(3, min(args.steps, 10))
            # launches are timed one by one on one stream here: the engine's second stream (shortcuts concurrent with the
            # first convs of their block) is switched off for this pass -- the step time above includes the overlap
            overlap = getattr(engine, "overlap_shortcuts", None)
            if overlap is not None:
                engine.overlap_shortcuts = False
            for _ in range(reps):
                engine(x_dev)                      # rank-local: no collective in this pass
            torch.cuda.synchronize()
            if overlap is not None:
                engine.overlap_shortcuts = overlap
            BF.pack_activations, BF.bconv2d, BF.bconv2d_fused, BF.shortcut = orig_pack, orig_conv, orig_fused, orig_short
            for h in hooks:
                h.remove()
            packs = [r for r in records if r[0] == "pack"]
            convs = [r for r in records if r[0] == "conv"]
            # median over the repetitions: a single preempted launch must not skew a layer's time
            per_step_c, per_step_p = len(convs) // reps, len(packs) // reps

            def med(recs, per_step, slot):
                if per_step == 0:
                    return 0.0
                return statistics.median(recs[rep * per_step + slot][1].elapsed_time(recs[rep * per_step + slot][2])
                                         for rep in range(reps))

            if order:                                   # unfused: module hooks give the layer names
                for i, name in enumerate(order[:per_step_c]):
                    per_layer[name] = {"pack_ms": med(packs, per_step_p, i) if per_step_p == per_step_c else 0.0,
                                       "conv_ms": med(convs, per_step_c, i)}
            else:                                       # fused engine: conv launches in execution order
                names = fused_layer_order(model)
                matched = per_step_c == len(names)          # one launch per binarized layer (shortcuts included)
                for i in range(per_step_c):
                    per_layer[names[i] if matched else f"launch{i}"] = {"pack_ms": 0.0, "conv_ms": med(convs, per_step_c, i)}
                if not matched:
                    # some block ran unfused: per-layer attribution is lost, the totals are still the layers' totals
                    per_layer["_all_binarized_layers"] = {"pack_ms": 0.0, "conv_ms": 0.0,
                                                          "bytes": sum(v["bytes"] for v in algo.values()),
                                                          "bmac": sum(v["bmac"] for v in algo.values())}
                per_layer["_pack_launches_total"] = {"pack_ms": sum(med(packs, per_step_p, i) for i in range(per_step_p)),
                                                     "conv_ms": 0.0, "bytes": 0, "bmac": 0, "n_per_step": per_step_p}
            for name, d in per_layer.items():
                d.setdefault("bytes", 0); d.setdefault("bmac", 0)
                if name in algo:
                    d.update(algo[name])
                    d["conv_tbmac_s"] = d["bmac"] / (d["conv_ms"] * 1e-3) * 1e-12
                    d["path_gb_s"] = d["bytes"] / ((d["pack_ms"] + d["conv_ms"]) * 1e-3) * 1e-9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    traffic = measured_traffic(args.config)
    n_conv = sum(1 for k, d in per_layer.items() if d.get("conv_ms", 0) > 0)
    total_bytes = sum(d["bytes"] for d in per_layer.values())
    total_bmac = sum(d["bmac"] for d in per_layer.values())
    path_ms = sum(d["pack_ms"] + d["conv_ms"] for d in per_layer.values())
    conv_ms = sum(d["conv_ms"] for d in per_layer.values())
    achieved = total_bytes / (path_ms * 1e-3) * 1e-9
    try:
        popc_gops = BF.ubench(0, 200)                 # POPC warp-lane ops / s on this GPU, now
        mix_gwords = BF.ubench(2, 200)
    except native.NativeError:
        popc_gops = mix_gwords = None
    popc_peak_tbmac = popc_gops * 32 * 1e-3 if popc_gops else None     # one POPC = 32 binary MACs
    images = B * world
    value = images / (ms / args.steps * 1e-3)
    e2e_value = images / (ms_e2e / args.steps * 1e-3)
    achieved_tbmac = total_bmac / (conv_ms * 1e-3) * 1e-12 if conv_ms else None
    stem_name = getattr(engine, "stem_kernel_used", None)
    config = workload_config(args, world)
    line = {
        "metric": cfgd["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config,
        "engine": {"l2": f"input batch {B * 3 * RES * RES * 4 / 1e6:.0f} MB + activations exceed the 126 MB L2",
                   "launch": "cuda_graph" if graph is not None else "eager",
                   "fusion": "per-layer" if args.no_fuse else "bnn_b200.fuse.optimize (BN/act/residual/sign in conv epilogues)",
                   "glue": ("torch fp32 (TF32 off): stem conv+BN+ReLU(+maxpool), BN/act/add per layer, avgpool, fc"
                            if args.no_fuse else f"stem kernel = {stem_name}; torch fp32 only for global avgpool + fc")},
        "ms_per_step_stats": {"median": statistics.median(per_step), "p10": percentile(per_step, 0.1),
                              "p90": percentile(per_step, 0.9), "what": "per-step CUDA-event times of the timed region, rank 0"},
        "clocks": clocks,
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": "images/s",
                "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": images * 1000 * 4 * world,
                "ms_per_step": ms_e2e / args.steps, "max_abs_diff_vs_resident_logits": e2e_diff,
                "h2d_ceiling_gbs": h2d_ceiling_gbs,
                "h2d_frac_of_ceiling": (h2d_bytes * world / (ms_e2e / args.steps * 1e-3) * 1e-9) / h2d_ceiling_gbs,
                "h2d_ceiling_what": f"{ncopies} back-to-back cudaMemcpyAsync of the same pinned {h2d_bytes / 1e6:.0f} MB batch "
                                    f"on {world} rank(s) at once, no compute: aggregate GB/s (max time over ranks)",
                "how": "bnn_b200.pipeline.HostPipeline: pinned-host batch -> H2D -> fused engine -> D2H logits, "
                       "double-buffered (upload of step i+1 overlaps the forward of step i)"
                       + (f"; rank 0 host threads and pinned buffers bound to NUMA node {numa_node}" if numa_node is not None else "")},
        "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else 0,
        "roofline": {"bound": "popc", "achieved": achieved_tbmac, "peak": popc_peak_tbmac, "unit": "T bMAC/s",
                     "frac": (achieved_tbmac / popc_peak_tbmac) if (popc_peak_tbmac and achieved_tbmac) else None,
                     "peak_kind": "bnn_ubench(0): POPC warp-lane ops/s measured on this GPU in this run x 32 binary MACs "
                                  "(SASS loop: profiles/r02_ubench_popc_sass.txt); nominal 148 SM x 16 lanes x 1.965 GHz x 32 = 148.9",
                     "launches": n_conv, "algorithmic_bmac_per_launch": total_bmac / max(1, n_conv),
                     "avg_launch_ms": conv_ms / max(1, n_conv), "lop3_popc_iadd_gwords_s": mix_gwords,
                     "traffic": (traffic["binarized_path_dram_bytes_per_step"] / max(1, n_conv)
                                 if traffic and not args.no_fuse and B == cfgd["batch"] else None),
                     "traffic_what": "ncu dram__bytes_read+write of the same launches, per conv launch "
                                     "(profiles/traffic.json; fused engine at bs 256): below the algorithmic bytes "
                                     "because fused layers exchange bit planes instead of fp32 NCHW tensors",
                     "what": f"XNOR-popcount launches of the {len(algo)} binarized layers: {total_bmac / 1e9:.1f} G binary MACs "
                             f"per step over {conv_ms:.3f} ms; every 3x3 layer is POPC-bound (SURVEY.md 8(d))",
                     "hbm": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / peaks["hbm_gbs"], "peak_kind": peak_kind,
                             "algorithmic_bytes_per_launch": total_bytes / max(1, n_conv),
                             "what": f"same launches (+ bit-pack launches) as algorithmic bytes at the drop-in contract: "
                                     f"{total_bytes / 1e9:.3f} GB/step over {path_ms:.3f} ms/step"}},
    }
    if e2e_u8 is not None:
        line["e2e_u8"] = e2e_u8
    if dropin is not None:
        line["dropin"] = dropin
    if popc_peak_tbmac:
        # north_star's yardstick: per layer t >= max(bytes / BW_HBM, bMAC / P_popc), summed over the
        # binarized layers, against (i) the time of those launches and (ii) the WHOLE step (stem, classifier included)
        bound_ms = sum(max(d["bytes"] / (peaks["hbm_gbs"] * 1e9), d["bmac"] / (popc_peak_tbmac * 1e12)) * 1e3
                       for d in per_layer.values() if d.get("bmac"))
        hbm_ms = total_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
        line["tighter_roofline"] = {
            "bound": "popc", "lower_bound_ms_per_step": bound_ms, "hbm_only_lower_bound_ms_per_step": hbm_ms,
            "binarized_path_frac": bound_ms / path_ms if path_ms else None,
            "whole_step_frac": bound_ms / (ms / args.steps),
            "what": "sum over the binarized layers of max(algorithmic bytes / measured HBM rate, binary MACs / measured "
                    "POPC rate), against the time of those launches and against the whole step (stem, shortcuts, "
                    "classifier included): the figure north_star's '>= 60 % of the tighter roofline' refers to"}
    if world == 1 and not args.no_cpu_baseline:
        sample, iters = min(64, B), 10          # ~6-20 s of host work depending on the box
        rate, threads, kind, what = cpu_floatsim_rate(args.config, args.variant, model_cpu, sample, iters, RES)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": kind,
                                "sample": f"{iters} forwards of {sample} images ({what}, torch CPU fp32, "
                                          f"{threads} of {os.cpu_count()} host threads: the fastest of a small sweep)"}
    if args.layers_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
        with open(args.layers_out, "w") as f:
            json.dump({"per_layer": per_layer, "line": line}, f, indent=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
