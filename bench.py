#!/usr/bin/env python
"""bench.py -- images/sec of the ResNet-18 XNOR-Net forward (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference ...                            # reference CPU float-sim arm
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward pass of the prepared model over one synthetic batch of 256 images per
GPU (BASELINE configs[1]; 224x224, random-init weights, randomised BatchNorm statistics --
SURVEY.md section 8(d)).  One JSON line is printed by rank 0.

  value      whole-job images/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public module API with PINNED HOST input every step
             (H2D of the batch and D2H of the logits inside the timed region)
  roofline   binarized-layer path (bit-pack + XNOR-popcount kernels): algorithmic bytes at the
             drop-in contract (fp32 NCHW in + fp32 NCHW out per layer) divided by
             the CUDA-event time of those launches, against MEASURED_PEAKS.json hbm_gbs
  popc       same launches as binary MAC/s against the POPC-pipe peak measured by bnn_ubench
  cpu_baseline  the oracle's float simulation (oracle/floatsim.py, a torch-CPU restatement of
             the reference's forward) timed on this box's host cores on a bounded sample
"""
import argparse
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

METRIC = "images/sec ResNet-18 XNOR fwd @bs256"
BATCH_PER_GPU = 256
RES = 224


def build_model(variant: str):
    import bnn_b200 as bnn
    from bnn_b200 import workloads
    from bnn_b200.ops import BasicInputBinarizer, XNORWeightBinarizer
    torch.manual_seed(0)
    if variant == "pre_prelu":
        model = workloads.resnet18(workloads.PreBasicBlock, nn.PReLU)
    else:
        model = workloads.resnet18()
    cfg = bnn.BConfig(BasicInputBinarizer, bnn.Identity,
                      XNORWeightBinarizer.with_args(compute_alpha=True, center_weights=True))
    model = bnn.prepare_binary_model(model, cfg, ignore_layers_name=["_first_", "_last_"])
    workloads.randomize_batchnorm(model, seed=1)
    return model.eval()


def layer_algorithmics(model, batch, res):
    """Algorithmic bytes / binary MACs per binarized layer at the drop-in contract."""
    import bnn_b200 as bnn
    shapes = {}
    hooks = []
    for name, m in model.named_modules():
        if isinstance(m, bnn.layers.Conv2d):
            hooks.append(m.register_forward_hook(
                lambda mod, inp, out, name=name: shapes.__setitem__(name, (tuple(inp[0].shape), tuple(out.shape)))))
    with torch.no_grad(), bnn.runtime.floatsim_enabled():
        twin_in = torch.zeros(1, 3, res, res)
        import copy
        copy.deepcopy(model).cpu()(twin_in)
    for h in hooks:
        h.remove()
    table = {}
    for name, m in model.named_modules():
        if name in shapes:
            (_, ci, h, w), (_, co, ho, wo) = shapes[name]
            k = ci * m.kernel_size[0] * m.kernel_size[1]
            table[name] = dict(bytes=batch * 4 * (ci * h * w + co * ho * wo) + co * k // 8 + 8 * co,
                               bmac=batch * co * ho * wo * k)
    return table


def fused_layer_order(model):
    """Names of the binarized convs in the order the fused engine launches them (shortcut first)."""
    names = []
    for lname in ("layer1", "layer2", "layer3", "layer4"):
        for bi, blk in enumerate(getattr(model, lname)):
            if getattr(blk, "downsample", None) is not None:
                names.append(f"{lname}.{bi}.downsample.1")
            names += [f"{lname}.{bi}.conv1", f"{lname}.{bi}.conv2"]
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


def measured_traffic():
    """DRAM bytes of the binarized-path launches of one step, from the committed ncu capture (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def bind_host_to_gpu_numa(dev_index: int):
    """Multi-GPU runs: run this rank's host threads (and therefore first-touch its pinned upload buffers) on the NUMA
    node the GPU hangs off, so eight ranks do not pull their 154 MB batches across the socket interconnect.  Best
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


def pick_cpu_threads(twin):
    """The float simulation is many small torch ops; on a many-core host the default (all cores) is not always
    the fastest setting.  Give the CPU arm its best case: try a few intra-op thread counts on a tiny batch."""
    ncpu = os.cpu_count() or 1
    cands = sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 16)}, reverse=True)
    x = torch.randn(8, 3, RES, RES, generator=torch.Generator().manual_seed(0))
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


def cpu_floatsim_rate(model, sample_batch, iters, threads):
    """images/s of the oracle float simulation on host cores (bounded sample)."""
    from oracle import floatsim
    twin = floatsim.mirror_model(model)
    threads = pick_cpu_threads(twin)
    x = torch.randn(sample_batch, 3, RES, RES, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        twin(x)                                           # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            twin(x)
        dt = time.perf_counter() - t0
    return sample_batch * iters / dt, threads


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU float-sim forward (oracle port: the reference is pure Python
    over torch and is not installed on the GPU box) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    model = build_model(args.variant)
    from oracle import floatsim
    twin = floatsim.mirror_model(model)
    threads = pick_cpu_threads(twin)
    sample = args.ref_batch
    x = torch.randn(sample, 3, RES, RES, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            twin(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            twin(x)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"resnet18 XNOR-Net ({args.variant}, first/last fp32) {RES}x{RES}",
                   "sample": f"{sample} images per step on CPU"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} forwards of {sample} images, oracle/floatsim.py (torch CPU fp32), "
                                   f"{threads} of {os.cpu_count()} host threads (fastest of a small sweep)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="basic_relu", choices=["basic_relu", "pre_prelu"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=64, help="images per CPU step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph")
    ap.add_argument("--no-fuse", action="store_true", help="per-layer kernels + torch glue (no cross-module fusion)")
    ap.add_argument("--stem", default="mma", choices=["mma", "fma"],
                    help="fused engine's stem kernel: mma.sync split-fp16 (default) or the fp32 fma chain")
    ap.add_argument("--layers-out", default=None, help="write the per-layer table to this JSON file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 engine has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_host_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    # fp32 glue (stem conv, fc) in true fp32: the reference's CPU float-sim is the parity target
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    import bnn_b200 as bnn
    from bnn_b200 import functional as BF
    from bnn_b200 import native, sharded

    model_cpu = build_model(args.variant)
    algo = layer_algorithmics(model_cpu, args.batch, RES)
    model = model_cpu.to(dev)
    from bnn_b200 import fuse
    engine = model if args.no_fuse else fuse.optimize(model, stem=args.stem)   # public API: bnn_b200.fuse.optimize
    B = args.batch
    x_dev = torch.randn(B, 3, RES, RES, device=dev)       # 154 MB at bs256 > 126 MB L2
    x_host = torch.randn(B, 3, RES, RES).pin_memory()
    stream = torch.cuda.current_stream()

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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # started before warm-up: nvidia-smi needs a moment to deliver samples
    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        sync_all()

        graph = None
        if not args.no_graph:
            # the whole forward as one CUDA graph: ~60 launches per step stop costing host time
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_resident()
                side.synchronize()
                engine(x_dev)
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):      # the forward only; the collective stays outside
                    graph_out = engine(x_dev)
            torch.cuda.current_stream().wait_stream(side)

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
            ms = timed(step_graph, args.steps)
            # a replay re-issues every captured launch; count them from one eager step
            l0 = native.launch_count(); step_resident(); launches_per_step = native.launch_count() - l0
        else:
            ms = timed(step_resident, args.steps)
            launches_per_step = (native.launch_count() - launches0) // args.steps
        if rank == 0:
            sampler.window(t_region0, time.perf_counter())
        clocks = sampler.stop() if rank == 0 else None

        # end to end through the public API (bnn_b200.pipeline.HostPipeline): every step uploads the batch from
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
        assert torch.isfinite(last_logits).all()

        # per-launch CUDA-event timing of the binarized path (same data, same stream, warm)
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
            for _ in range(reps):
                engine(x_dev)                      # rank-local: no collective in this pass
            torch.cuda.synchronize()
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
    traffic = measured_traffic()
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
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"resnet18 XNOR-Net ({args.variant}, 19 binarized convs, first/last fp32) {RES}x{RES} bs{B}/GPU",
                   "global_batch": images, "parallelism": f"dp{world}",
                   "l2": f"input batch {B * 3 * RES * RES * 4 / 1e6:.0f} MB + activations exceed the 126 MB L2",
                   "launch": "cuda_graph" if graph is not None else "eager",
                   "fusion": "per-layer" if args.no_fuse else "bnn_b200.fuse.optimize (BN/act/residual/sign in conv epilogues)",
                   "glue": ("torch fp32 (TF32 off): stem conv7x7+BN+ReLU+maxpool, BN/act/add per layer, avgpool, fc"
                            if args.no_fuse else f"stem = {'bnn_stem_mma_fwd' if args.stem == 'mma' else 'bnn_stem_fwd'}; "
                                              "torch fp32 only for global avgpool + fc")},
        "clocks": clocks,
        "e2e": {"value": images / (ms_e2e / args.steps * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": B * 3 * RES * RES * 4 * world, "d2h_bytes_per_step": images * 1000 * 4 * world,
                "ms_per_step": ms_e2e / args.steps,
                "how": "bnn_b200.pipeline.HostPipeline: pinned-host batch -> H2D -> fused engine -> D2H logits, "
                       "double-buffered (upload of step i+1 overlaps the forward of step i)"
                       + (f"; rank 0 host threads and pinned buffers bound to NUMA node {numa_node}" if numa_node is not None else "")},
        "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else 0,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"],
                     "traffic": (traffic["binarized_path_dram_bytes_per_step"] / max(1, n_conv)
                                 if traffic and not args.no_fuse and B == BATCH_PER_GPU else None),
                     "traffic_what": "ncu dram__bytes_read+write of the same launches, per conv launch "
                                     "(profiles/traffic.json; fused engine at bs 256): below the algorithmic bytes "
                                     "because fused layers exchange bit planes instead of fp32 NCHW tensors",
                     "launches": n_conv, "algorithmic_bytes_per_launch": total_bytes / max(1, n_conv),
                     "avg_launch_ms": path_ms / max(1, n_conv), "peak_kind": peak_kind,
                     "what": "bit-pack + XNOR-popcount launches of the 19 binarized layers, algorithmic bytes "
                             f"{total_bytes / 1e9:.3f} GB/step over {path_ms:.3f} ms/step; these layers are "
                             "POPC-bound, see 'popc'"},
        "popc": {"achieved_tbmac_s": total_bmac / (conv_ms * 1e-3) * 1e-12, "peak_tbmac_s": popc_peak_tbmac,
                 "frac": (total_bmac / (conv_ms * 1e-3) * 1e-12 / popc_peak_tbmac) if popc_peak_tbmac else None,
                 "peak_kind": "bnn_ubench(POPC) x 32 on this GPU", "lop3_popc_iadd_gwords_s": mix_gwords,
                 "conv_ms_per_step": conv_ms, "binarized_path_ms_per_step": path_ms},
    }
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
                    "POPC rate); the POPC term is the larger one for every 3x3 layer, so roofline.frac (HBM) is low by "
                    "construction and this object is the one north_star's '>= 60 % of the tighter roofline' refers to"}
    if world == 1 and not args.no_cpu_baseline:
        sample, iters = 64, 10          # ~6-20 s of host work depending on the box
        rate, threads = cpu_floatsim_rate(model_cpu, sample, iters, None)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": f"{iters} forwards of {sample} images (oracle/floatsim.py, torch CPU fp32, "
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
