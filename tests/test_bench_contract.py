"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints one JSON line with the
keys the driver reads, and the B200 arm refuses to run without CUDA instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, timeout=timeout,
                          capture_output=True, text=True)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-batch", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["config"]["global_batch"] == 256 and "bs256/GPU" in line["config"]["workload"]
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("images/sec ResNet-18 XNOR fwd") and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["scaling"] == "weak" and line["dtype"] == "f32" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    from oracle import build as oracle_build
    want_kind = "reference" if os.path.exists(os.path.join(oracle_build.REF_PKG, "__init__.pyc")) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       cwd=ROOT, env=env, timeout=300, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a CUDA-less host")
def test_b200_arm_refuses_to_run_without_cuda():
    r = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and "no CPU fallback" in (r.stderr + r.stdout)


def test_product_package_never_touches_the_oracle():
    """oracle/ is the checker: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
    import re
    pkg = os.path.join(ROOT, "binary-networks-pytorch_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", text, flags=re.M):
                    offenders.append(os.path.relpath(os.path.join(dirpath, name), ROOT))
    assert offenders == []
    shim = open(os.path.join(ROOT, "bnn_b200.py")).read()
    assert "oracle" not in shim
    # bench.py reaches the oracle only inside its CPU-baseline / reference-arm functions
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle import", bench)]
    assert uses, "bench.py's CPU legs are expected to use oracle/"
    for pos in uses:
        func = re.findall(r"^def (\w+)\(", bench[:pos], flags=re.M)[-1]
        # the two builders of the CPU arm's model; their callers are the reference arm, cpu_baseline and the parity CHECK
        assert func in ("build_reference_model", "cpu_twin"), func
    callers = set()
    for m in re.finditer(r"(?<!def )\b(cpu_twin|build_reference_model)\(", bench):
        callers.add(re.findall(r"^def (\w+)\(", bench[:m.start()], flags=re.M)[-1])
    assert callers <= {"build_reference_model", "cpu_twin", "cpu_floatsim_rate", "run_reference", "main"}, callers
    # ... and in main() only inside the parity block, never inside what is timed
    main_src = bench[bench.index("def main()"):]
    assert main_src.count("cpu_twin(") == 1 and main_src.index("cpu_twin(") > main_src.index("parity of what was just timed")
