"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`: the CPU restatement of the reference's
algorithm on the host cores) prints exactly one JSON line with the agreed keys, and nothing that runs on the GPU box reads the reference tree."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys(oracle):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "chain_iters_per_sec" and d["unit"] == "chain-iterations/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 1e3 and "workload" in d["config"] and "seeds" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d.get("gpu_launches", 0) == 0


def test_nothing_that_runs_on_the_gpu_box_reads_the_reference_tree():
    # /root/reference does not exist on the GPU box: bench.py, the driver hooks, the product and the GPU tests must not open it
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += [os.path.join(ROOT, "tests", f) for f in os.listdir(os.path.join(ROOT, "tests")) if f.startswith("test_gpu_")]
    pkg = os.path.join(ROOT, "mamba.jl_b200")
    for base, _, names in os.walk(pkg):
        if "build" in base.split(os.sep):
            continue
        files += [os.path.join(base, n) for n in names if n.endswith((".py", ".cu", ".cuh", ".hpp", ".h"))]
    for f in files:
        src = open(f, errors="ignore").read()
        assert "/root/reference" not in src, f
    # the golden tests that do look at the reference tree are CPU tests and skip when it is absent
    src = open(os.path.join(ROOT, "tests", "test_api_host.py")).read()
    assert re.search(r"if not os\.path\.exists\(.*\):\s*\n\s*pytest\.skip", src)


def test_own_arm_declares_the_contract_keys():
    # static: the JSON line of the GPU arm is assembled from these keys (the run itself needs a GPU: profiles/r1_seeds_bench.json)
    src = open(os.path.join(ROOT, "bench.py")).read()
    for k in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"', '"scaling"', '"vs_baseline"',
              '"dtype"', '"data"', '"config"', '"roofline"', '"bound"', '"achieved"', '"peak"', '"frac"', '"traffic"', '"cpu_baseline"', '"e2e"',
              '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"gpu_launches"', '"clocks"', '"sm_mhz"', '"sm_max_mhz"', '"reasons"'):
        assert k in src, k
    line = json.loads(open(os.path.join(ROOT, "profiles", "r1_seeds_bench.json")).read().strip().splitlines()[-1])
    assert line["roofline"]["frac"] == line["roofline"]["achieved"] / line["roofline"]["peak"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0 and line["clocks"]["reasons"] == []
