"""bench.py's reference arm on the host cores: the JSON line the driver parses (prompt section 4: same metric, unit,
config and higher_is_better as the GPU arm, plus impl / cpu_baseline / e2e with zero transfer bytes).  The GPU arm
needs a device; its line is checked by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--seconds", "5"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] and d["metric"].split(" at ")[0] in base["metric"]
    assert d["unit"] == "audio-s/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # `config` is the workload and nothing else: the very object the GPU arm prints for the same command line
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("q4_44k_2ch", 64, 5.0)
    assert d["config"]["streams_per_gpu"] == 64 and d["config"]["seconds_per_stream"] == 5.0


def test_reference_arm_runs_on_rank_0_only():
    """under torchrun the other ranks exit 0 without work and without output"""
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
