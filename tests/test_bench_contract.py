"""bench.py's output contract (the driver parses ONE JSON line): the reference arm is run for real on a tiny config;
the GPU arm's keys are checked on the committed measurement of this build (profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _last_json_line(text):
    lines = [ln for ln in text.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _last_json_line(r.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "relgat_train_edges_per_sec_fwd_bwd" and d["unit"] == "edges/s" and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--config", "tiny"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r01_bench_c2_fp32_final.json")) as f:
        d = _last_json_line(f.read())
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    roof = d["roofline"]
    assert roof["bound"] in ("hbm", "tensor") and roof["unit"] in ("GB/s", "TFLOP/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-3 and roof["traffic"] is None or roof["traffic"] > 0
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] != d["value"]
    assert d["gpu_launches"] > 0 and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    for name in ("r01_bench_n2_peer.json", "r01_bench_n4_peer.json", "r01_bench_n8_peer.json"):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            m = _last_json_line(f.read())
        assert BASE_KEYS <= set(m) and m["scaling"] == "weak" and m["config"]["exchange"] == "peer"
        assert m["n_gpus"] == int(name.split("_n")[1][0])
