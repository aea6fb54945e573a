"""The bench.py contract (CPU): the reference arm runs here and prints one well-formed JSON line; the committed B200 lines
under profiles/ carry every key the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def _last_json_line(text):
    lines = [ln for ln in text.strip().splitlines() if ln.startswith("{")]
    assert lines, text
    return json.loads(lines[-1])


def test_reference_arm_runs_on_the_cpu_and_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    d = _last_json_line(out.stdout)
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    from oracle import ref as R
    assert cb["kind"] == ("reference" if R.available() else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_a_non_zero_rank_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_b200_lines_carry_the_contract_keys():
    for name, n in (("r01_bench_v4_n1_rgb_e2e.json", 1), ("r01_bench_v4_n2_scatter_e2e.json", 2), ("r01_bench_v4_n4_scatter_e2e.json", 4)):
        d = _last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
        assert BASE_KEYS | {"roofline", "clocks"} <= set(d), name
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["data"] == "synthetic"
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        e = d["e2e"]
        assert e["value"] > 0 and e["d2h_bytes_per_step"] == 1920 * 1080 * 12 and e["last_frame_on_host_equals_device_image"] is True
        assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
