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


def test_committed_round2_lines_carry_the_contract_keys():
    """This round's B200 lines (the final kernel, one GPU and eight): every contract key, a roofline whose fraction follows from
    its own numbers and whose kernel time fits the step, a passed precision gate, a verified read-back and — at N > 1 — the
    exchanged frame checked against one GPU rendering alone."""
    for name, n, cfg in (("r02_bench_c2_n1_final.json", 1, "c2"), ("r02_bench_c4_n1_final.json", 1, "c4"), ("r02_bench_c3_n1_final.json", 1, "c3"),
                         ("r02_bench_c2_n8_final.json", 8, "c2"), ("r02_bench_c4_n8_final.json", 8, "c4")):
        d = _last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
        assert BASE_KEYS | {"roofline", "clocks", "precision_gate", "exact", "exchange_check"} <= set(d), name
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["data"] == "synthetic" and d["config"]["name"] == cfg
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel_ms", "kernel_ms_per_launch", "frames_per_launch"} <= set(r)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "hbm"
        assert r["kernel_ms"] <= d["ms_per_step"] * 1.08, name              # consecutive launches overlap a little in the drain
        g = d["precision_gate"]
        assert g["tolerance"] == 1e-6 and g["passed"] == (max(g["per_channel_mse_fast_vs_exact"]) < 1e-6)
        assert ("fast" in d["config"]["precision"][:5]) == g["passed"]       # the fast build is quoted only inside the tolerance
        assert (d["exact"] is None) == (not g["passed"])
        e = d["e2e"]
        assert e["value"] > 0 and e["last_frame_on_host_equals_device_image"] is True
        assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["exchange_check"] is None
        else:
            assert d["exchange_check"]["exchange_equals_single_gpu"] is True
