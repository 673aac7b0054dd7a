"""The bench line is a contract with the driver: keys of the committed GPU run (profiles/) and of a live run of the
reference arm (CPU only, one bounded sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_committed_bench_line_has_every_contract_key():
    j = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_final.json")))
    assert BASE_KEYS <= set(j)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "numeric-factorization" in base["metric"] and j["metric"] == "fp64_numeric_factorization_gflops"
    assert j["unit"] == "GFLOP/s" and j["higher_is_better"] is True and j["dtype"] == "f64" and j["vs_baseline"] is None
    assert j["config"]["workload"].startswith("cfg2") and "model" not in j["config"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"]) and not j["clocks"]["reasons"]
    assert j["gpu_launches"] > 0
    e = j["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 8e9 and 0 < e["value"] < j["value"]
    r = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0 < r["frac"] < 1 and r["traffic"] > 0
    assert "SparseFrame_factorize" in e["api"]                 # the reference-facing symbol of the drop-in library
    c = j["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "reference" and c["cores"] <= 16
    assert c["sample"].startswith("cfg2: the whole matrix") and {"reference_policy_threads", "used_threads"} <= set(c["threads"])
    assert j["value"] == abs(j["value"]) and abs(j["value"] - j["config"]["flops_per_step"] / j["ms_per_step"] / 1e6) < 1e-6 * j["value"]


def test_committed_multi_gpu_lines_carry_parity_and_the_same_config_baseline():
    """N > 1: the distributed factor is checked inside the bench run, and the one-GPU point of the curve is measured on the
    same matrix in the same run."""
    for name, n in (("r02_bench_n2_cfg2_v2_streamed_e2e.json", 2), ("r02_bench_n4_cfg5_v1.json", 4), ("r02_bench_n8_cfg5_v1.json", 8)):
        j = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert BASE_KEYS <= set(j) and j["n_gpus"] == n and j["scaling"] == "strong"
        p = j["parity"]
        assert p["ok"] is True and p["backward"] <= 1e-12 and p["max_rel_vs_single_gpu"] <= 1e-10
        assert p["entries_gathered"] == p["entries_expected"] > 1e9
        s1 = j["single_gpu_same_config"]
        assert s1["ms_per_step"] > j["ms_per_step"] and 1.0 < s1["ms_per_step"] / j["ms_per_step"] <= n
        x = j["exchange"]
        assert x["remote_update_GB_per_step"] > 0 and x["pushed_GB_per_step"] > 0 and len(x["image_GB_per_rank"]) == n
        assert max(x["image_GB_per_rank"]) < j["config"]["factor_GB"]            # compact images: nobody holds the whole factor
        assert j["e2e"]["d2h_bytes_per_step"] >= 8 * p["entries_expected"] * 0.99 and j["e2e"]["value"] < j["value"]


def test_reference_arm_prints_one_json_line_with_the_same_metric():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsfref_chol.so")):
        import pytest
        pytest.skip("compiled reference not built")
    # cfg1 here (2 s per step on 8 cores); the driver's run is cfg2, the config of the GPU line
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--config", "cfg1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout                       # the reference's own printf goes to stderr
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and BASE_KEYS <= set(j)
    assert j["metric"] == "fp64_numeric_factorization_gflops" and j["unit"] == "GFLOP/s" and j["higher_is_better"] is True
    assert j["config"]["workload"].startswith("cfg1") and j["config"]["same_config_as_gpu_arm"] is True
    assert j["steps"] == 2 and j["warmup"] == 1
    assert abs(j["config"]["sample"].count("1.923e+10") - 1) == 0          # flops of cfg1, from the reference's arrays
    assert j["e2e"] == {"value": j["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = j["cpu_baseline"]
    assert c["kind"] == "reference" and c["value"] == j["value"] and 1 <= c["cores"] <= 16 and "the whole matrix" in c["sample"]
    assert {"host_cores", "affinity_cores", "loadavg_1_5_15"} <= set(c["host"])
    assert 1.0 < j["value"] < 5000.0                       # GFLOP/s of a CPU, not of a GPU

    # the other ranks of a torchrun launch print nothing and exit 0
    q = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert q.returncode == 0 and q.stdout.strip() == ""
