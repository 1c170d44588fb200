"""CPU test of the driver-facing bench.py contract: the reference arm runs without a GPU, prints exactly one JSON line with
the required keys, and non-zero ranks of a torchrun launch exit quietly."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(extra_env, *args):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run_bench({"RANK": "0", "WORLD_SIZE": "1"}, "--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench({"RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_definition_matches_baseline_config():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.M_KL == 100 and bench.R_SHIFTS == 16 and bench.LEVELS == list(range(7))   # BASELINE.json configs[1]
    assert [16 << l for l in bench.LEVELS][-1] == 1024                                      # grid 2^(l+4)
    assert bench.points_per_shift(0) == 1 << 20 and bench.points_per_shift(6) == 2048
    sh = bench.level_shifts()
    assert sh.shape == (7, 16, 100) and (sh >= 0).all() and (sh < 1).all()


def test_committed_bench_lines_carry_every_contract_key():
    """the round's measured bench lines under profiles/ (N = 1, 2, 4, 8): every key of the bench contract, consistent values"""
    import json
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    single = None
    for n, fn in ((1, "r01_final_bench_1gpu.json"), (2, "r01_bench_2gpu.json"), (4, "r01_bench_4gpu.json"), (8, "r01_bench_8gpu.json")):
        d = json.loads(open(os.path.join(ROOT, "profiles", fn)).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in d, (fn, key)
        assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
        assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert str(base.get("metric", d["metric"])).split()[0].lower() in d["metric"].lower() or "samples" in d["metric"]
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0 < r["frac"] < 1
        c = d["clocks"]
        assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        per_step = d["config"]["samples_per_step_per_gpu"] * n
        assert abs(d["value"] - per_step / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
        if n == 1:
            single = d["value"]
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
        else:
            assert 0.95 * n * single <= d["value"] <= 1.05 * n * single      # weak scaling


def test_round2_bench_lines_carry_every_contract_key():
    """round 2: the N = 1 and N = 8 lines under profiles/, with the extra keys this round added (strong scaling, C3 / C4 / C5,
    run-level end-to-end, pipe-busy and traffic figures read from the committed ncu summary)"""
    one = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_1gpu.json")).read().strip().splitlines()[-1])
    eight = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_8gpu.json")).read().strip().splitlines()[-1])
    for n, d in ((1, one), (8, eight)):
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "extra"):
            assert key in d, (n, key)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
        assert d["warmup"] >= 3 and d["gpu_launches"] == 7 * d["steps"]      # one launch per level and step, no finalize launch
        r = d["roofline"]
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0.3 < r["frac"] < 1
        assert 0 < d["e2e"]["value"] <= d["value"] * 1.02 and d["e2e"]["h2d_bytes_per_step"] > 0
        for cfg in ("c3", "c4", "c5"):
            assert d["extra"][cfg]["value"] > 0, cfg
        per_step = d["config"]["samples_per_step_per_gpu"] * n
        assert abs(d["value"] - per_step / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert 0.95 * 8 * one["value"] <= eight["value"] <= 1.05 * 8 * one["value"]
    assert eight["strong"]["scaling"] == "strong" and eight["strong"]["value"] > 5 * one["value"]
    assert len(one["roofline"]["pipe_busy_ncu_per_level"]) == 7 and all(0.4 < x < 1 for x in one["roofline"]["pipe_busy_ncu_per_level"])
    assert one["roofline"]["traffic"] > 0 and one["cpu_baseline"]["kind"] == "port"
    tols = one["run_e2e"]["tols"]
    assert len(tols) >= 3 and all(t["same_decisions"] for t in tols) and tols[-1]["speedup"] > 10
