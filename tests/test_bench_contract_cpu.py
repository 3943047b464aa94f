"""The bench line contract (task statement, "Measurement"): the committed lines of the final state carry every key the driver
parses, with consistent values.  (The lines themselves are produced on a B200: `tools/collect_r7_final.sh`.)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_headline_line_has_every_contract_key():
    d = _load("r7_bench_n1.json")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    # BASELINE.json names the metric in prose ("train-step frames/sec @512x512, 30k Gaussians, ..."); the line carries its slug
    assert "frames/sec" in base["metric"] and "30k Gaussians" in base["metric"]
    assert d["metric"] == "train_step_frames_per_sec_512x512_30k_gaussians" and d["unit"] == "frames/s"
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = frames of all ranks / time
    frames = d["config"]["global_batch"]
    assert abs(d["value"] - frames / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == d["unit"] and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is None or r["traffic"] > 0
    b = d["cpu_baseline"]
    assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"] and b["unit"] == d["unit"]
    # the extra keys this repo adds
    for k in ("b1", "full_model", "full_model_b1", "roofline_blend", "roofline_conv", "library_share_of_step", "kernels"):
        assert k in d, k
    assert d["library_share_of_step"]["own_kernels"] > 0.95


def test_reference_arm_line_has_every_contract_key():
    d, h = _load("r7_bench_reference_arm.json"), _load("r7_bench_n1.json")
    assert d["impl"] == "reference"
    assert d["metric"] == h["metric"] and d["unit"] == h["unit"] and d["higher_is_better"] == h["higher_is_better"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_multi_gpu_lines_report_whole_job_throughput():
    for name, n in (("r7_bench_n2.json", 2), ("r7_bench_n8.json", 8)):
        d = _load(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert d["config"]["global_batch"] == n * d["config"]["frames_per_step_per_gpu"]
        assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
        assert "strong_scaling" in d
