"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"], env={"BK_REF_BUDGET_S": "5"})
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "target_regions_assembled_per_sec" and d["unit"] == "regions/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    # both arms describe the workload with the same config object
    import bench
    assert d["config"] == bench.config_dict("C1", 1)


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(["--regions", "60", "--steps", "8", "--warmup", "3", "--no-cpu-baseline"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "target_regions_assembled_per_sec" and d["n_gpus"] == 1 and d["steps"] == 8 and d["warmup"] == 3
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["from_files"]["value"] > 0 and d["from_files"]["n_contigs"] > 0
    assert d["with_ref_kmer_cache"]["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["sharding_check"]["equal_to_single_gpu_run"] is True and d["sharding_check"]["regions"] == 60
    assert d["e2e_dropin"]["value"] > 0 and d["run"]["host_threads_per_rank"] == 1
    assert all(v["frac"] is None or v["frac"] >= 0 for v in d["roofline_per_kernel"].values())
