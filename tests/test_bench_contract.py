"""bench.py contract (task statement, sections "Measurement" and "How to work"), checked on the CPU:
the reference arm really runs here (it is the reference's CPU implementation, no GPU involved), and
the committed bench lines under profiles/ carry every key the driver and the judge read."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config"}


def _line(path):
    with open(path) as f:
        return json.loads([l for l in f if l.startswith("{")][0])


def test_reference_arm_runs_on_the_host_and_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--ref-step", "200000"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one JSON line"
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"] == "incr_mops_c2" and d["unit"] == "Mops/s" and d["higher_is_better"] is True
    assert "workload" in d["config"] and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_lines_carry_the_contract_keys(n):
    d = _line(os.path.join(ROOT, "profiles", f"r1_bench_n{n}.json"))
    assert BASE_KEYS <= set(d)
    assert d["metric"] == "incr_mops_c2" and d["unit"] == "Mops/s" and d["n_gpus"] == n
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u32"
    assert d["warmup"] >= 3 and abs(d["ms_per_step"] * d["steps"] * d["value"] * 1e3
                                    - d["config"]["timed_ops"]) < 1e-6 * d["config"]["timed_ops"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    k = d["checks"]                   # full-size invariants: every op added 1, odd queries all miss
    assert k["value_sum_ok"] and k["hit_fraction_exact"] and k["rows_ok"] and k["value_sum"] == k["ops_applied"]
    if n == 1:
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] == 8 * d["config"]["ops_per_step"] and e["d2h_bytes_per_step"] > 0
        assert 0 < e["value"] < d["value"]         # host buffers can only be slower than resident ones
        cb = d["cpu_baseline"]
        assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        assert r["random_sector"]["incr_frac"] > 0.4          # north_star: >= 40 % of the random-sector roofline


ROUND2 = [("r2_bench_n1.json", "incr_mops_c2", 1), ("r2_bench_n1_steps20.json", "incr_mops_c2", 1),
          ("r2_bench_c3_n1.json", "incr_mops_c3", 1), ("r2_bench_c4_n1.json", "getrow_mpairs_c4", 1),
          ("r2_bench_n2.json", "incr_mops_c2", 2), ("r2_bench_n8.json", "incr_mops_c2", 8),
          ("r2_bench_c5_n2.json", "incr_mops_c5", 2), ("r2_bench_c5_n8.json", "incr_mops_c5", 8),
          # the final build of the round (256-slice write chunks, slice-ordered point reads)
          ("r2h_bench_n1_steps20.json", "incr_mops_c2", 1), ("r2h_bench_c3_n1.json", "incr_mops_c3", 1)]


@pytest.mark.parametrize("name,metric,n", ROUND2)
def test_round2_bench_lines(name, metric, n):
    """Every committed round-2 line: the contract keys, zero parity mismatches against the CPU reference
    (incl. sharded getrow at N > 1), the size-independent checks, roofline + e2e + cpu_baseline present."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not recorded yet")
    d = _line(path)
    assert BASE_KEYS <= set(d) and d["metric"] == metric and d["n_gpus"] == n
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u32" and d["warmup"] >= 3
    assert "workload" in d["config"] and d["gpu_launches"] > 0 and d["value"] > 0
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    p = d["parity"]
    assert p["mismatches"] == 0 and p["ranks"] == n and p["checker"] == "reference"
    assert p["gets"] > 0 and p["rowlens"] > 0 and p["getrow_pairs"] > 0
    assert all(v for k, v in d["checks"].items() if k.endswith("_ok") or k.endswith("_exact"))
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    if n > 1:
        assert r["nvlink"]["remote_bytes_per_rank_per_step"] > 0 and d["scaling"] in ("weak", "strong")
    if metric == "incr_mops_c2" and n == 1:
        assert r["random_sector"]["incr_frac"] > 0.4          # north_star: >= 40 % of the random-sector roofline
    if name.startswith("r2h_"):
        g = r["get"]        # the slice-ordered look-ups answered every query, with the answers of the input order, faster
        assert g["sliced_fraction"] == 1.0 and g["input_order"]["same_answers"] and d["checks"]["get_orders_agree"]
        assert d["get_mops"] > 1.5 * g["input_order"]["get_mops"]


def test_full_scale_parity_record():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_fullscale_parity.json")))
    assert d["ok"] and d["row_digest_mismatches"] == 0 and d["get_mismatches"] == 0
    assert d["rows_checked"] == 13_000_000 and d["pairs_compared"] == d["gpu_nnz"] == 1_510_576_950
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_fullscale_parity.json")))
    assert d["ok"] and d["row_digest_mismatches"] == 0 and d["get_mismatches"] == 0
    assert d["rows_checked"] == 13_000_000 and d["pairs_compared"] == d["gpu_nnz"] == 1_510_576_950
