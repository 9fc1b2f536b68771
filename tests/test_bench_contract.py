"""bench.py's CPU-runnable parts: the reference arm prints ONE JSON line with the contract's keys, and the workload table is
consistent with BASELINE.json's configs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workloads_follow_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    import mgcfd_b200 as M
    kind, dims, variant, _ = bench.WORKLOADS["c2"]           # Onera-M6-shaped: 300K/165K/111K/81K nodes, ~930K fine edges, 4 levels
    nodes = [d[0] * d[1] * d[2] for d in dims]
    assert len(dims) == 4 and variant == M.MESH_M6_WING
    for got, want in zip(nodes, (300e3, 165e3, 111e3, 81e3)):
        assert abs(got - want) / want < 0.03
    e0 = 3 * dims[0][0] ** 2 * (dims[0][0] - 1)
    assert abs(e0 - 930e3) / 930e3 < 0.06
    kind, dims, variant, _ = bench.WORKLOADS["c1"]           # fvcorr.domn.097K-shaped: single level, ~97K cells
    assert len(dims) == 1 and variant == M.MESH_FVCORR and abs(6 * dims[0][0] * dims[0][1] * dims[0][2] - 97e3) / 97e3 < 0.02
    kind, dims, variant, _ = bench.WORKLOADS["c3"]           # 8M-node tetrahedral box, 4 levels
    assert len(dims) == 4 and kind == M.GEN_TET_BOX and abs(dims[0][0] ** 3 - 8e6) / 8e6 < 0.03
    assert bench.units_per_cycle([(0, 10, 0, 0), (0, 5, 0, 0), (0, 2, 0, 0)]) == 3 * (10 + 2 * 5 + 2)
