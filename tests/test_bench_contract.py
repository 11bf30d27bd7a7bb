"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) times the CPU oracle and prints
ONE JSON line with the keys the driver reads; the algorithmic-byte table matches DESIGN.md's formulas."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-edges", "20000"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "DMPNN layer fwd+bwd edges/sec" and j["unit"] == "edges/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_kernel_bytes_table():
    sys.path.insert(0, ROOT)
    import bench
    N, E, H = 2_000_000, 40_000_000, 128
    row = 4 * H
    assert bench.kernel_bytes("segment_reduce.dQd_bwd", N, E, H) == E * row + N * row + 4 * E + 4 * (N + 1)
    # U = S + coef*P arrives pre-combined from the dual projection: read U, write out, two endpoint rows per edge
    # (one per edge when mirrored pairs share them)
    assert bench.kernel_bytes("edge_update", N, E, H) == 4 * E * row + 12 * E + row
    assert bench.kernel_bytes("edge_update", N, E, H, mirrored=True) == 3 * E * row + 12 * E + row
    assert bench.kernel_bytes("edge_backward", N, E, H) == 2 * E * row + 5 * E
