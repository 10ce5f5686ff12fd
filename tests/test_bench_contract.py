"""CPU test of bench.py's output contract: the reference arm (`--impl reference`, the oracle port timed on the host
cores) prints exactly ONE JSON line on stdout with the keys the driver reads, and the algorithmic byte / flop helpers the
roofline uses agree with SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    sys.path.insert(0, ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("MC-DropBlock fwd passes/s") and d["unit"] == "passes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] >= 1 and d["vs_baseline"] is None
    # kind "reference": the unmodified reference modules (from /root/reference here, from oracle/_ref byte-code on the GPU box)
    from oracle import ref_shims
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shims.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "passes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_algorithmic_work_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    # 500.46 GFLOP per forward, 500.03 of them in the tensor-core convs (everything but the first conv and the head)
    assert abs(bench.CONV_FLOP_PER_FORWARD / 1e9 - 500.03) < 0.05
    # conv family: every 3x3 / transposed conv reads its bf16 input once and writes its bf16 output once
    one, ten = bench.conv_bytes_per_step(1), bench.conv_bytes_per_step(10)
    act = (ten - one) / 9                                                       # activations scale with the batch ...
    assert 62.0e6 < one - act < 62.1e6                                          # ... the 31.0 M bf16 conv weights do not
    assert 0.75e9 < act < 0.78e9 and 7.6e9 < ten < 7.8e9
    eb = bench.elementwise_bytes_per_step(10)
    assert set(eb) >= {"b2u_gn_apply", "b2u_gn_apply_pool", "b2u_head_fwd", "b2u_conv_first_fwd", "b2u_dropblock_dilate"}
    assert abs(eb["b2u_head_fwd"] / 1e6 - 474.3) < 1.0 and abs(eb["b2u_gn_apply_pool"] / 1e6 - 1943.7) < 1.0
