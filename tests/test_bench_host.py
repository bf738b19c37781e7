"""bench.py without a GPU: the reference arm runs (on a shrunk segment) and prints the contract's line; both arms name the workload with
the same `config` object; the synthetic stream of config #5 is reproducible and inside the CI ranges."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    a = types.SimpleNamespace(workload="segment", config="b19807080", stark_config="standard_fast", shrink=0, log_n=20)
    a.__dict__.update(kw)
    return a


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shrink", "8", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                  # exactly one JSON line on stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-9 * d["value"] + 1e-12      # value and ms_per_step are the same measurement
    # the same `config` object our arm prints for these flags
    a = _args(shrink=8)
    assert d["config"] == bench.config_record(a, bench.segment_shape(a))
    # the other ranks of a torchrun launch exit without work and without output
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--shrink", "8", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_shapes_follow_the_survey():
    assert bench.segment_shape(_args()) == [17, 14, 19, 17, 13, 16, 21, 19, 19]
    assert bench.segment_shape(_args(config="b3_b6")) == [16, 10, 16, 12, 8, 10, 18, 16, 7]           # config #3
    assert bench.segment_shape(_args(workload="cpu_table")) == [None, None, 20, None, None, None, None, None, None]   # config #2
    assert "witness_b3_b6" in bench.config_record(_args(config="b3_b6"), bench.segment_shape(_args(config="b3_b6")))["workload"]
    hs = bench.stream_heights()
    assert hs == bench.stream_heights() and len(hs) == 64                                               # config #5: seeded, reproducible
    for h in hs:
        assert all(lo <= v < hi for v, (lo, hi) in zip(h, bench.GENERIC_RANGES))                        # Rust half-open ranges lo..hi
    assert len(set(hs)) > 32                                                                            # varied heights, not one shape
    assert len(bench.PUBLIC_VALUES) == 2217                                                             # flatten_public_values' length
