#!/usr/bin/env python3
"""Per-kernel device timings (CUDA events inside libzkgpu) for the roofline tables in DESIGN.md / profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_evm_b200 as zk

PEAK = 6462.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

ctx = zk.Context(0)
rows = []
for ncols, lg in [(128, 16), (85, 20), (16, 21), (512, 14), (2431, 12)]:
    ms, launches = ctx.bench_ntt(ncols, 1 << lg, 5)
    gb = 16.0 * ncols * (1 << lg) / 1e9
    rows.append({"kernel": "ntt_dif", "ncols": ncols, "log_n": lg, "ms": ms, "launches": launches,
                 "algo_GBps": gb / ms * 1e3, "frac_of_measured_hbm": gb / ms * 1e3 / PEAK})
for ncols, lg in [(85, 21), (30, 22), (8, 21), (135, 18), (2431, 14)]:
    ms = ctx.bench_leaf_hash(ncols, 1 << lg, 3)
    nperm = ((ncols + 7) // 8) * (1 << lg)
    gb = (8.0 * ncols + 32) * (1 << lg) / 1e9
    rows.append({"kernel": "leaf_hash", "ncols": ncols, "log_rows": lg, "ms": ms, "perms_per_s": nperm / ms * 1e3,
                 "algo_GBps": gb / ms * 1e3, "frac_of_measured_hbm": gb / ms * 1e3 / PEAK})
for lg in [21, 16]:
    ms = ctx.bench_merkle_levels(1 << lg, 3)
    rows.append({"kernel": "merkle_levels", "log_leaves": lg, "ms": ms, "perms_per_s": ((1 << lg) - 16) / ms * 1e3,
                 "algo_GBps": 96.0 * (1 << lg) / 1e9 / ms * 1e3})
for r in rows:
    print(json.dumps(r))
