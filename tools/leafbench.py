#!/usr/bin/env python3
"""leaf-hash throughput against the number of resident blocks per SM (ZKGPU_LEAF_BLOCKS=5..9, unset = the library's choice)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_evm_b200 as zk
ctx = zk.Context(0)
for ncols, lg in [(2431, 18), (85, 20), (523, 17), (30, 22), (116, 18), (438, 14)]:
    ms = ctx.bench_leaf_hash(ncols, 1 << lg, 3)
    nperm = ((ncols + 7) // 8) * (1 << lg)
    print(json.dumps({"leaf_blocks": os.environ.get("ZKGPU_LEAF_BLOCKS", "auto"), "ncols": ncols, "log_rows": lg, "ms": round(ms, 3),
                      "Mperm_per_s": round(nperm / ms / 1e3, 1)}))
