#!/usr/bin/env python3
"""Peak device memory of one segment proof against SegmentProver's admission estimate (zk_evm_b200.scheduler.estimate_segment_bytes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zk_evm_b200 as zk
from bench import SEGMENT_CONFIGS, NUM_COLUMNS, PUBLIC_VALUES

P = 0xFFFFFFFF00000001
for name in ("b3_b6", "b19807080"):
    log_ns = SEGMENT_CONFIGS[name]
    rng = np.random.default_rng(7)
    traces = [rng.integers(0, P, size=(NUM_COLUMNS[t], 1 << lg), dtype=np.uint64) for t, lg in enumerate(log_ns)]
    ctx = zk.Context(0)
    zk.prove_with_traces(ctx, traces, PUBLIC_VALUES, zk.StarkConfig.standard_fast(), zk.KernelLabels(1, 2, 3, 4))
    st = ctx.stats()
    est = zk.estimate_segment_bytes(traces)
    print(json.dumps({"segment": name, "log_ns": list(log_ns), "trace_bytes": int(sum(t.nbytes for t in traces)), "bytes_peak": int(st["bytes_peak"]),
                      "estimate": est, "estimate_over_peak": round(est / st["bytes_peak"], 3)}))
    ctx.close()
