#!/usr/bin/env python3
"""Stage spans (CUDA events on the library stream, zkgpu_ctx_set_timing) of ONE full-size segment proof with one segment in flight:
where a single segment's latency goes, table by table.  python tools/stage_report.py [--shrink k] [--config b19807080|b3_b6]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shrink", type=int, default=0)
    ap.add_argument("--config", default="b19807080")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import zk_evm_b200 as zk
    heights = bench.SEGMENT_CONFIGS[args.config] if hasattr(bench, "SEGMENT_CONFIGS") else bench.SEGMENT_LOG_NS
    log_ns = [max(4, lg - args.shrink) for lg in heights]
    dev = torch.device("cuda", 0)
    ctx = zk.Context(0)
    cfg, labels = zk.StarkConfig(*bench.STANDARD_FAST), zk.KernelLabels(*bench.LABELS)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    tr = [torch.randint(0, 2 ** 63 - 1, (bench.NUM_COLUMNS[t], 1 << log_ns[t]), dtype=torch.int64, device=dev, generator=g) for t in range(9)]
    ptrs = [(d.data_ptr(), d.shape[1]) for d in tr]
    torch.cuda.synchronize()
    for _ in range(2):
        zk.prove_with_traces(ctx, None, bench.PUBLIC_VALUES, cfg, labels, device_ptrs=ptrs)
    best = None
    for _ in range(args.reps):
        ctx.set_timing(True)
        ctx.timer_start()
        zk.prove_with_traces(ctx, None, bench.PUBLIC_VALUES, cfg, labels, device_ptrs=ptrs)
        total = ctx.timer_stop()
        rep = ctx.timing_report()
        ctx.set_timing(False)
        if best is None or total < best[0]:
            best = (total, rep)
    total, rep = best
    print("total %.2f ms (one segment in flight)" % total)
    agg = {}
    for name, ms in rep:
        print("%-28s %9.3f" % (name, ms))
        agg[name] = agg.get(name, 0.0) + ms
    print("---- by stage name")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print("%-28s %9.3f" % (k, v))
    print(json.dumps({"total_ms": total, "log_ns": log_ns, "spans": rep}))
    ctx.close()


if __name__ == "__main__":
    main()
