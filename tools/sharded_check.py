#!/usr/bin/env python3
"""Parity of the table-sharded layout, run under torchrun on N >= 2 GPUs of one node:
     torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py [--full]
Every rank takes part in `prove_with_traces_sharded` with a plan that splits the trace commitments over all ranks; rank 0 also proves
the same segment alone (`prove_with_traces`) and the two AllProofs must agree word for word (proofs, CTL challenges, trace caps).
Small valid segment (restated verifier's input, tests/traces.py) with every table split, then (--full) the bench segment with the
default plan, traces resident in HBM and from host memory."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--precompute", type=int, default=1, help="evaluate the alpha-independent constraint values ahead of the relay (0: fused evaluator in the relay)")
    ap.add_argument("--time", type=int, default=0, help="with --full: time this many sharded proofs and print the phase breakdown of the last")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import zk_evm_b200 as zk
    import bench
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = zk.Context(local)
    comm = zk.TorchComm(device=dev)
    labels = zk.KernelLabels(*bench.LABELS)
    ok = True

    def compare(name, ap, ref):
        nonlocal ok
        same = np.array_equal(ap.ctl_challenges, ref.ctl_challenges) and np.array_equal(np.asarray(ap.trace_caps).ravel(), np.asarray(ref.trace_caps).ravel())
        for t in range(9):
            a, b = ap.stark_proofs[t], ref.stark_proofs[t]
            same = same and ((a is None) == (b is None)) and (a is None or np.array_equal(a, b))
        print("[sharded_check] %-60s %s" % (name, "identical to the one-GPU proof" if same else "MISMATCH"), flush=True)
        ok = ok and same

    # 1. small valid segment, every table split over all ranks
    from tests import traces
    for cfgw, cname in ((bench.TEST_CONFIG, "test config"), (bench.STANDARD_FAST, "standard_fast")):
        cfg = zk.StarkConfig(*cfgw)
        tr = traces.valid_segment(seed=5, k=9)
        in_use = [x is not None for x in tr]
        log_ns = [None if x is None else int(np.log2(x.shape[1])) for x in tr]
        plan = zk.shard_plan(world, log_ns, split_min_bytes=0)
        be = zk.ZkGpuBackend(ctx, cfg, labels, precompute_constraints=bool(args.precompute))
        pv = np.arange(1000, 1037, dtype=np.uint64)
        ap = zk.prove_with_traces_sharded(be, comm, tr, in_use, pv, plan=plan, gather=True)
        if rank == 0:
            compare("valid segment 2^9 (%s), all tables split over %d GPUs" % (cname, world), ap, zk.prove_with_traces(ctx, tr, pv, cfg, labels))
    # 2. the bench segment
    if args.full:
        cfg = zk.StarkConfig(*bench.STANDARD_FAST)
        log_ns = list(bench.SEGMENT_LOG_NS)
        plan = zk.shard_plan(world, log_ns)
        rig = bench.Rig(torch, dev, log_ns, seed=4, mine=[True] * 9 if rank == 0 else plan.needs(rank), common_seed=True)
        be = zk.ZkGpuBackend(ctx, cfg, labels, precompute_constraints=bool(args.precompute))
        for host in (False, True):
            tr = rig.host_traces if host else rig.ptrs
            ap = zk.prove_with_traces_sharded(be, comm, tr, rig.in_use, bench.PUBLIC_VALUES, plan=plan, gather=True)
            if rank == 0:
                ref = zk.prove_with_traces(ctx, None, bench.PUBLIC_VALUES, cfg, labels, device_ptrs=rig.ptrs)
                compare("bench segment, %s traces, plan %s" % ("host" if host else "resident", plan.describe()), ap, ref)
        if args.time:
            import time
            for host in ((False,) if os.environ.get("ZK_SHARD_BEGIN") else (False, True)):
                tr = rig.host_traces if host else rig.ptrs
                for _ in range(2):
                    zk.prove_with_traces_sharded(be, comm, tr, rig.in_use, bench.PUBLIC_VALUES, plan=plan, gather=False)
                torch.cuda.synchronize(); dist.barrier()
                t0 = time.perf_counter()
                each = []
                phs = []
                for _ in range(args.time):
                    t1 = time.perf_counter()
                    be.phase_ms = {"__nosync__": 1}
                    zk.prove_with_traces_sharded(be, comm, tr, rig.in_use, bench.PUBLIC_VALUES, plan=plan, gather=False)
                    each.append(round((time.perf_counter() - t1) * 1e3, 1))
                    phs.append([round(v, 1) for k, v in be.phase_ms.items() if k != "__nosync__"])
                    be.phase_ms = None
                if rank == 0:
                    for row in phs:
                        print("[sharded_check]    host ms between marks:", row, flush=True)
                torch.cuda.synchronize(); dist.barrier()
                dt = (time.perf_counter() - t0) / args.time * 1e3
                print("[sharded_check] rank %d back-to-back proofs, no barrier between: %s" % (rank, each), flush=True)
                be.phase_ms = {}
                zk.prove_with_traces_sharded(be, comm, tr, rig.in_use, bench.PUBLIC_VALUES, plan=plan, gather=False)
                ph, be.phase_ms = be.phase_ms, None
                walls = []
                for _ in range(3):
                    be.phase_ms = {"__nosync__": 1}
                    torch.cuda.synchronize(); dist.barrier()
                    t1 = time.perf_counter()
                    zk.prove_with_traces_sharded(be, comm, tr, rig.in_use, bench.PUBLIC_VALUES, plan=plan, gather=False)
                    walls.append((time.perf_counter() - t1) * 1e3)
                    ph2, be.phase_ms = be.phase_ms, None
                ph2.pop("__nosync__")
                print("[sharded_check] rank %d %s: walls %s; host time between marks WITHOUT syncs: %s" % (rank, "host" if host else "resident", [round(w, 1) for w in walls], {k: round(v, 1) for k, v in ph2.items()}), flush=True)
                print("[sharded_check] rank %d %s traces: %.1f ms per proof (wall, %d proofs); phases of one more proof with a sync at every mark: %s"
                      % (rank, "host" if host else "resident", dt, args.time, {k: round(v, 1) for k, v in ph.items()}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
