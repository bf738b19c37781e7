#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 STARK proving path (contract: task statement / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the zkgpu C ABI)
  python bench.py --impl reference --gpus N ...            # reference arm: the CPU restatement (oracle) on all host cores

Metric (BASELINE.json): segment proofs/sec.  A step = one segment proof PER SEGMENT STREAM (--streams S, default 3: S segments in
flight per GPU, each on its own context + CUDA stream + host thread, proving its segments back to back); a segment proof =
prove_with_traces over all nine STARK tables of a synthetic segment with the table heights of the `witness_b19807080` CI ranges: per table the trace
commitment, CTL / lookup auxiliary columns + commitment, fused quotient evaluation + commitment, openings, FRI commit phase,
proof of work and query answers, all tables chained through one Fiat-Shamir transcript.
  value : proofs/s with the traces already resident in HBM, device time (CUDA events on the library's stream), max over ranks
  e2e   : the same call with the traces in pinned host memory (H2D inside the timed region) and the proofs read back (D2H)
N > 1 (one process per GPU, torchrun):
  --parallelism segments (default)  every rank proves its own segment — segments are independent proofs (own transcript), the
                                    way the reference spreads them over workers; no data-path collective; "weak" scaling
  --parallelism tables              the north-star layout: the nine tables of ONE segment spread over the ranks, one all-gather
                                    of trace caps + the transcript relay (DESIGN.md "Multi-GPU"); "strong" scaling
--workload cpu_table  times BASELINE config #2 (single CpuStark table, 2^20 rows) instead.
"""
import argparse
import contextlib
import ctypes as C
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_CPU = 2
TABLE_NAMES = ("Arithmetic", "BytePacking", "Cpu", "Keccak", "KeccakSponge", "Logic", "Memory", "MemBefore", "MemAfter")
NUM_COLUMNS = (116, 71, 85, 2431, 438, 523, 30, 12, 12)
# top of the CI height ranges of artifacts/witness_b19807080.json (scripts/prove_stdio.rs:89-101)
SEGMENT_LOG_NS = (17, 14, 19, 17, 13, 16, 21, 19, 19)
STANDARD_FAST = (100, 2, 1, 4, 16, 4, 5, 84)
LABELS = (0x1234, 0x77, 0x4000, 0x5000)
PUBLIC_VALUES = np.arange(1, 2180, dtype=np.uint64)       # ~2.2k observed elements (get_challenges.rs:202-227), synthetic
BG = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
STATE0 = np.arange(1, 13, dtype=np.uint64)
METRIC = "segment proofs/sec"
# DRAM bytes per algorithmic byte of the dominant kernel (leaf_hash), from the committed `ncu --set full` capture
# profiles/r1j_ncu_leaf_hash.raw.csv: Keccak trace leaves, dram__bytes_read.sum + dram__bytes_write.sum = 5.1175 GB + 0.0162 GB for
# (8 * 2431 + 32) * 2^18 = 5.1066 GB algorithmic (every LDE column read once, one digest written per row)
LEAF_HASH_TRAFFIC_PER_ALGORITHMIC_BYTE = 5.13375 / 5.10657
FAMILIES = ("leaf_hash", "merkle_levels", "ntt", "quotient", "aux_columns", "openings", "fri", "pow")


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def segment_shape(args):
    if args.workload == "cpu_table":
        return [args.log_n if t == T_CPU else None for t in range(9)]
    return [max(4, lg - args.shrink) for lg in SEGMENT_LOG_NS]


def describe(log_ns):
    return ", ".join("%s 2^%d x %d" % (TABLE_NAMES[t], lg, NUM_COLUMNS[t]) for t, lg in enumerate(log_ns) if lg is not None)


def kernel_stats(zk, ctx):
    out = {}
    for i, name in enumerate(FAMILIES):
        l, ms, b = C.c_uint64(), C.c_double(), C.c_double()
        zk._lib.check(zk.lib().zkgpu_ctx_kernel_stats(ctx._h, C.c_uint32(i), C.byref(l), C.byref(ms), C.byref(b)))
        out[name] = {"launches": l.value, "ms": ms.value, "bytes": b.value}
    return out


def run_ours(args):
    import torch
    import zk_evm_b200 as zk
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = zk.Context(local_rank)
    # --streams S: S segments in flight per GPU, each on its own context (own CUDA stream) driven by its own host thread (ctypes
    # releases the GIL): the latency-bound tails (small Merkle levels, transcript round trips) of one overlap the kernels of another
    nstreams = 1 if (args.parallelism == "tables" and world > 1) else max(1, args.streams)
    ctxs = [ctx] + [zk.Context(local_rank) for _ in range(nstreams - 1)]
    cfg = zk.StarkConfig(*STANDARD_FAST)
    labels = zk.KernelLabels(*LABELS)
    log_ns = segment_shape(args)
    in_use = [lg is not None for lg in log_ns]
    sharded = args.parallelism == "tables" and world > 1
    owner = zk.default_owner(world, weights=[(0 if lg is None else (1 << lg) * NUM_COLUMNS[t] * ((NUM_COLUMNS[t] + 7) // 8 + 6))
                                             for t, lg in enumerate(log_ns)]) if sharded else [rank] * 9
    mine = [in_use[t] and owner[t] == rank for t in range(9)]

    # synthetic traces: uniform random canonical field elements (seed 4 + rank; with rate_bits = 1 the prover does the
    # same work on any trace), resident in HBM and mirrored in pinned host memory for the e2e leg
    g = torch.Generator(device=dev)
    g.manual_seed(4 + (0 if sharded else rank))
    dev_traces, host_traces = [None] * 9, [None] * 9
    for t in range(9):
        if not in_use[t]:
            continue
        x = torch.randint(0, 2 ** 63 - 1, (NUM_COLUMNS[t], 1 << log_ns[t]), dtype=torch.int64, device=dev, generator=g)
        if mine[t]:
            dev_traces[t] = x
            h = torch.empty(x.shape, dtype=torch.int64, pin_memory=True)
            h.copy_(x)
            host_traces[t] = h.numpy().view(np.uint64)
        del x
    torch.cuda.synchronize()
    ptrs = [None if d is None else (d.data_ptr(), d.shape[1]) for d in dev_traces]
    h2d_bytes = sum(8 * NUM_COLUMNS[t] * (1 << log_ns[t]) for t in range(9) if mine[t])
    comm = zk.TorchComm(device=dev) if sharded else None
    backend = zk.ZkGpuBackend(ctx, cfg, labels)
    d2h = [0]

    # --finish-on-device: the Keccak and Logic traces (72 % of a segment's bytes) are not uploaded but finished on the device from the
    # permutation inputs / operations (zkgpu_keccak_generate_trace, zkgpu_logic_generate_trace), inside the timed region
    T_KECCAK, T_LOGIC = 3, 5
    fin_rng = np.random.default_rng(40 + rank)
    fin_keccak = fin_logic = None
    if args.finish_on_device and not sharded:
        if mine[T_KECCAK]:
            nperm = (1 << log_ns[T_KECCAK]) // 24
            fin_keccak = (fin_rng.integers(0, 1 << 64, size=(nperm, 25), dtype=np.uint64), np.arange(1, nperm + 1, dtype=np.uint64))
        if mine[T_LOGIC]:
            fin_logic = fin_rng.integers(0, 1 << 64, size=(1 << log_ns[T_LOGIC], 9), dtype=np.uint64)
            fin_logic[:, 0] %= np.uint64(3)
    fin_h2d_bytes = h2d_bytes - sum(8 * NUM_COLUMNS[t] * (1 << log_ns[t]) for t, f in ((T_KECCAK, fin_keccak), (T_LOGIC, fin_logic)) if f is not None) \
        + (fin_keccak[0].nbytes + fin_keccak[1].nbytes if fin_keccak is not None else 0) + (fin_logic.nbytes if fin_logic is not None else 0)

    def one_segment(cx, host):
        if host == 2:
            tr, made = list(host_traces), []
            if fin_keccak is not None:
                tr[T_KECCAK] = zk.keccak_generate_trace(cx, fin_keccak[0], fin_keccak[1], 1 << log_ns[T_KECCAK])
                made.append(tr[T_KECCAK])
            if fin_logic is not None:
                tr[T_LOGIC] = zk.logic_generate_trace(cx, fin_logic, 1 << log_ns[T_LOGIC])
                made.append(tr[T_LOGIC])
            try:
                return zk.prove_with_traces(cx, tr, PUBLIC_VALUES, cfg, labels)
            finally:
                for m in made:
                    m.free()
        if host:
            return zk.prove_with_traces(cx, host_traces, PUBLIC_VALUES, cfg, labels)
        return zk.prove_with_traces(cx, None, PUBLIC_VALUES, cfg, labels, device_ptrs=ptrs)

    def step(host, single=False):
        if sharded:
            ap = zk.prove_with_traces_sharded(backend, comm, host_traces if host else ptrs, in_use, PUBLIC_VALUES, owner=owner, gather=False)
        elif host:
            ap = zk.prove_with_traces(ctx, host_traces, PUBLIC_VALUES, cfg, labels)
        else:
            ap = zk.prove_with_traces(ctx, None, PUBLIC_VALUES, cfg, labels, device_ptrs=ptrs)
        d2h[0] = sum(8 * len(p) for p in ap.stark_proofs if p is not None)
        return ap

    def prove_stream(cx, host, steps, first_barrier=None):
        """`steps` segment proofs back to back on one context.  Host traces: the uploads of segment s+1 are queued before segment s is
        proved (zkgpu_segment_upload / zkgpu_prove_segment_uploaded), so every H2D chain but the first runs under the previous proof."""
        r = None
        if host == 1 and not sharded and args.prefetch:
            nxt = zk.upload_traces(cx, host_traces, cfg)
            for k in range(steps):
                cur, nxt = nxt, (zk.upload_traces(cx, host_traces, cfg) if k + 1 < steps else None)
                r = zk.prove_with_traces(cx, None, PUBLIC_VALUES, cfg, labels, upload=cur)
        else:
            for k in range(steps):
                r = one_segment(cx, host)
        return r

    def run_steps(host, steps, single=False):
        """`steps` segment proofs on every stream.  With several streams each one is a host thread that proves its segments back to
        back (a stream of segments, as a prover node would see it); stream i starts i * stagger later so that the latency-bound phases
        of one segment (small Merkle levels, transcript round trips) fall under the throughput-bound phases of another.  With host
        traces the library's upload gate already staggers the streams (one H2D chain at a time)."""
        if sharded:
            for _ in range(steps):
                step(host)
            return
        if nstreams == 1 or single:
            r = prove_stream(ctx, host, steps)
            if r is not None:
                d2h[0] = sum(8 * len(p) for p in r.stark_proofs if p is not None)
            return
        errs = []

        def worker(i):
            try:
                if not host and args.stagger_ms > 0:
                    time.sleep(i * args.stagger_ms / 1e3)
                r = prove_stream(ctxs[i], host, steps)
                if i == 0 and r is not None:
                    d2h[0] = nstreams * sum(8 * len(p) for p in r.stark_proofs if p is not None)
            except Exception as e:      # noqa: BLE001
                errs.append(e)
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(nstreams)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errs:
            raise RuntimeError("a segment stream failed: %r" % (errs[0],))

    def barrier():
        torch.cuda.synchronize()
        for cx in ctxs:
            cx.sync()
        if dist is not None:
            dist.barrier()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(host, steps, single=False):
        barrier()
        multi = nstreams > 1 and not single
        if multi:
            # several library streams: bracket with events on torch's stream, made to wait for / be waited on by a full device sync
            ev0.record()
            torch.cuda.synchronize()
        else:
            ctx.timer_start()
        t0 = time.perf_counter()
        run_steps(host, steps, single)
        if multi:
            for cx in ctxs:
                cx.sync()
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
        else:
            ms = ctx.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        if dist is not None:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall

    run_steps(False, args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sum(cx.stats()["kernel_launches"] for cx in ctxs)
    ms, wall = timed(False, args.steps)
    launches = sum(cx.stats()["kernel_launches"] for cx in ctxs) - l0
    # kernel-family breakdown: a second timed region of the same K steps on ONE stream with the library's CUDA-event brackets on
    # (with several segments in flight the brackets of one stream would also count the other streams' kernels)
    zk._lib.check(zk.lib().zkgpu_ctx_set_profiling(ctx._h, 0 if args.no_kernel_events else 1))
    ms1, wall1 = (ms, wall) if False else timed(False, args.steps, single=True)
    kst = kernel_stats(zk, ctx)
    zk._lib.check(zk.lib().zkgpu_ctx_set_profiling(ctx._h, 0))
    run_steps(True, min(args.warmup, 1))
    e_ms, e_wall = timed(True, args.steps)
    f_ms = f_wall = None
    if args.finish_on_device and not sharded:
        run_steps(2, min(args.warmup, 1))
        f_ms, f_wall = timed(2, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        segs = args.steps * (1 if sharded else world) * nstreams          # segment proofs completed by all ranks
        tot_ms = sum(v["ms"] for v in kst.values()) or 1.0
        top = max(kst, key=lambda k: kst[k]["ms"])

        def fam(k):
            v = kst[k]
            per = v["ms"] / max(v["launches"], 1)
            ach = v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 else 0.0
            return {"launches_per_step": v["launches"] / args.steps, "ms_per_step": v["ms"] / args.steps, "share_of_kernel_time": v["ms"] / tot_ms,
                    "avg_launch_ms": per, "algorithmic_GB_per_step": v["bytes"] / args.steps / 1e9, "achieved_GBps": ach, "frac_of_peak": ach / peak}
        traffic = None
        if top == "leaf_hash" and kst[top]["launches"]:
            # per launch, like `achieved`: measured DRAM bytes per algorithmic byte (ncu capture above) x algorithmic bytes per launch
            traffic = LEAF_HASH_TRAFFIC_PER_ALGORITHMIC_BYTE * kst[top]["bytes"] / kst[top]["launches"]
        roof = {"bound": "hbm", "kernel": top, "achieved": fam(top)["achieved_GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": fam(top)["frac_of_peak"], "traffic": traffic,
                "traffic_source": "profiles/r1j_ncu_leaf_hash.raw.csv (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                  "scaled to the average launch of the timed region); bytes" if traffic is not None else None,
                "how": "CUDA events recorded by the library on its launching stream around every launch group inside a timed region of the same "
                       "K steps run with ONE segment in flight (%.1f ms/step, %.3f proofs/s); achieved = algorithmic bytes of the group "
                       "(DESIGN.md) / event time, summed over the steps" % (ms1 / args.steps, (1 if sharded else world) * args.steps / (ms1 / 1e3)),
                "families": {k: fam(k) for k in FAMILIES}}
        if top == "leaf_hash":
            perms = sum(2 * (1 << log_ns[t]) * ((NUM_COLUMNS[t] + 7) // 8) for t in range(9) if mine[t] and NUM_COLUMNS[t] > 4)
            pps = perms * args.steps / max(kst["leaf_hash"]["ms"], 1e-9) * 1e3
            roof["note"] = "Poseidon is integer-issue bound (~2.6e4 instr per 64-byte absorb): HBM fraction reported because the metric asks " \
                           "for it; trace-leaf permutations alone: %.0f Mperm/s" % (pps / 1e6)
            # the roofline that actually bounds this kernel: warp-instruction issue (1 per clock per SM sub-partition).  26.2 k
            # instructions per permutation is the ncu count of the committed capture (smsp__inst_executed.sum / permutations).
            sm_mhz = (sampler.summary().get("sm_mhz") or 1965)
            peak_issue = 148 * 4 * 32 * sm_mhz * 1e6
            roof["issue"] = {"bound": "integer issue slots", "achieved": pps * 26.2e3 / 1e12, "peak": peak_issue / 1e12, "unit": "T thread-instr/s",
                             "frac": pps * 26.2e3 / peak_issue, "source": "profiles/r1j_ncu_leaf_hash.raw.csv: 26.2 k instr / permutation, ALU pipe 82 %, FMA-heavy pipe 75 % busy"}
        cpu = None if args.no_cpu_baseline else cpu_baseline(args, log_ns)
        line = {"metric": METRIC, "value": segs / (ms / 1e3), "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": ("segment proof (AllStark, 9 tables, heights of witness_b19807080's CI ranges): " if args.workload == "segment"
                                        else "single-table prove (BASELINE config #2): ") + describe(log_ns) + "; standard_fast_config",
                           "parallelism": ("tables of one segment sharded over %d GPUs (owner %s)" % (world, owner)) if sharded
                           else ("%d independent segment(s) in flight, %d per GPU (one context + CUDA stream + host thread each, each proving its segments back to back%s)"
                                 % (world * nstreams, nstreams, ", stream i started %g ms after stream i-1 inside the timed region" % args.stagger_ms if nstreams > 1 else "")),
                           "l2": "inputs larger than L2 (%.2f GB of trace per segment)" % (sum(8 * NUM_COLUMNS[t] * (1 << log_ns[t]) for t in range(9) if in_use[t]) / 1e9),
                           "timing": "CUDA events on the library stream, max over ranks"},
                "wall_ms_per_step": wall / args.steps,
                "e2e": {"value": segs / (e_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": h2d_bytes * nstreams, "d2h_bytes_per_step": d2h[0],
                        "ms_per_step": e_ms / args.steps, "wall_ms_per_step": e_wall / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}
        if f_ms is not None:
            line["e2e_finish_on_device"] = {"value": segs / (f_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": fin_h2d_bytes * nstreams,
                                            "ms_per_step": f_ms / args.steps, "wall_ms_per_step": f_wall / args.steps,
                                            "what": "as e2e, but the Keccak and Logic traces are finished on the device from permutation inputs / operations "
                                                    "(generation inside the timed region) instead of being uploaded"}
    for cx in ctxs:
        cx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


# ---- CPU side: the oracle (C++17 + OpenMP restatement of the plonky2 / starky prover) ----------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_segment_time(log_ns, threads=None):
    from tests import oracle_lib
    orc = oracle_lib.load()
    # all the host threads the process may use — torchrun exports OMP_NUM_THREADS=1 to its workers, which is not what "the box's host
    # cores" means for the CPU arm
    orc.lib.orc_set_num_threads(int(threads or host_threads()))
    rng = np.random.default_rng(4)
    traces = [None if lg is None else rng.integers(0, 2 ** 63 - 1, size=(NUM_COLUMNS[t], 1 << lg), dtype=np.uint64) for t, lg in enumerate(log_ns)]
    t0 = time.perf_counter()
    oracle_lib.orc_prove_segment(orc, STANDARD_FAST, traces, PUBLIC_VALUES, labels=LABELS)
    return time.perf_counter() - t0, orc.lib.orc_num_threads()


def cpu_sample_shape(args, log_ns, extra=0):
    sh = args.cpu_shrink + extra
    return [None if lg is None else max(4, lg - sh) for lg in log_ns], float(1 << sh)


def cpu_full_size_time(args, log_ns):
    """Bounded sample of the CPU prover, extrapolated to the full segment.  A proof has costs that grow with the rows (NTTs, hashing,
    constraint evaluation) and costs that do not (proof-of-work grind, 84 query rounds, transcript): the segment is proved with every
    table 2^s and 2^(s+1) times shorter, T(k) = F + R 2^-k is solved for F and R, and T(0) = F + R is reported (scaling one sample
    by 2^s would multiply the fixed part too; the n log n terms make the linear model slightly favourable to the CPU)."""
    s = args.cpu_shrink
    t_s, threads = oracle_segment_time(cpu_sample_shape(args, log_ns)[0])
    t_s1, _ = oracle_segment_time(cpu_sample_shape(args, log_ns, 1)[0])
    per_row = max(t_s - t_s1, 0.0) * 2.0 * (1 << s)            # R
    fixed = max(2.0 * t_s1 - t_s, 0.0)                          # F
    return fixed + per_row, threads, ("oracle prove_segment with every table 2^%d x and 2^%d x shorter took %.2f s and %.2f s on %d threads; "
                                      "fixed part %.2f s + row-proportional part %.2f s at full size" % (s, s + 1, t_s, t_s1, threads, fixed, per_row))


def cpu_baseline(args, log_ns):
    secs, threads, how = cpu_full_size_time(args, log_ns)
    return {"value": 1.0 / secs, "unit": "proofs/s", "cores": int(threads), "kind": "port", "sample": how}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    log_ns = segment_shape(args)
    sample, _ = cpu_sample_shape(args, log_ns)
    for _ in range(min(args.warmup, 1)):
        oracle_segment_time([None if lg is None else max(4, lg - 4) for lg in sample])
    t, threads, how = 0.0, 1, ""
    for _ in range(args.steps):
        s, threads, how = cpu_full_size_time(args, log_ns)
        t += s
    per_step = t / args.steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    val = 1.0 / per_step
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": ("segment proof (AllStark, 9 tables): " if args.workload == "segment" else "single-table prove: ") + describe(log_ns)
                               + "; standard_fast_config; CPU restatement (oracle/, C++17 + OpenMP) of the plonky2/starky prover — the Rust "
                                 "reference cannot be built here (no cargo/rustc, crates not vendored)"},
        "cpu_baseline": {"value": val, "unit": "proofs/s", "cores": int(threads), "kind": "port",
                         "sample": "each step: " + how},
        "e2e": {"value": val, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="segment", choices=["segment", "cpu_table"])
    ap.add_argument("--parallelism", default="segments", choices=["segments", "tables"])
    ap.add_argument("--log-n", type=int, default=20, help="cpu_table workload: log2 of the trace length (BASELINE config #2: 20)")
    ap.add_argument("--shrink", type=int, default=0, help="segment workload: make every table 2^shrink times shorter (smoke runs)")
    ap.add_argument("--cpu-shrink", type=int, default=5, help="CPU legs prove tables 2^k times shorter and scale the time")
    ap.add_argument("--streams", type=int, default=3, help="segments in flight per GPU (parallelism=segments); measured 1: 3.34, 2: 3.78, 3: 3.95, 4: 3.86 proofs/s")
    ap.add_argument("--prefetch", type=int, default=0,
                    help="e2e: queue the uploads of segment s+1 before proving segment s (1) or upload inside each prove call (0); measured "
                         "(profiles/r1w): upload-ahead is slower on this box (3 streams 2.94 vs 3.87 proofs/s), the per-call upload chain "
                         "already hides under the other segments' kernels")
    ap.add_argument("--stagger-ms", type=float, default=80.0, help="start offset between the segment streams of a GPU (inside the timed region)")
    ap.add_argument("--finish-on-device", action="store_true",
                    help="extra e2e leg: Keccak / Logic traces finished on the device from their inputs instead of uploaded (e2e_finish_on_device)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiler runs)")
    ap.add_argument("--no-kernel-events", action="store_true", help="do not bracket kernel families with CUDA events")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints meanwhile (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    out = buf.getvalue().strip()
    if out:
        print(out.splitlines()[-1], flush=True)


if __name__ == "__main__":
    main()
