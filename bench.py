#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 STARK proving path (contract: task statement / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the zkgpu C ABI)
  python bench.py --impl reference --gpus N ...            # reference arm: the CPU restatement (oracle) on all host cores

Metric (BASELINE.json): segment proofs/sec.  A step = one segment proof PER SEGMENT STREAM (--streams S, default 4: S segments in
flight per GPU, each on its own context + CUDA stream + host thread, proving its segments back to back); a segment proof =
prove_with_traces over all nine STARK tables of a synthetic segment with the table heights of the `witness_b19807080` CI ranges
(SURVEY.md 8d config #4): per table the trace commitment, CTL / lookup auxiliary columns + commitment, fused quotient evaluation +
commitment, openings, FRI commit phase, proof of work and query answers, all tables chained through one Fiat-Shamir transcript.
  value : proofs/s with the traces already resident in HBM, device time (CUDA events on the library's stream), max over ranks
  e2e   : the same call with the traces in pinned host memory (H2D inside the timed region) and the proofs read back (D2H)
N > 1 (one process per GPU, torchrun):
  --parallelism segments (default)  every rank proves its own segments — segments are independent proofs (own transcript), the
                                    way the reference spreads them over workers; no data-path collective; "weak" scaling
  --parallelism tables              the north-star layout: the nine tables of ONE segment spread over the ranks (DESIGN.md
                                    "Multi-GPU"); "strong" scaling.  The default run at N > 1 also measures this layout and reports it
                                    as the `table_sharded` sub-record of its line.
Sub-records of the default line (the other BASELINE configs, so that they are in the driver's record; --no-extras skips them):
  config2_cpu_table   single CpuStark prove, 2^20 rows (BASELINE config #2)
  config3_b3_b6       the small segment of witness_b3_b6's CI ranges, standard_fast_config and TEST_STARK_CONFIG (config #3)
  config5_stream      a stream of 64 segments with heights drawn from the generic CI ranges, spread over the GPUs (config #5)
  table_sharded       N > 1: one segment's tables over the N GPUs, ms per proof and speed-up over one GPU (config #4's layout)
"""
import argparse
import contextlib
import ctypes as C
import io
import json
import os
import queue
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_CPU, T_KECCAK, T_LOGIC = 2, 3, 5
TABLE_NAMES = ("Arithmetic", "BytePacking", "Cpu", "Keccak", "KeccakSponge", "Logic", "Memory", "MemBefore", "MemAfter")
NUM_COLUMNS = (116, 71, 85, 2431, 438, 523, 30, 12, 12)
SEGMENT_CONFIGS = {
    # top of the CI height ranges of artifacts/witness_b19807080.json (scripts/prove_stdio.rs:89-101)
    "b19807080": (17, 14, 19, 17, 13, 16, 21, 19, 19),
    # a small segment inside the CI ranges of artifacts/witness_b3_b6.json (scripts/prove_stdio.rs:102-114)
    "b3_b6": (16, 10, 16, 12, 8, 10, 18, 16, 7),
}
SEGMENT_LOG_NS = SEGMENT_CONFIGS["b19807080"]
# the generic CI ranges (scripts/prove_stdio.rs:115-127), Rust half-open lo..hi: config #5 draws its heights from these
GENERIC_RANGES = ((16, 18), (8, 15), (9, 20), (7, 18), (8, 14), (5, 17), (17, 22), (16, 20), (7, 20))
STANDARD_FAST = (100, 2, 1, 4, 16, 4, 5, 84)
TEST_CONFIG = (1, 1, 1, 4, 1, 4, 5, 1)              # TEST_STARK_CONFIG, evm_arithmetization/src/testing_utils.rs:41-52 (ci.yml:196)
STARK_CONFIGS = {"standard_fast": STANDARD_FAST, "test": TEST_CONFIG}
LABELS = (0x1234, 0x77, 0x4000, 0x5000)
PUBLIC_VALUES = np.arange(1, 2218, dtype=np.uint64)       # 2217 observed elements (flatten_public_values, get_challenges.rs:202-227), synthetic
METRIC = "segment proofs/sec"
# DRAM bytes per algorithmic byte of the dominant kernel (leaf_hash), from the committed `ncu --set full` capture of its largest launch with
# the round-2 permutation, profiles/r2r_ncu_leaf_hash_keccak.raw.csv.gz: Keccak trace leaves (2^18 rows x 2431 columns), 71.0 ms,
# dram__bytes_read.sum + dram__bytes_write.sum = 5.1152 GB + 0.0150 GB for (8 * 2431 + 32) * 2^18 = 5.1066 GB algorithmic (every LDE
# column read once, one digest written per row)
LEAF_HASH_TRAFFIC_PER_ALGORITHMIC_BYTE = (5.115167 + 0.014989) / 5.10657
FAMILIES = ("leaf_hash", "merkle_levels", "ntt", "quotient", "aux_columns", "openings", "fri", "pow")
# thread-instructions per Poseidon permutation of the production leaf-hash kernel: ncu smsp__inst_executed.sum x 32 / permutations —
# Keccak: 63.78e9 warp-instructions for 2^18 leaves x 304 permutations (profiles/r2r_ncu_leaf_hash_keccak.raw.csv.gz), the same figure
# for Memory (2^22 x 4) and Cpu (2^20 x 11) in profiles/r2p_ncu_leaf_merkle_summary.csv
INSTR_PER_PERMUTATION = 25.6e3


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
    return [max(4, lg - args.shrink) for lg in SEGMENT_CONFIGS[args.config]]


def describe(log_ns):
    return ", ".join("%s 2^%d x %d" % (TABLE_NAMES[t], lg, NUM_COLUMNS[t]) for t, lg in enumerate(log_ns) if lg is not None)


def trace_bytes(log_ns, which=None):
    return sum(8 * NUM_COLUMNS[t] * (1 << lg) for t, lg in enumerate(log_ns) if lg is not None and (which is None or which[t]))


def config_record(args, log_ns):
    """names the workload; built from the command line only, so both arms print the same object"""
    if args.workload == "cpu_table":
        what = "single-table prove (BASELINE config #2): "
    else:
        what = "segment proof (AllStark, 9 tables, heights of witness_%s's CI ranges): " % args.config
    return {"workload": what + describe(log_ns) + "; " + args.stark_config + ("_config" if args.stark_config == "standard_fast" else " (TEST_STARK_CONFIG)"),
            "l2": "inputs larger than L2 (%.2f GB of trace per segment)" % (trace_bytes(log_ns) / 1e9),
            "timing": "our arm: CUDA events on the library stream, max over ranks; reference arm: steady_clock around the prover call"}


def kernel_stats(zk, ctx):
    out = {}
    for i, name in enumerate(FAMILIES):
        l, ms, b = C.c_uint64(), C.c_double(), C.c_double()
        zk._lib.check(zk.lib().zkgpu_ctx_kernel_stats(ctx._h, C.c_uint32(i), C.byref(l), C.byref(ms), C.byref(b)))
        out[name] = {"launches": l.value, "ms": ms.value, "bytes": b.value}
    return out


class Rig:
    """Synthetic traces of one segment shape: uniform random canonical field elements (with rate_bits = 1 the prover does the same
    work on any trace), resident in HBM and mirrored in pinned host memory for the e2e legs."""

    def __init__(self, torch, dev, log_ns, seed, mine=None, host=True, common_seed=False):
        self.log_ns = list(log_ns)
        self.in_use = [lg is not None for lg in log_ns]
        self.mine = [self.in_use[t] and (mine is None or mine[t]) for t in range(9)]
        self.dev_traces, self.host_traces = [None] * 9, [None] * 9
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        for t in range(9):
            if not self.in_use[t]:
                continue
            if common_seed:
                g.manual_seed(seed * 16 + t)     # table t is the same on whichever rank makes it
                if not self.mine[t]:
                    continue
            x = torch.randint(0, 2 ** 63 - 1, (NUM_COLUMNS[t], 1 << log_ns[t]), dtype=torch.int64, device=dev, generator=g)
            if self.mine[t]:
                self.dev_traces[t] = x
                if host:
                    h = torch.empty(x.shape, dtype=torch.int64, pin_memory=True)
                    h.copy_(x)
                    self.host_traces[t] = h.numpy().view(np.uint64)
                    self._keep = getattr(self, "_keep", []) + [h]
            del x
        torch.cuda.synchronize()
        self.ptrs = [None if d is None else (d.data_ptr(), d.shape[1]) for d in self.dev_traces]
        self.h2d_bytes = trace_bytes(log_ns, self.mine)

    def free(self):
        self.dev_traces = self.host_traces = self.ptrs = None
        self._keep = []


class Runner:
    """steps of `prove_with_traces` on the contexts of one GPU, timed on the device"""

    def __init__(self, torch, dist, zk, ctxs, dev, args):
        self.torch, self.dist, self.zk, self.ctxs, self.dev, self.args = torch, dist, zk, ctxs, dev, args
        self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.d2h = 0

    def barrier(self):
        self.torch.cuda.synchronize()
        for cx in self.ctxs:
            cx.sync()
        if self.dist is not None:
            self.dist.barrier()

    def run_streams(self, fn, steps, nstreams, stagger_ms=0.0):
        """fn(ctx) proves one segment; every stream proves `steps` segments back to back"""
        if nstreams == 1:
            r = None
            for _ in range(steps):
                r = fn(self.ctxs[0])
            return r
        errs, res = [], [None]

        def worker(i):
            try:
                if stagger_ms > 0:
                    time.sleep(i * stagger_ms / 1e3)
                r = None
                for _ in range(steps):
                    r = fn(self.ctxs[i])
                if i == 0:
                    res[0] = r
            except Exception as e:      # noqa: BLE001
                errs.append(e)
        ths = [threading.Thread(target=worker, args=(i,)) for i in range(nstreams)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errs:
            raise RuntimeError("a segment stream failed: %r" % (errs[0],))
        return res[0]

    def timed(self, body, single_ctx):
        """device time of body(): the library's own events on its stream when one context is in use, else events on torch's stream
        bracketing full device syncs; max over ranks"""
        self.barrier()
        if single_ctx:
            self.ctxs[0].timer_start()
        else:
            self.ev0.record()
            self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        body()
        if single_ctx:
            ms = self.ctxs[0].timer_stop()
        else:
            for cx in self.ctxs:
                cx.sync()
            self.ev1.record()
            self.torch.cuda.synchronize()
            ms = self.ev0.elapsed_time(self.ev1)
        wall = (time.perf_counter() - t0) * 1e3
        self.barrier()
        if self.dist is not None:
            t = self.torch.tensor([ms, wall], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall


def proof_bytes(ap):
    return sum(8 * len(p) for p in ap.stark_proofs if p is not None)


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
    sharded = args.parallelism == "tables" and world > 1
    # --streams S: S segments in flight per GPU, each on its own context (own CUDA stream) driven by its own host thread (ctypes
    # releases the GIL): the latency-bound tails (small Merkle levels, transcript round trips) of one overlap the kernels of another
    nstreams = 1 if sharded else max(1, args.streams)
    ctxs = [zk.Context(local_rank) for _ in range(nstreams)]
    ctx = ctxs[0]
    cfg = zk.StarkConfig(*STARK_CONFIGS[args.stark_config])
    labels = zk.KernelLabels(*LABELS)
    log_ns = segment_shape(args)
    run = Runner(torch, dist, zk, ctxs, dev, args)
    comm = zk.TorchComm(device=dev) if world > 1 else None

    def sharded_setup(shape):
        plan = zk.shard_plan(world, shape) if hasattr(zk, "shard_plan") else None
        owner = plan.owner if plan is not None else zk.default_owner(world, weights=[
            (0 if lg is None else (1 << lg) * NUM_COLUMNS[t] * ((NUM_COLUMNS[t] + 7) // 8 + 6)) for t, lg in enumerate(shape)])
        need = plan.needs(rank) if plan is not None else [shape[t] is not None and owner[t] == rank for t in range(9)]
        r = Rig(torch, dev, shape, seed=4, mine=need, common_seed=True)
        be = zk.ZkGpuBackend(ctx, cfg, labels, precompute_constraints=True)
        return plan, owner, r, be

    def sharded_step(plan, owner, r, be, host):
        tr = r.host_traces if host else r.ptrs
        if plan is not None:
            return zk.prove_with_traces_sharded(be, comm, tr, r.in_use, PUBLIC_VALUES, plan=plan, gather=False)
        return zk.prove_with_traces_sharded(be, comm, tr, r.in_use, PUBLIC_VALUES, owner=owner, gather=False)

    plan = owner = backend = None
    if sharded:
        plan, owner, rig, backend = sharded_setup(log_ns)
    else:
        rig = Rig(torch, dev, log_ns, seed=4 + rank)

    # --finish-on-device leg: the Keccak and Logic traces (72 % of a segment's bytes) are not uploaded but finished on the device from
    # the permutation inputs / operations (zkgpu_keccak_generate_trace, zkgpu_logic_generate_trace), inside the timed region
    fin_rng = np.random.default_rng(40 + rank)
    fin_keccak = fin_logic = None
    if not sharded and args.workload == "segment":
        nperm = (1 << log_ns[T_KECCAK]) // 24
        if nperm:
            fin_keccak = (fin_rng.integers(0, 1 << 64, size=(nperm, 25), dtype=np.uint64), np.arange(1, nperm + 1, dtype=np.uint64))
        fin_logic = fin_rng.integers(0, 1 << 64, size=(1 << log_ns[T_LOGIC], 9), dtype=np.uint64)
        fin_logic[:, 0] %= np.uint64(3)
    fin_h2d_bytes = rig.h2d_bytes - sum(8 * NUM_COLUMNS[t] * (1 << log_ns[t]) for t, f in ((T_KECCAK, fin_keccak), (T_LOGIC, fin_logic)) if f is not None) \
        + (fin_keccak[0].nbytes + fin_keccak[1].nbytes if fin_keccak is not None else 0) + (fin_logic.nbytes if fin_logic is not None else 0)

    BG = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
    STATE0 = np.arange(1, 13, dtype=np.uint64)

    class TableProofs:      # what one_table returns: looks like an AllProof to proof_bytes()
        def __init__(self, words):
            self.stark_proofs = [words]

    def one_table(cx, host, r, c):
        """BASELINE config #2: PolynomialBatch::from_values + get_ctl_data + prove_single_table of the one table of the rig
        (prover.rs:100-107, 137-143, 301-341), CTL challenges and transcript state fixed"""
        t = r.in_use.index(True)
        if host:
            tb = zk.PolynomialBatch.from_values(cx, r.host_traces[t], c.rate_bits, c.cap_height, keep_values=True)
        else:
            tb = zk.PolynomialBatch.from_device_values(cx, r.ptrs[t][0], NUM_COLUMNS[t], r.ptrs[t][1], c.rate_bits, c.cap_height, keep_values=True)
        ctl = zk.get_ctl_data(cx, t, tb, BG[:2 * c.num_challenges], c.num_challenges)
        proof, _ = zk.prove_single_table(cx, t, c, tb, ctl, STATE0, labels=labels)
        words = proof.words
        proof.free(); ctl.free(); tb.free()
        return TableProofs(words)

    def one_segment(cx, host, r=None, c=None):
        r, c = r or rig, c or cfg
        if sum(r.in_use) == 1:
            return one_table(cx, 1 if host else 0, r, c)
        if host == 2:
            tr, made = list(r.host_traces), []
            if fin_keccak is not None:
                tr[T_KECCAK] = zk.keccak_generate_trace(cx, fin_keccak[0], fin_keccak[1], 1 << log_ns[T_KECCAK])
                made.append(tr[T_KECCAK])
            if fin_logic is not None:
                tr[T_LOGIC] = zk.logic_generate_trace(cx, fin_logic, 1 << log_ns[T_LOGIC])
                made.append(tr[T_LOGIC])
            try:
                return zk.prove_with_traces(cx, tr, PUBLIC_VALUES, c, labels)
            finally:
                for m in made:
                    m.free()
        if host:
            return zk.prove_with_traces(cx, r.host_traces, PUBLIC_VALUES, c, labels)
        return zk.prove_with_traces(cx, None, PUBLIC_VALUES, c, labels, device_ptrs=r.ptrs)

    def steps_of(host, steps, single=False, r=None, c=None):
        """-> (device ms, wall ms) of `steps` segment proofs on every stream (one stream when `single`)"""
        def body():
            if sharded:
                ap = None
                for _ in range(steps):
                    ap = sharded_step(plan, owner, rig, backend, host)
                run.d2h = proof_bytes(ap)
                return
            ns = 1 if single else nstreams
            ap = run.run_streams(lambda cx: one_segment(cx, host, r, c), steps, ns, 0.0 if host else args.stagger_ms)
            if ap is not None:
                run.d2h = ns * proof_bytes(ap)
        return run.timed(body, single_ctx=(sharded or single or nstreams == 1))

    steps_of(False, args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sum(cx.stats()["kernel_launches"] for cx in ctxs)
    ms, wall = steps_of(False, args.steps)
    launches = sum(cx.stats()["kernel_launches"] for cx in ctxs) - l0
    # kernel-family breakdown: a second timed region of the same K steps on ONE stream with the library's CUDA-event brackets on
    # (with several segments in flight the brackets of one stream would also count the other streams' kernels)
    zk._lib.check(zk.lib().zkgpu_ctx_set_profiling(ctx._h, 0 if args.no_kernel_events else 1))
    ms1, wall1 = steps_of(False, args.steps, single=True)
    kst = kernel_stats(zk, ctx)
    zk._lib.check(zk.lib().zkgpu_ctx_set_profiling(ctx._h, 0))
    steps_of(True, min(args.warmup, 1))
    e_ms, e_wall = steps_of(True, args.steps)
    d2h_bytes = run.d2h
    f_ms = f_wall = None
    if fin_keccak is not None and not args.no_extras:
        steps_of(2, min(args.warmup, 1))
        f_ms, f_wall = steps_of(2, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    single_latency_ms = ms1 / args.steps

    # ---- the other BASELINE configs as sub-records of the same line -------------------------------------------------------------
    extras = {}
    if not args.no_extras and not sharded and args.workload == "segment" and args.shrink == 0:
        xsteps = max(2, min(args.steps, 5))
        rig.free()
        torch.cuda.empty_cache()
        # config #2: single CpuStark table, 2^20 rows
        shape2 = [20 if t == T_CPU else None for t in range(9)]
        r2 = Rig(torch, dev, shape2, seed=2 + rank)
        steps_of(False, 2, r=r2)
        m2, _ = steps_of(False, xsteps, r=r2)
        m2s, _ = steps_of(False, xsteps, single=True, r=r2)
        e2, _ = steps_of(True, xsteps, r=r2)
        extras["config2_cpu_table"] = {
            "workload": "single CpuStark prove, 2^20 rows x 85 columns, standard_fast_config (BASELINE config #2), uniform random trace",
            "value": world * nstreams * xsteps / (m2 / 1e3), "unit": "table proofs/s", "e2e": world * nstreams * xsteps / (e2 / 1e3),
            "ms_per_proof_one_in_flight": m2s / xsteps, "steps": xsteps}
        r2.free()
        # config #3: the b3_b6 segment under both StarkConfigs
        r3 = Rig(torch, dev, SEGMENT_CONFIGS["b3_b6"], seed=3 + rank)
        rec3 = {"workload": "segment proof: " + describe(SEGMENT_CONFIGS["b3_b6"]) + " (BASELINE config #3, CI height ranges of witness_b3_b6, "
                            "scripts/prove_stdio.rs:102-114)"}
        for name, words in STARK_CONFIGS.items():
            c3 = zk.StarkConfig(*words)
            steps_of(False, 2, r=r3, c=c3)
            m3, _ = steps_of(False, xsteps, r=r3, c=c3)
            m3s, _ = steps_of(False, xsteps, single=True, r=r3, c=c3)
            e3, _ = steps_of(True, xsteps, r=r3, c=c3)
            rec3[name] = {"value": world * nstreams * xsteps / (m3 / 1e3), "unit": "proofs/s", "e2e": world * nstreams * xsteps / (e3 / 1e3),
                          "ms_per_proof_one_in_flight": m3s / xsteps, "steps": xsteps}
        extras["config3_b3_b6"] = rec3
        r3.free()
        # config #5: a stream of 64 segments of varied heights over all GPUs
        extras["config5_stream"] = stream_workload(torch, dist, zk, run, ctxs, dev, cfg, labels, rank, world, nstreams)
        torch.cuda.empty_cache()
        # config #4's layout: the tables of one segment over the N GPUs
        if world > 1:
            for cx in ctxs[1:]:
                cx.sync()
            sp, so, sr, sb = sharded_setup(log_ns)

            def sh_body(host, n):
                for _ in range(n):
                    sharded_step(sp, so, sr, sb, host)
            run.timed(lambda: sh_body(False, 2), True)
            ts_ms, _ = run.timed(lambda: sh_body(False, xsteps), True)
            run.timed(lambda: sh_body(True, 1), True)
            te_ms, _ = run.timed(lambda: sh_body(True, xsteps), True)
            lat = torch.tensor([single_latency_ms], dtype=torch.float64, device=dev)
            dist.broadcast(lat, src=0)
            extras["table_sharded"] = {
                "what": "the nine tables of ONE segment over the %d GPUs (north-star layout): ms per segment proof, traces resident in HBM; "
                        "speed-up over one segment in flight on one GPU (rank 0's ms_per_proof_one_in_flight)" % world,
                "ms_per_proof": ts_ms / xsteps, "proofs_per_s": xsteps / (ts_ms / 1e3), "e2e_ms_per_proof": te_ms / xsteps,
                "one_gpu_ms_per_proof": float(lat[0]), "speedup": float(lat[0]) / (ts_ms / xsteps), "steps": xsteps,
                "plan": sp.describe() if sp is not None else {"owner": list(so)},
                "phases_ms": getattr(sb, "phase_ms", None)}
            sr.free()

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
                "traffic_source": "profiles/r2r_ncu_leaf_hash_keccak.raw.csv.gz (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                  "scaled to the average launch of the timed region); bytes" if traffic is not None else None,
                "how": "CUDA events recorded by the library on its launching stream around every launch group inside a timed region of the same "
                       "K steps run with ONE segment in flight (%.1f ms/step, %.3f proofs/s); achieved = algorithmic bytes of the group "
                       "(DESIGN.md) / event time, summed over the steps" % (ms1 / args.steps, (1 if sharded else world) * args.steps / (ms1 / 1e3)),
                "families": {k: fam(k) for k in FAMILIES}}
        if top == "leaf_hash":
            perms = sum(2 * (1 << log_ns[t]) * ((NUM_COLUMNS[t] + 7) // 8) for t in range(9) if rig.mine[t] and NUM_COLUMNS[t] > 4)
            pps = perms * args.steps / max(kst["leaf_hash"]["ms"], 1e-9) * 1e3
            roof["note"] = "Poseidon is integer-issue bound (~2.6e4 instr per 64-byte absorb): HBM fraction reported because the metric asks " \
                           "for it; trace-leaf permutations alone: %.0f Mperm/s" % (pps / 1e6)
            # the roofline that actually bounds this kernel: warp-instruction issue (1 per clock per SM sub-partition)
            sm_mhz = (clocks.get("sm_mhz") or 1965)
            peak_issue = 148 * 4 * 32 * sm_mhz * 1e6
            roof["issue"] = {"bound": "integer issue slots", "achieved": pps * INSTR_PER_PERMUTATION / 1e12, "peak": peak_issue / 1e12, "unit": "T thread-instr/s",
                             "frac": pps * INSTR_PER_PERMUTATION / peak_issue,
                             "source": "profiles/r2r_ncu_leaf_hash_keccak.raw.csv.gz: %.1f k thread-instructions per permutation (ncu smsp__inst_executed.sum x 32 / permutations), ALU pipe 81 %% busy" % (INSTR_PER_PERMUTATION / 1e3)}
        # the CPU baseline is timed at N = 1 only (at N > 1 the other ranks' host threads would share the cores with it)
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(args, log_ns)
        line = {"metric": METRIC, "value": segs / (ms / 1e3), "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config_record(args, log_ns),
                "schedule": ("tables of one segment sharded over %d GPUs: %s" % (world, plan.describe() if plan is not None else owner)) if sharded
                else ("%d independent segment(s) in flight, %d per GPU (one context + CUDA stream + host thread each, each proving its segments back to back%s)"
                      % (world * nstreams, nstreams, ", stream i started %g ms after stream i-1 inside the timed region" % args.stagger_ms if nstreams > 1 else "")),
                "wall_ms_per_step": wall / args.steps,
                "single_segment_latency_ms": single_latency_ms,
                "e2e": {"value": segs / (e_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": rig.h2d_bytes * nstreams, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": e_ms / args.steps, "wall_ms_per_step": e_wall / args.steps,
                        "mode": "host traces: all nine traces in pinned host memory, uploaded inside the timed region, proofs read back"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        line["e2e_host_traces"] = dict(line["e2e"])
        if f_ms is not None:
            line["e2e_finish_on_device"] = {"value": segs / (f_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": fin_h2d_bytes * nstreams,
                                            "d2h_bytes_per_step": d2h_bytes, "ms_per_step": f_ms / args.steps, "wall_ms_per_step": f_wall / args.steps,
                                            "mode": "finish on device: the Keccak and Logic traces (72 % of a segment's bytes) are finished on the device from "
                                                    "host permutation inputs / operations inside the timed region (zkgpu_keccak_generate_trace, "
                                                    "zkgpu_logic_generate_trace: more device work, fewer PCIe bytes), the other seven traces uploaded, proofs read back"}
            # the end-to-end number is the better of the two host-facing ways to call the path (both always reported): at one GPU the
            # uploads hide under the commitments either way, at eight the shared PCIe / host memory path does not keep up with full uploads
            if line["e2e_finish_on_device"]["value"] > line["e2e"]["value"]:
                line["e2e"] = dict(line["e2e_finish_on_device"])
        line.update(extras)
    for cx in ctxs:
        cx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def stream_heights(count=64, seed=5):
    rng = np.random.default_rng(seed)
    return [tuple(int(rng.integers(lo, hi)) for lo, hi in GENERIC_RANGES) for _ in range(count)]


def stream_workload(torch, dist, zk, run, ctxs, dev, cfg, labels, rank, world, nstreams, count=64):
    """BASELINE config #5: `count` segments with heights drawn (seed 5) from the generic CI ranges, segment s proved by rank s mod N,
    `nstreams` in flight per GPU through the product's SegmentProver (zk_evm_b200/scheduler.py), traces resident / in pinned host memory."""
    heights = stream_heights(count)
    mine = [h for s, h in enumerate(heights) if s % world == rank]
    top = [max(h[t] for h in heights) for t in range(9)]
    pool = Rig(torch, dev, top, seed=50 + rank)           # one random pool per table; a segment uses a prefix of it

    def dev_seg(h):
        return [(pool.dev_traces[t].data_ptr(), 1 << h[t]) for t in range(9)]

    def host_seg(h):
        return [pool.host_traces[t].reshape(-1)[:NUM_COLUMNS[t] << h[t]].reshape(NUM_COLUMNS[t], 1 << h[t]) for t in range(9)]

    class Worker:       # a pre-warmed context per scheduler thread (no `close`: the contexts outlive the scheduler)
        def __init__(self, cx):
            self.cx = cx
    free = queue.Queue()
    nbytes = [0]

    def make_worker(device):
        return Worker(free.get())

    def prove(state, traces, public_values, lab, abort_flag):
        if isinstance(traces[0], tuple):
            ap = zk.prove_with_traces(state.cx, None, public_values, cfg, labels, device_ptrs=traces, abort_flag=abort_flag)
        else:
            ap = zk.prove_with_traces(state.cx, traces, public_values, cfg, labels, abort_flag=abort_flag)
        nbytes[0] += proof_bytes(ap)
        return None

    def go(host):
        for cx in ctxs:
            free.put(cx)
        sp = zk.SegmentProver(device=dev.index, streams=nstreams, config=cfg, labels=labels, make_worker=make_worker, prove=prove)
        sp.prove_all(((host_seg(h) if host else dev_seg(h)), PUBLIC_VALUES) for h in mine)
    go(False)                                             # warm-up pass (memory pools, twiddle tables of every height)
    ms, _ = run.timed(lambda: go(False), single_ctx=False)
    nbytes[0] = 0
    e_ms, _ = run.timed(lambda: go(True), single_ctx=False)
    rec = {"workload": "%d segments, heights drawn (seed 5) uniformly from the generic CI ranges scripts/prove_stdio.rs:115-127 %s, "
                       "segment s on GPU s mod %d, %d in flight per GPU (SegmentProver), standard_fast_config" % (count, list(GENERIC_RANGES), world, nstreams),
           "value": count / (ms / 1e3), "unit": "proofs/s", "e2e": count / (e_ms / 1e3), "segments": count,
           "mean_trace_GB_per_segment": float(np.mean([trace_bytes(h) for h in heights])) / 1e9,
           "h2d_bytes_total_rank0": int(sum(trace_bytes(h) for h in mine)), "d2h_bytes_total_rank0": int(nbytes[0])}
    pool.free()
    return rec


# ---- CPU side: the oracle (C++17 + OpenMP restatement of the plonky2 / starky prover) ----------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_segment_time(log_ns, threads=None, stark_config=STANDARD_FAST):
    from tests import oracle_lib
    orc = oracle_lib.load()
    # all the host threads the process may use — torchrun exports OMP_NUM_THREADS=1 to its workers, which is not what "the box's host
    # cores" means for the CPU arm
    orc.lib.orc_set_num_threads(int(threads or host_threads()))
    rng = np.random.default_rng(4)
    traces = [None if lg is None else rng.integers(0, 2 ** 63 - 1, size=(NUM_COLUMNS[t], 1 << lg), dtype=np.uint64) for t, lg in enumerate(log_ns)]
    t0 = time.perf_counter()
    if sum(tr is not None for tr in traces) == 1:        # BASELINE config #2: one table through commit + ctl data + prove_single_table
        t = [tr is not None for tr in traces].index(True)
        bg = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
        oracle_lib.orc_prove_table(orc, t, stark_config, traces[t], bg[:2 * stark_config[1]], np.arange(1, 13, dtype=np.uint64))
    else:
        oracle_lib.orc_prove_segment(orc, stark_config, traces, PUBLIC_VALUES, labels=LABELS)
    return time.perf_counter() - t0, orc.lib.orc_num_threads()


def oracle_stage_report():
    from tests import oracle_lib
    orc = oracle_lib.load()
    try:
        orc.lib.orc_stage_report.restype = C.c_size_t
        buf = C.create_string_buffer(8192)
        orc.lib.orc_stage_report(buf, C.c_size_t(8192), 1)
        return {k: float(v) for k, v in (l.split("\t") for l in buf.value.decode().splitlines())}
    except Exception:
        return None


def cpu_baseline(args, log_ns):
    """In-arm CPU baseline: ONE bounded sample — the same segment with every table 2^s times shorter, proved once by the oracle on all
    host threads (about 10-30 s) — scaled by 2^s.  The row-proportional work (NTTs, hashing, constraint evaluation) dominates at these
    sizes; the reference arm (`--impl reference`) proves the full-size segment and is the number to compare with."""
    s = args.cpu_shrink
    shape = [None if lg is None else max(4, lg - s) for lg in log_ns]
    oracle_stage_report()
    t, threads = oracle_segment_time(shape, stark_config=STARK_CONFIGS[args.stark_config])
    return {"value": 1.0 / (t * (1 << s)), "unit": "proofs/s", "cores": int(threads), "kind": "port",
            "sample": "oracle prove_segment (C++17 + OpenMP restatement, oracle/) of the same segment with every table 2^%d x shorter: %.2f s on %d "
                      "threads, scaled by 2^%d to full size (n log n terms ignored: favourable to the CPU)" % (s, t, threads, s),
            "sample_seconds": t, "stages_seconds": oracle_stage_report()}


def run_reference(args):
    """The reference arm: the CPU implementation of the path on the box's host cores.  The Rust prover cannot be built here (no
    cargo / rustc, crates not vendored), so this is the oracle port — every step proves the FULL-SIZE segment of the configuration
    once, really (no extrapolation); the number of steps is capped by --reference-budget seconds and the steps actually run are
    what the line reports."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    log_ns = segment_shape(args)
    sc = STARK_CONFIGS[args.stark_config]
    warm = 0
    if args.warmup > 0:      # touches the library, the thread pool and the FFT root tables: a small segment, not a timed step
        oracle_segment_time([None if lg is None else max(4, lg - 6) for lg in log_ns], stark_config=sc)
        warm = 1
    oracle_stage_report()
    times, threads = [], 1
    t_begin = time.perf_counter()
    for k in range(args.steps):
        t, threads = oracle_segment_time(log_ns, stark_config=sc)
        times.append(t)
        spent = time.perf_counter() - t_begin
        if k + 1 < args.steps and spent + 1.15 * t > args.reference_budget:      # the next step (and making its traces) would not fit
            break
    per_step = sum(times) / len(times)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    val = 1.0 / per_step
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "proofs/s", "n_gpus": world, "steps": len(times),
        "warmup": warm, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config_record(args, log_ns),
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "steps_note": "every step is one full-size segment proof on the CPU (%s s each); steps capped by --reference-budget %g s; the warm-up is a "
                      "2^6 x shorter segment (library / thread pool / root tables), not a timed step" % (", ".join("%.1f" % t for t in times), args.reference_budget),
        "cpu_baseline": {"value": val, "unit": "proofs/s", "cores": int(threads), "kind": "port",
                         "sample": "full-size segment, oracle prove_segment (C++17 + OpenMP restatement of the plonky2 / starky prover, oracle/) on %d "
                                   "threads, %d step(s) measured; the Rust reference cannot be built here (no cargo / rustc, crates not vendored)"
                                   % (threads, len(times)),
                         "stages_seconds": oracle_stage_report()},
        "e2e": {"value": val, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="segment", choices=["segment", "cpu_table"])
    ap.add_argument("--config", default="b19807080", choices=sorted(SEGMENT_CONFIGS), help="segment workload: table heights (BASELINE configs #4 / #3)")
    ap.add_argument("--stark-config", default="standard_fast", choices=sorted(STARK_CONFIGS))
    ap.add_argument("--parallelism", default="segments", choices=["segments", "tables"])
    ap.add_argument("--log-n", type=int, default=20, help="cpu_table workload: log2 of the trace length (BASELINE config #2: 20)")
    ap.add_argument("--shrink", type=int, default=0, help="segment workload: make every table 2^shrink times shorter (smoke runs)")
    ap.add_argument("--cpu-shrink", type=int, default=3, help="in-arm cpu_baseline: the sample proves tables 2^k times shorter")
    ap.add_argument("--reference-budget", type=float, default=280.0, help="reference arm: stop starting new full-size steps after this many seconds")
    ap.add_argument("--streams", type=int, default=4, help="segments in flight per GPU (parallelism=segments); measured (profiles/r2t_streams_sweep.jsonl) "
                                                             "2: 3.94, 3: 4.13, 4: 4.17, 5: 4.18 proofs/s (e2e 3.91, 4.07, 4.14, 4.14); round 1: 1: 3.34, 3: 3.95, 4: 3.86")
    ap.add_argument("--stagger-ms", type=float, default=80.0, help="start offset between the segment streams of a GPU (inside the timed region)")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (configs #2, #3, #5, table_sharded, e2e_finish_on_device)")
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
