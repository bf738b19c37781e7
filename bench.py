#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 STARK proving path (contract: see the task statement / DESIGN.md §measurement).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the zkgpu C ABI)
  python bench.py --impl reference --gpus N ...            # reference arm: the CPU restatement (oracle) on all host cores

A step = one complete proof of the workload through the reference-shaped API (PolynomialBatch.from_values ->
get_ctl_data -> prove_single_table): trace commitment, CTL/lookup auxiliary columns + commitment, quotient evaluation +
commitment, openings, FRI commit phase, proof of work, query answers.
  value : proofs/s with the trace already resident in HBM (device-to-device ingest), device time from CUDA events on
          the library's stream, max over ranks
  e2e   : the same through the host API with the trace in pinned host memory (H2D inside the timed region) and the
          serialised proof read back (D2H)
N > 1 : one process per GPU (torchrun), every rank proves its own instance of the workload ("weak" scaling: the path
shards by table / by segment with no data-path collective).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_CPU, T_MEMORY = 2, 6
TABLE_NAMES = {0: "ArithmeticStark", 1: "BytePackingStark", 2: "CpuStark", 3: "KeccakStark", 4: "KeccakSpongeStark",
               5: "LogicStark", 6: "MemoryStark", 7: "MemBeforeStark", 8: "MemAfterStark"}
STANDARD_FAST = (100, 2, 1, 4, 16, 4, 5, 84)
LABELS = (0x1234, 0x77, 0x4000, 0x5000)
BG = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
STATE0 = np.arange(1, 13, dtype=np.uint64)
METRIC = "segment proofs/sec"


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


def workload(args):
    import zk_evm_b200 as zk
    table = T_CPU if zk.lib().zkgpu_table_info(C.c_uint32(T_CPU), C.c_uint32(2), None, None, None, None) == 0 else T_MEMORY
    if args.table is not None:
        table = args.table
    info = zk.table_info(table, 2)
    return table, args.log_n, info


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
    table, log_n, info = workload(args)
    ncols, n = info["num_columns"], 1 << log_n
    na = info["num_lookup_columns"] + info["num_ctl_helper_columns"] + info["num_ctl_zs"]
    dev = torch.device("cuda", local_rank)
    ctx = zk.Context(local_rank)
    cfg = zk.StarkConfig(*STANDARD_FAST)
    labels = zk.KernelLabels(*LABELS)
    g = torch.Generator(device=dev)
    g.manual_seed(2 + rank)
    trace_dev = torch.randint(0, 2 ** 63 - 1, (ncols, n), dtype=torch.int64, device=dev, generator=g)   # canonical: < p
    trace_host = torch.empty((ncols, n), dtype=torch.int64, pin_memory=True)
    trace_host.copy_(trace_dev)
    trace_np = trace_host.numpy().view(np.uint64)
    torch.cuda.synchronize()
    pow_w = [None]

    def step_device():
        tb = zk.PolynomialBatch.from_device_values(ctx, trace_dev.data_ptr(), ncols, n, keep_values=True)
        ctl = zk.get_ctl_data(ctx, table, tb, BG, 2)
        proof, _ = zk.prove_single_table(ctx, table, cfg, tb, ctl, STATE0, labels=labels)
        nwords = len(proof.words)
        pow_w[0] = int(proof.words[-1])
        proof.free(); ctl.free(); tb.free()
        return nwords

    def step_host():
        tb = zk.PolynomialBatch.from_values(ctx, trace_np, keep_values=True)
        ctl = zk.get_ctl_data(ctx, table, tb, BG, 2)
        proof, _ = zk.prove_single_table(ctx, table, cfg, tb, ctl, STATE0, labels=labels)
        nwords = len(proof.words)
        proof.free(); ctl.free(); tb.free()
        return nwords

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        ctx.timer_start()
        t0 = time.perf_counter()
        for _ in range(steps):
            nw = fn()
        ms = ctx.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        if dist is not None:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall, nw

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.stats()["kernel_launches"]
    ms, wall, nwords = timed(step_device, args.steps)
    launches = ctx.stats()["kernel_launches"] - l0
    for _ in range(min(args.warmup, 1)):
        step_host()
    e_ms, e_wall, _ = timed(step_host, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: the Poseidon leaf hash over the 2n LDE rows of the trace (profiles/: launch-list share)
        N = 2 * n
        lh_ms = ctx.bench_leaf_hash(ncols, N, 3)
        lh_bytes = (8.0 * ncols + 32.0) * N
        ntt_ms, _ = ctx.bench_ntt(ncols, n, 3)
        roof = {"bound": "hbm", "kernel": "leaf_hash_kernel (Poseidon sponge over %d LDE rows x %d cols)" % (N, ncols),
                "achieved": lh_bytes / lh_ms / 1e6, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": lh_bytes / lh_ms / 1e6 / peak, "traffic": None,
                "perms_per_s": ((ncols + 7) // 8) * N / lh_ms * 1e3,
                "note": "Poseidon is integer-ALU bound (~2e4 int ops per 64 B); HBM fraction reported because the metric asks for it",
                "ntt": {"kernel": "ntt_dif_pass_kernel size-n batch of %d cols" % ncols, "achieved": 16.0 * ncols * n / ntt_ms / 1e6,
                        "frac": 16.0 * ncols * n / ntt_ms / 1e6 / peak, "unit": "GB/s"}}
        cpu = None if args.no_cpu_baseline else cpu_baseline(table, args, full_log_n=log_n)
        line = {"metric": METRIC, "value": world * args.steps / (ms / 1e3), "unit": "proofs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks, exact)", "data": "synthetic",
                "config": {"workload": "single %s table prove, trace 2^%d rows x %d cols (+%d aux cols), standard_fast_config; "
                                       "uniform random trace (synthetic, seed 2+rank)" % (TABLE_NAMES[table], log_n, ncols, na),
                           "l2": "inputs larger than L2 (trace %.0f MiB, LDE %.0f MiB)" % (8.0 * ncols * n / 2 ** 20, 16.0 * ncols * n / 2 ** 20),
                           "timing": "CUDA events on the library stream, max over ranks", "pow_witness": pow_w[0]},
                "wall_ms_per_step": wall / args.steps,
                "e2e": {"value": world * args.steps / (e_ms / 1e3), "unit": "proofs/s", "h2d_bytes_per_step": 8 * ncols * n,
                        "d2h_bytes_per_step": 8 * nwords, "ms_per_step": e_ms / args.steps, "wall_ms_per_step": e_wall / args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def oracle_prove_time(table, log_n, threads=None):
    from tests import oracle_lib
    orc = oracle_lib.load()
    if threads:
        orc.lib.orc_set_num_threads(int(threads))
    ncols = orc.lib.orc_table_num_columns(C.c_uint32(table))
    rng = np.random.default_rng(2)
    tr = rng.integers(0, 2 ** 63 - 1, size=(ncols, 1 << log_n), dtype=np.uint64)
    t0 = time.perf_counter()
    oracle_lib.orc_prove_table(orc, table, STANDARD_FAST, tr, BG, STATE0, labels=LABELS)
    return time.perf_counter() - t0, orc.lib.orc_num_threads()


def cpu_baseline(table, args, full_log_n):
    """The CPU restatement (oracle, C++/OpenMP) on the host cores, on a bounded sample: the same table at 2^sample rows,
    scaled linearly in rows to the workload size (n log n terms make this slightly favourable to the CPU)."""
    sample = min(full_log_n, args.cpu_sample_log_n)
    secs, threads = oracle_prove_time(table, sample)
    scale = float(1 << (full_log_n - sample))
    return {"value": 1.0 / (secs * scale), "unit": "proofs/s", "cores": int(threads), "kind": "port",
            "sample": "oracle prove_table of the same table at 2^%d rows took %.2f s on %d threads; scaled x%d in rows to 2^%d"
                      % (sample, secs, threads, int(scale), full_log_n)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, ROOT)
    from tests import oracle_lib
    orc = oracle_lib.load()
    table = args.table
    if table is None:
        table = T_CPU if orc.lib.orc_table_supported(C.c_uint32(T_CPU)) else T_MEMORY
    ncols = orc.lib.orc_table_num_columns(C.c_uint32(table))
    sample = min(args.log_n, args.cpu_sample_log_n)
    scale = float(1 << (args.log_n - sample))
    for _ in range(min(args.warmup, 1)):
        oracle_prove_time(table, max(8, sample - 4))
    t = 0.0
    threads = 1
    for _ in range(args.steps):
        s, threads = oracle_prove_time(table, sample)
        t += s
    per_step = t / args.steps * scale
    world = int(os.environ.get("WORLD_SIZE", "1"))
    val = 1.0 / per_step
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (Goldilocks, exact)", "data": "synthetic",
        "config": {"workload": "single %s table prove, trace 2^%d rows x %d cols, standard_fast_config; CPU restatement "
                               "(oracle/, C++17 + OpenMP) of the plonky2/starky prover: the Rust reference cannot be built here "
                               "(no cargo/rustc, crates not vendored)" % (TABLE_NAMES[table], args.log_n, ncols)},
        "cpu_baseline": {"value": val, "unit": "proofs/s", "cores": int(threads), "kind": "port",
                         "sample": "each step proves the table at 2^%d rows; time scaled x%d in rows to 2^%d" % (sample, int(scale), args.log_n)},
        "e2e": {"value": val, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 of the trace length (BASELINE config #2: 20)")
    ap.add_argument("--table", type=int, default=None, help="table id (default: CpuStark)")
    ap.add_argument("--cpu-sample-log-n", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiler runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
