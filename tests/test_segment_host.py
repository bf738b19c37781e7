"""Host-side segment logic without a GPU: the transcript of the C library against the oracle's, and the table-sharded
prove_with_traces over a 2-rank gloo group with the oracle standing in for the device (checker used as test double)."""
import os
import socket
import numpy as np
import pytest
import torch.multiprocessing as mp
import zk_evm_b200 as zk
from tests import traces, oracle_lib
from tests.oracle_lib import TEST_CONFIG, orc_prove_table, orc_prove_segment, orc_verify_segment

PUBLIC_VALUES = np.arange(1000, 1000 + 37, dtype=np.uint64)


def test_challenger_matches_oracle(oracle):
    rng = np.random.default_rng(5)
    ops, ch, got = [], zk.Challenger(), []
    for _ in range(200):
        k = rng.integers(0, 3)
        if k == 0:
            xs = oracle_lib.rand_field(rng, int(rng.integers(1, 20)))
            ops += [('o', int(x)) for x in xs]
            ch.observe_elements(xs)
        elif k == 1:
            m = int(rng.integers(1, 12))
            ops += [('c',)] * m
            got += list(ch.get_n_challenges(m))
        else:
            ops.append(('k',))
            ch.compact()
    want, st = oracle.challenger_run(ops + [('k',)])
    assert np.array_equal(np.array(got, dtype=np.uint64), want)
    assert np.array_equal(ch.compact(), st)


def test_challenger_rejects_non_canonical():
    ch = zk.Challenger()
    with pytest.raises(zk.ZkGpuError):
        ch.observe_elements(np.array([0xFFFFFFFF00000001], dtype=np.uint64))


def test_segment_challenges_match_oracle(oracle):
    tr = traces.valid_segment(seed=1)
    _, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PUBLIC_VALUES)
    in_use = [t is not None for t in tr]
    bg2, _ = zk.segment_challenges(caps, in_use, TEST_CONFIG[3], PUBLIC_VALUES, TEST_CONFIG[1])
    assert np.array_equal(bg, bg2)
    with pytest.raises(zk.ZkGpuError):      # a mandatory table cannot be left out
        zk.segment_challenges(caps, [False] + in_use[1:], TEST_CONFIG[3], PUBLIC_VALUES, TEST_CONFIG[1])


def test_default_owner_is_a_partition():
    for w in (1, 2, 4, 8):
        o = zk.default_owner(w)
        assert len(o) == 9 and set(o) <= set(range(w)) and (w > 8 or len(set(o)) == min(w, 9) or w == 8)


class OracleBackend:
    """test double for ZkGpuBackend: same interface, computed by the CPU oracle"""

    def __init__(self, cfg):
        self.cfg, self.orc = cfg, oracle_lib.load()

    def commit(self, table, trace):
        return (table, np.ascontiguousarray(trace))

    def cap(self, handle):
        return self.orc.commit(handle[1], self.cfg[2], self.cfg[3])[3]

    def segment_challenges(self, caps, in_use, pv):
        return zk.segment_challenges(caps, in_use, self.cfg[3], pv, self.cfg[1])

    def begin(self, table, handle, beta_gamma):
        return (table, handle[1], np.asarray(beta_gamma))

    def finish(self, job, state, forced_pow=None):
        table, trace, bg = job
        return orc_prove_table(self.orc, table, self.cfg, trace, bg, state, forced_pow=forced_pow)


class OracleSplitCommit:
    """test double for zk_evm_b200.SplitCommit: the same three steps (column slice through ifft + LDE and all-gather; leaf block + Merkle
    levels under the rank's cap entries and all-gather; assembly on the owner) computed with the oracle's primitives and exchanged over
    gloo, so that the split branch of prove_with_traces_sharded — ordering of the collectives, slicing, block layout — runs without a GPU"""

    def __init__(self, backend, comm, table, trace):
        self.be, self.comm, self.table, self.trace = backend, comm, table, np.ascontiguousarray(trace)
        orc, k, r = backend.orc, comm.world, comm.rank
        ncols, n = self.trace.shape
        self.ncols, self.n, self.N = ncols, n, 2 * n
        cpr = -(-ncols // k)
        c0, c1 = min(ncols, r * cpr), min(ncols, (r + 1) * cpr)
        slot = np.zeros((cpr, self.N), dtype=np.uint64)
        if c1 > c0:
            coef = orc.ntt(self.trace[c0:c1], 1)                                    # ifft
            padded = np.concatenate([coef, np.zeros_like(coef)], axis=1)
            lde = orc.ntt(padded, 2, oracle_lib.GENERATOR)                          # coset fft on g <w_2n>, natural order
            lg = self.N.bit_length() - 1
            rev = np.array([int(("{:0%db}" % lg).format(j)[::-1], 2) for j in range(self.N)])
            slot[:c1 - c0] = lde[:, rev]                                            # rows in bit-reversed order, as the leaves are
        self.lde = comm.all_gather(slot).reshape(k * cpr, self.N)[:ncols]

    def hash_block(self):
        orc, k, r = self.be.orc, self.comm.world, self.comm.rank
        per = self.N // k
        level = orc.hash_rows_colmajor(np.ascontiguousarray(self.lde[:, r * per:(r + 1) * per]))   # (per, 4) leaf digests of the block
        packed = [level]
        while per * k > 16:                                                          # down to the cap level (cap_height 4)
            level = np.array([orc.two_to_one(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)], dtype=np.uint64).reshape(-1, 4)
            packed.append(level)
            per //= 2
        self.sizes = [len(l) for l in packed]
        self.packed = self.comm.all_gather(np.concatenate(packed))                  # (k, sum of the block's level sizes, 4)

    def finish(self, is_owner):
        if not is_owner:
            return None
        # the cap = last level of every block, block after block; must be the cap of the one-device commitment
        off = sum(self.sizes[:-1])
        cap = np.concatenate([self.packed[b, off:off + self.sizes[-1]] for b in range(self.comm.world)])
        want = self.be.orc.commit(self.trace, self.be.cfg[2], self.be.cfg[3])[3]
        assert np.array_equal(cap.reshape(-1, 4), want.reshape(-1, 4)), "split commitment of table %d: cap differs" % self.table
        return (self.table, self.trace)


OracleBackend.split_commit = lambda self, comm, table, trace: OracleSplitCommit(self, comm, table, trace)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = traces.valid_segment(seed=3)
        in_use = [t is not None for t in tr]
        owner = zk.default_owner(world)
        local = [t if (t is not None and owner[i] == rank) else None for i, t in enumerate(tr)]   # a rank holds only its tables
        ap = zk.prove_with_traces_sharded(OracleBackend(TEST_CONFIG), zk.TorchComm(), local, in_use, PUBLIC_VALUES, owner=owner)
        q.put((rank, ap.stark_proofs, ap.ctl_challenges, owner))
    finally:
        dist.destroy_process_group()


def _worker_split(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tr = traces.valid_segment(seed=3)
        in_use = [t is not None for t in tr]
        log_ns = [None if t is None else t.shape[1].bit_length() - 1 for t in tr]
        plan = zk.shard_plan(world, log_ns, split_min_bytes=8 * 12 * 128)          # Memory and Cpu are split, the small tables are not
        local = [t if (t is not None and plan.needs(rank)[i]) else None for i, t in enumerate(tr)]
        ap = zk.prove_with_traces_sharded(OracleBackend(TEST_CONFIG), zk.TorchComm(), local, in_use, PUBLIC_VALUES, plan=plan)
        q.put((rank, ap.stark_proofs, ap.ctl_challenges, plan.describe()))
    finally:
        dist.destroy_process_group()


def test_sharded_prove_with_split_commitments_two_ranks_gloo(oracle):
    """the table-sharded layout WITH split trace commitments (ShardPlan.split) on two gloo ranks: proofs == the single-process proofs"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_split, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, bg, _ = orc_prove_segment(oracle, TEST_CONFIG, traces.valid_segment(seed=3), PUBLIC_VALUES)
    assert len(res[0][3]["split_over_all_gpus"]) >= 2 and len(set(res[0][3]["owner"])) == 2
    for rank, proofs, ctl_ch, _ in res:
        assert np.array_equal(ctl_ch, bg)
        for t in range(9):
            assert (proofs[t] is None) == (want[t] is None)
            if want[t] is not None:
                assert np.array_equal(proofs[t], want[t]), "rank %d table %d differs from the single-process proof" % (rank, t)


def test_shard_plan_properties():
    import bench
    for w in (1, 2, 4, 8):
        plan = zk.shard_plan(w, bench.SEGMENT_LOG_NS)
        assert len(plan.owner) == 9 and set(plan.owner) <= set(range(w))
        assert (w == 1 and not any(plan.split)) or (w > 1 and plan.split[3] and plan.split[6] and not plan.split[1])   # Keccak, Memory split; BytePacking not
        for r in range(w):
            need = plan.needs(r)
            assert all(need[t] == (plan.split[t] or plan.owner[t] == r) for t in range(9))
    # an optional table that is not in use is nobody's
    plan = zk.shard_plan(4, [16, None, 16, None, None, None, 18, 16, None])
    assert plan.needs(0)[1] is False and not plan.split[1]
    # a split commitment hands every rank whole cap subtrees: other world sizes shard by owner only
    for w in (3, 5, 6, 7, 32):
        plan = zk.shard_plan(w, bench.SEGMENT_LOG_NS)
        assert not any(plan.split) and len(set(plan.owner)) == min(w, 9)
    assert not any(zk.shard_plan(4, bench.SEGMENT_LOG_NS, cap_height=1).split)      # 4 ranks, 2 cap entries
    assert any(zk.shard_plan(2, bench.SEGMENT_LOG_NS, cap_height=1).split)


def test_sharded_prove_two_ranks_gloo(oracle):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, bg, _ = orc_prove_segment(oracle, TEST_CONFIG, traces.valid_segment(seed=3), PUBLIC_VALUES)
    owners = res[0][3]
    assert len(set(owners)) == 2
    for rank, proofs, ctl_ch, _ in res:
        assert np.array_equal(ctl_ch, bg)
        for t in range(9):
            assert (proofs[t] is None) == (want[t] is None)
            if want[t] is not None:
                assert np.array_equal(proofs[t], want[t]), "rank %d table %d differs from the single-process proof" % (rank, t)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, res[0][1], PUBLIC_VALUES)
    assert ok, err


def test_sharded_prove_three_ranks_gloo(oracle):
    """a world size that is not a power of two (three GPUs of a node): shard_plan keeps whole tables with their owners, the relay and the
    cap all-gather work as for two ranks, proofs == the single-process proofs"""
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_split, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, bg, _ = orc_prove_segment(oracle, TEST_CONFIG, traces.valid_segment(seed=3), PUBLIC_VALUES)
    assert res[0][3]["split_over_all_gpus"] == [] and len(set(res[0][3]["owner"])) == 3
    for rank, proofs, ctl_ch, _ in res:
        assert np.array_equal(ctl_ch, bg)
        for t in range(9):
            assert (proofs[t] is None) == (want[t] is None)
            if want[t] is not None:
                assert np.array_equal(proofs[t], want[t]), "rank %d table %d differs from the single-process proof" % (rank, t)


def test_sharded_prove_local_comm_equals_segment(oracle):
    tr = traces.valid_segment(seed=4, k=17)
    in_use = [t is not None for t in tr]
    ap = zk.prove_with_traces_sharded(OracleBackend(TEST_CONFIG), zk.LocalComm(), tr, in_use, PUBLIC_VALUES)
    want, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PUBLIC_VALUES)
    assert np.array_equal(ap.ctl_challenges, bg) and np.array_equal(ap.trace_caps, caps)
    for t in range(9):
        assert (ap.stark_proofs[t] is None) == (want[t] is None)
        if want[t] is not None:
            assert np.array_equal(ap.stark_proofs[t], want[t])
