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
