"""The compiled-language host side: include/zkgpu.hpp (C++17 mirror of the reference interface over the C ABI) through the harness
tests/native/host_mirror.cpp.  CPU: it builds and links against libzkgpu.so, fails loudly without a device (no CPU path), its Challenger
agrees with the oracle's, its typed StarkProof decoder round-trips oracle proofs.  GPU: prove_with_traces / PolynomialBatch::from_values /
the abort signal from C++ on a segment handed over as a trace file, proofs bit for bit against the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import traces
from tests.oracle_lib import orc_prove_segment, orc_prove_table, TEST_CONFIG, STANDARD_FAST, DEFAULT_LABELS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "host_mirror.cpp")
BIN = os.path.join(HERE, "native", "host_mirror")
LIBDIR = os.path.join(ROOT, "zk_evm_b200")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def harness():
    import zk_evm_b200
    zk_evm_b200.lib()                      # libzkgpu.so must exist (built by __graft_entry__.build())
    deps = [SRC, os.path.join(ROOT, "include", "zkgpu.hpp"), os.path.join(ROOT, "include", "zkgpu.h"),
            os.path.join(LIBDIR, "csrc", "stark", "proof.h")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", BIN, SRC,
                               "-L", LIBDIR, "-lzkgpu", "-lpthread", "-Wl,-rpath," + LIBDIR])
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(p for p in (env.get("LD_LIBRARY_PATH"), "/usr/local/cuda/lib64") if p)

    def run(*args, ok=True):
        r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=900)
        if ok:
            assert r.returncode == 0, (r.stdout, r.stderr)
        return r
    return run


def test_cpp_host_builds_and_has_no_cpu_path(harness):
    import torch
    out = harness("nodevice").stdout
    if not torch.cuda.is_available():
        assert out.startswith("error -2"), out            # ZKGPU_ERR_CUDA


def test_cpp_challenger_matches_oracle(harness, oracle):
    words = [int(x) for x in harness("challenger", 20).stdout.split()]
    ops = [("o", v) for v in range(1, 21)] + [("c",)] * 3 + [("o", 7), ("k",)]
    ch, st = oracle.challenger_run(ops)
    assert words[:3] == [int(c) for c in ch[:3]]
    assert words[3:15] == [int(x) for x in st]
    ch2, _ = oracle.challenger_run(ops + [("c",)])
    assert words[15] == int(ch2[3])


def test_cpp_segment_stream_scheduling_logic(harness):
    """zkgpu::SegmentStream with a stand-in prover: proofs in segment order from 3 workers, at most 3 segments alive, a failing segment
    and abort() stop the stream early (the product instance, segment_prover(), plugs Context + prove_with_traces into the same template)"""
    out = harness("stream").stdout.split()
    f = dict(zip(out[0::2], out[1::2]))
    assert f["order"] == "1" and f["failure"] == "1" and f["abort"] == "1"
    assert int(out[out.index("failure") + 3]) < 500 and int(out[out.index("abort") + 3]) < 100000
    # admission by device memory (with_memory_budget): two side by side / one at a time / all three, and the C++ estimate of a segment's
    # device bytes equals the Python scheduler's
    from zk_evm_b200.scheduler import estimate_segment_bytes

    class Shape:
        def __init__(self, c, lg):
            self.shape = (c, 1 << lg)
    widths = [116, 71, 85, 2431, 438, 523, 30, 12, 12]
    want = estimate_segment_bytes([Shape(widths[t], lg) for t, lg in enumerate([17, 14, 19, 17, 13, 16, 21, 19, 19])])
    assert out[out.index("budget") + 1] == "1"
    assert abs(int(out[out.index("estimate") + 1]) - want) <= 2


def _random_public_values(seed):
    from zk_evm_b200 import public_values as pvm
    rng = np.random.default_rng(seed)
    h = lambda: rng.bytes(32)
    r = lambda bits: int.from_bytes(rng.bytes(bits // 8), "big")
    return pvm.PublicValues(
        pvm.TrieRoots(h(), h(), h()), pvm.TrieRoots(h(), h(), h()),
        pvm.BlockMetadata(rng.bytes(20), r(32), r(32), r(32), h(), r(32), r(32), r(64), r(32), r(64), r(64), h(), [r(256) for _ in range(8)]),
        pvm.BlockHashes([h() for _ in range(256)], h()),
        pvm.ExtraBlockData(h(), [int(x) for x in rng.integers(0, 0xFFFFFFFF00000001, size=4, dtype=np.uint64)], r(32), r(32), r(32), r(32)))


def _pack_public_values(pv):
    be = lambda x: int(x).to_bytes(32, "big")
    m, e = pv.block_metadata, pv.extra_block_data
    out = b"".join(bytes(x) for rr in (pv.trie_roots_before, pv.trie_roots_after) for x in (rr.state_root, rr.transactions_root, rr.receipts_root))
    out += bytes(m.block_beneficiary) + be(m.block_timestamp) + be(m.block_number) + be(m.block_difficulty) + bytes(m.block_random)
    out += be(m.block_gaslimit) + be(m.block_chain_id) + be(m.block_base_fee) + be(m.block_gas_used) + be(m.block_blob_gas_used)
    out += be(m.block_excess_blob_gas) + bytes(m.parent_beacon_block_root) + b"".join(be(w) for w in m.block_bloom)
    out += b"".join(bytes(x) for x in pv.block_hashes.prev_hashes) + bytes(pv.block_hashes.cur_hash)
    out += bytes(e.checkpoint_state_trie_root) + np.array(e.checkpoint_consolidated_hash, dtype="<u8").tobytes()
    out += be(e.txn_number_before) + be(e.txn_number_after) + be(e.gas_used_before) + be(e.gas_used_after)
    return out


def test_public_values_flattening_python_and_cpp(harness, tmp_path):
    """observe_public_values as an element list (get_challenges.rs:202-227): the reference's own sizes, a hand-derived root, and the
    C++ and Python host mirrors against each other on random values"""
    from zk_evm_b200 import public_values as pvm
    pv = pvm.PublicValues()
    assert len(pvm.flatten_public_values(pv)) == 24 * 2 + 97 + 2056 + 16          # proof.rs:652-655, 981, 1193, 1382, 1469
    assert len(pvm.flatten_public_values(pv, eth_mainnet=False)) == 24 * 2 + 85 + 2056 + 16
    pv.trie_roots_before.state_root = bytes(range(32))
    # observe_root (get_challenges.rs:11-19): root.into_uint().0 = little-endian u64 limbs of the big-endian integer, low half first
    assert [int(x) for x in pvm.flatten_public_values(pv)[:8]] == [0x1c1d1e1f, 0x18191a1b, 0x14151617, 0x10111213, 0x0c0d0e0f, 0x08090a0b,
                                                                   0x04050607, 0x00010203]
    pv = _random_public_values(17)
    path = tmp_path / "pv.bin"
    path.write_bytes(_pack_public_values(pv))
    got = np.array([int(x) for x in harness("pubvals", path).stdout.split()], dtype=np.uint64)
    assert np.array_equal(got, pvm.flatten_public_values(pv))
    # u256_to_u32 refuses what does not fit (ProgramError::IntegerTooLarge), on both sides
    pv.block_metadata.block_timestamp = 1 << 32
    with pytest.raises(pvm.IntegerTooLarge):
        pvm.flatten_public_values(pv)
    path.write_bytes(_pack_public_values(pv))
    assert harness("pubvals", path, ok=False).returncode != 0


def test_memory_extra_looking_values_python_and_cpp(harness, tmp_path):
    """get_memory_extra_looking_values / _sum (verifier.rs:319-512, 547-737): the C++ and Python host mirrors against each other on random
    public values (the Python one is checked field by field in test_public_values.py)"""
    from zk_evm_b200 import public_values as pvm
    rng = np.random.default_rng(23)
    pv = _random_public_values(19)
    u = lambda: int.from_bytes(rng.bytes(32), "big")
    pv.registers_before = pvm.RegistersData(u(), 1, u(), u(), u(), u())
    pv.registers_after = pvm.RegistersData(u(), 0, u(), u(), u(), u())
    kh, klen, beta, gamma = rng.bytes(32), 61234, int(rng.integers(1, 1 << 62)), int(rng.integers(1, 1 << 62))
    be = lambda x: int(x).to_bytes(32, "big")
    blob = _pack_public_values(pv)
    for r in (pv.registers_before, pv.registers_after):
        blob += b"".join(be(x) for x in (r.program_counter, r.is_kernel, r.stack_len, r.stack_top, r.context, r.gas_used))
    blob += kh + be(klen) + be(beta) + be(gamma)
    path = tmp_path / "pv_extra.bin"
    path.write_bytes(blob)
    got = [int(x) for x in harness("pubvals", path).stdout.split()]
    rows = pvm.memory_extra_looking_values(pv, kh, klen)
    assert len(got) == 13 * 301 + 1
    assert got[:-1] == [v for r in rows for v in r]
    assert got[-1] == pvm.memory_extra_looking_sum(pv, beta, gamma, kh, klen)


@pytest.mark.parametrize("table,lg,cfg", [(traces.T_MEM_AFTER, 7, TEST_CONFIG), (traces.T_LOGIC, 6, STANDARD_FAST)])
def test_cpp_proof_decoder_on_oracle_proofs(harness, oracle, tmp_path, table, lg, cfg):
    tr = traces.memcont_trace(lg, 3) if table == traces.T_MEM_AFTER else traces.logic_trace(lg, 3)
    bg = np.array([5, 6, 7, 8], dtype=np.uint64)[:2 * cfg[1]]
    proof, _ = orc_prove_table(oracle, table, cfg, tr, bg, np.arange(12, dtype=np.uint64))
    path = tmp_path / "proof.words"
    proof.astype("<u8").tofile(path)
    out = harness("decode", path).stdout.split()
    f = dict(zip(out[0::2], out[1::2]))
    assert int(f["table"]) == table and int(f["degree_bits"]) == lg and int(f["roundtrip"]) == 1
    assert int(f["trace_cap"]) == int(f["quotient_cap"]) == 1 << cfg[3]
    assert int(f["local"]) == int(f["next"]) == traces.NUM_COLUMNS[table]
    assert int(f["quotient"]) == 2 * cfg[1] and int(f["queries"]) == cfg[7] and int(f["pow"]) == int(proof[-1])
    # a truncated proof is refused, not misread
    proof[:-5].astype("<u8").tofile(path)
    assert harness("decode", path, ok=False).returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,cfg", [("test", TEST_CONFIG), ("fast", STANDARD_FAST)])
def test_cpp_prove_with_traces_matches_oracle(harness, oracle, tmp_path, name, cfg):
    from zk_evm_b200 import trace_file
    tr = traces.valid_segment(seed=21, k=19)
    pv = np.arange(1000, 1037, dtype=np.uint64)
    seg, out = tmp_path / "segment.trace", tmp_path / "proofs.bin"
    trace_file.save(seg, tr, pv, DEFAULT_LABELS)
    r = harness("prove", seg, out, name)
    assert "proved 5 tables" in r.stdout, (r.stdout, r.stderr)
    want, bg, caps = orc_prove_segment(oracle, cfg, tr, pv)
    w = np.fromfile(out, dtype="<u8")
    pos = 0

    def take():
        nonlocal pos
        n = int(w[pos]); v = w[pos + 1:pos + 1 + n]; pos += 1 + n
        return v
    assert np.array_equal(take(), bg)
    for t in range(9):
        assert np.array_equal(take(), caps[t].ravel()), "table %d cap" % t
        p = take()
        assert (p.size == 0) == (want[t] is None)
        assert want[t] is None or np.array_equal(p, want[t]), "table %d proof differs from the oracle" % t
    assert pos == w.size
