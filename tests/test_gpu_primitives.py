"""GPU parity: field / Poseidon / NTT / commitment kernels through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest
from tests.oracle_lib import P, GENERATOR, rand_field
import zk_evm_b200 as zk

pytestmark = pytest.mark.gpu


def test_poseidon_permutation_kats_and_random(ctx, oracle):
    rng = np.random.default_rng(11)
    st = rand_field(rng, (4096, 12))
    st[0] = 0
    st[1] = P - 1
    got = ctx.poseidon_permute(st)
    assert list(got[0][:4]) == [4330397376401421145, 14124799381142128323, 8742572140681234676, 14345658006221440202]
    assert np.array_equal(got, oracle.poseidon(st))


@pytest.mark.parametrize("width", [1, 2, 4, 5, 8, 9, 12, 16, 30, 85, 135])
def test_hash_rows(ctx, oracle, width):
    rng = np.random.default_rng(width)
    data = rand_field(rng, (width, 777))
    assert np.array_equal(ctx.poseidon_hash_rows(data), oracle.hash_rows_colmajor(data))


def test_empty_consolidated_blockhash_on_gpu(ctx):
    # proof.rs:505-510 known answer through the device sponge: one row of 2048 zeros
    out = ctx.poseidon_hash_rows(np.zeros((2048, 1), dtype=np.uint64))
    assert list(out[0]) == [5498946765822202150, 10724662260254836878, 9161393967331872654, 5704373722058976135]


def test_reference_bytecode_and_smt_kats_on_gpu(ctx):
    # smt_trie/src/code.rs:71-84 (non-zero multi-block sponge input) and smt_trie/src/smt_test.rs:30-47 through the DEVICE permutation
    from tests.test_oracle_kats import SOME_CODE, SOME_CODE_HASH, SMT_SINGLE_LEAF_ROOT, _hash_contract_bytecode, smt_single_leaf_root

    class Dev:      # the two helpers only need .poseidon(state) -> states
        @staticmethod
        def poseidon(st):
            return ctx.poseidon_permute(np.ascontiguousarray(st, dtype=np.uint64).reshape(-1, 12))
    assert _hash_contract_bytecode(Dev, SOME_CODE) == SOME_CODE_HASH
    assert smt_single_leaf_root(Dev.poseidon, [1, 0, 0, 0], 2) == SMT_SINGLE_LEAF_ROOT


@pytest.mark.parametrize("lg", [0, 1, 2, 3, 4, 7, 10, 12, 13, 14, 16, 17, 20, 21, 22])
def test_ntt_all_kinds(ctx, oracle, lg):
    rng = np.random.default_rng(lg)
    ncols = 3 if lg >= 16 else 9
    x = rand_field(rng, (ncols, 1 << lg))
    x[0] = P - 1
    assert np.array_equal(ctx.ntt(x), oracle.ntt(x, 0))
    assert np.array_equal(ctx.ntt(x, inverse=True), oracle.ntt(x, 1))
    assert np.array_equal(ctx.ntt(x, coset_shift=GENERATOR), oracle.ntt(x, 2, GENERATOR))
    assert np.array_equal(ctx.ntt(x, inverse=True, coset_shift=GENERATOR), oracle.ntt(x, 3, GENERATOR))


def test_largest_production_table_height(ctx, oracle):
    """The tallest trace the reference's default ranges allow is Memory at 2^23 rows (`.env` MEMORY_CIRCUIT_SIZE 17..24, half-open): its
    transforms have three passes over an 8 + 8 + 7 bit split, its commitment 2^24 leaves.  All four transform kinds and the whole
    commitment (coefficients, leaves, every digest, cap) against the oracle, on two columns"""
    lg = 23
    rng = np.random.default_rng(lg)
    x = rand_field(rng, (2, 1 << lg))
    x[0, :5] = P - 1
    assert np.array_equal(ctx.ntt(x), oracle.ntt(x, 0))
    assert np.array_equal(ctx.ntt(x, inverse=True), oracle.ntt(x, 1))
    assert np.array_equal(ctx.ntt(x, coset_shift=GENERATOR), oracle.ntt(x, 2, GENERATOR))
    assert np.array_equal(ctx.ntt(x, inverse=True, coset_shift=GENERATOR), oracle.ntt(x, 3, GENERATOR))
    b = zk.PolynomialBatch.from_values(ctx, x, rate_bits=1, cap_height=4)
    co, le, di = b.export()
    oco, ole, odi, ocap = oracle.commit(x, 1, 4)
    assert np.array_equal(co, oco) and np.array_equal(le, ole) and np.array_equal(b.cap, ocap) and np.array_equal(di, odi)
    b.free()


def test_ntt_roundtrip_config1(ctx):
    # BASELINE config #1 shape on the device: 2^16 points, 128 columns
    rng = np.random.default_rng(1)
    x = rand_field(rng, (128, 1 << 16))
    assert np.array_equal(ctx.ntt(ctx.ntt(x), inverse=True), x)


def test_ntt_large_roundtrip_and_linearity(ctx):
    # size-independent properties at a size the oracle would take long on: 2^22, round trip + linearity
    rng = np.random.default_rng(22)
    n = 1 << 22
    a = rand_field(rng, (1, n)); b = rand_field(rng, (1, n))
    fa, fb = ctx.ntt(a), ctx.ntt(b)
    assert np.array_equal(ctx.ntt(fa, inverse=True), a)
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    fs = ctx.ntt(s)
    assert np.array_equal(fs, ((fa.astype(object) + fb.astype(object)) % P).astype(np.uint64))


@pytest.mark.parametrize("ncols,lg,cap", [(1, 4, 4), (4, 5, 4), (5, 6, 4), (12, 7, 4), (30, 10, 4), (85, 12, 4),
                                          (3, 13, 4), (16, 16, 4), (2, 3, 4), (7, 6, 2), (9, 8, 0)])
def test_commit_values_matches_oracle(ctx, oracle, ncols, lg, cap):
    rng = np.random.default_rng(100 * ncols + lg)
    cols = rand_field(rng, (ncols, 1 << lg))
    b = zk.PolynomialBatch.from_values(ctx, cols, rate_bits=1, cap_height=cap)
    co, le, di = b.export()
    oco, ole, odi, ocap = oracle.commit(cols, 1, cap)
    assert np.array_equal(co, oco)
    assert np.array_equal(le, ole)
    assert np.array_equal(b.cap, ocap)
    assert np.array_equal(di, odi)
    b.free()


@pytest.mark.parametrize("rate_bits,ncols,lg,cap", [(0, 5, 8, 4), (2, 9, 7, 4), (3, 12, 10, 4), (3, 135, 12, 4), (3, 3, 3, 1), (4, 7, 6, 2)])
def test_commit_values_other_blowups_match_oracle(ctx, oracle, rate_bits, ncols, lg, cap):
    """PolynomialBatch::from_values is not only the STARK's rate_bits = 1: the recursion layer's circuits commit their wires with rate_bits = 3
    (CircuitConfig::standard_recursion_config; 135 wires in TEST_RECURSION_CONFIG, testing_utils.rs:55-72).  Coefficients, leaves in
    plonky2's bit-reversed row order, the recursive digest layout and the cap against the oracle for blow-ups 1, 4, 8 and 16."""
    rng = np.random.default_rng(1000 * rate_bits + ncols)
    cols = rand_field(rng, (ncols, 1 << lg))
    b = zk.PolynomialBatch.from_values(ctx, cols, rate_bits=rate_bits, cap_height=cap)
    co, le, di = b.export()
    oco, ole, odi, ocap = oracle.commit(cols, rate_bits, cap)
    assert np.array_equal(co, oco) and np.array_equal(le, ole) and np.array_equal(b.cap, ocap) and np.array_equal(di, odi)
    b.free()


def test_commit_coeffs_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(5)
    cf = rand_field(rng, (4, 1 << 9))
    b = zk.PolynomialBatch.from_coeffs(ctx, cf)
    co, le, di = b.export()
    oco, ole, odi, ocap = oracle.commit(cf, 1, 4, from_coeffs=True)
    assert np.array_equal(co, cf) and np.array_equal(le, ole) and np.array_equal(di, odi) and np.array_equal(b.cap, ocap)


def test_commit_errors(ctx):
    with pytest.raises(zk.ZkGpuError):
        zk.PolynomialBatch.from_values(ctx, np.zeros((2, 24), dtype=np.uint64))      # not a power of two
    with pytest.raises(zk.ZkGpuError):
        zk.PolynomialBatch.from_values(ctx, np.zeros((2, 4), dtype=np.uint64), cap_height=4)   # cap > leaves
