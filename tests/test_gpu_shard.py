"""The split commitment (include/zkgpu.h "S1 split over several devices") on ONE device: the k column slices and the k leaf blocks are
computed one after the other into the buffers an all-gather would fill, and the assembled batch must be the one-device commitment bit
for bit (coefficients, leaves, plonky2 digest layout, cap).  The multi-process exchange itself: tools/sharded_check.py under torchrun."""
import ctypes as C
import numpy as np
import pytest
from tests import oracle_lib

pytestmark = pytest.mark.gpu


def _p(addr):
    import zk_evm_b200 as zk
    return C.cast(C.c_void_p(int(addr)), zk._lib.u64p)


@pytest.mark.parametrize("k", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("ncols,lg", [(12, 8), (85, 7), (5, 9), (3, 6)])
def test_split_commit_equals_one_device_commit(ctx, k, ncols, lg):
    import torch
    import zk_evm_b200 as zk
    from zk_evm_b200._lib import check, lib
    n, N = 1 << lg, 2 << lg
    cols = oracle_lib.rand_field(np.random.default_rng(100 * k + ncols), (ncols, n))
    want = zk.PolynomialBatch.from_values(ctx, cols, rate_bits=1, cap_height=4, keep_values=True)
    dev = torch.device("cuda", 0)
    cpr = -(-ncols // k)
    lde = torch.zeros((k * cpr, N), dtype=torch.int64, device=dev)
    coef = torch.zeros((k * cpr, n), dtype=torch.int64, device=dev)
    vals = torch.zeros((k * cpr, n), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for r in range(k):      # what rank r of k would compute
        c0, c1 = min(ncols, r * cpr), min(ncols, (r + 1) * cpr)
        sl = np.ascontiguousarray(cols[c0:c1])
        off = 8 * r * cpr
        check(lib().zkgpu_lde_slice(ctx._h, _p(sl.ctypes.data) if c1 > c0 else None, 0, C.c_size_t(c1 - c0), C.c_size_t(n), C.c_uint32(1),
                                    _p(vals.data_ptr() + off * n), _p(coef.data_ptr() + off * n), _p(lde.data_ptr() + off * N)))
    words = C.c_size_t()
    check(lib().zkgpu_merkle_block_words(C.c_size_t(N), C.c_uint32(4), C.c_uint32(k), C.byref(words)))
    packed = torch.zeros((k, words.value), dtype=torch.int64, device=dev)
    for r in range(k):
        check(lib().zkgpu_merkle_block(ctx._h, _p(lde.data_ptr()), C.c_size_t(N), C.c_size_t(ncols), C.c_size_t(N), C.c_uint32(4),
                                       C.c_uint32(k), C.c_uint32(r), _p(packed.data_ptr() + 8 * r * words.value)))
    h = C.c_void_p()
    check(lib().zkgpu_batch_assemble(ctx._h, _p(vals.data_ptr()), _p(coef.data_ptr()), _p(lde.data_ptr()), _p(packed.data_ptr()), C.c_uint32(k),
                                     C.c_size_t(ncols), C.c_size_t(n), C.c_uint32(1), C.c_uint32(4), C.byref(h)))
    got = zk.PolynomialBatch(ctx, h)
    assert np.array_equal(got.cap, want.cap)
    for a, b in zip(got.export(), want.export()):
        assert np.array_equal(a, b)
    # the assembled batch works as the trace commitment of a proof (values borrowed too)
    ctx.sync()
    got.free(); want.free()
    assert np.array_equal(vals[:ncols].cpu().numpy().view(np.uint64), cols)


def test_split_commit_rejects_bad_blocks(ctx):
    import zk_evm_b200 as zk
    from zk_evm_b200._lib import lib
    words = C.c_size_t()
    assert lib().zkgpu_merkle_block_words(C.c_size_t(256), C.c_uint32(4), C.c_uint32(3), C.byref(words)) == -1     # not a power of two
    assert lib().zkgpu_merkle_block_words(C.c_size_t(256), C.c_uint32(4), C.c_uint32(32), C.byref(words)) == -1    # more blocks than cap entries
    assert lib().zkgpu_merkle_block_words(C.c_size_t(256), C.c_uint32(4), C.c_uint32(8), C.byref(words)) == 0
    # levels 256, 128, 64, 32, 16 digests, an eighth of each
    assert words.value == 4 * (32 + 16 + 8 + 4 + 2)


@pytest.mark.parametrize("table", list(range(9)))
@pytest.mark.parametrize("cfgname", ["test", "std"])
def test_precomputed_constraint_values_give_the_same_proof(ctx, oracle, table, cfgname):
    """zkgpu_ctx_set_precompute_constraints: the constraint values recorded in the first half of a table job and Horner-combined with the
    alphas in the second half give the proof the fused evaluator gives — and that is the oracle's proof, word for word"""
    import zk_evm_b200 as zk
    from tests import traces
    from tests.oracle_lib import STANDARD_FAST, TEST_CONFIG, DEFAULT_LABELS, orc_prove_table
    cfgw = STANDARD_FAST if cfgname == "std" else TEST_CONFIG
    cfg, labels = zk.StarkConfig(*cfgw), zk.KernelLabels(*DEFAULT_LABELS)
    lg = 16 if table == traces.T_ARITHMETIC else 9
    tr = traces.random_trace(table, lg, 70 + table)
    bg = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)[:2 * cfgw[1]]
    state = np.arange(1, 13, dtype=np.uint64)
    be = zk.ZkGpuBackend(ctx, cfg, labels, precompute_constraints=True)
    words, st = be.finish(be.begin(table, be.commit(table, tr), bg), state)
    want, st_want = orc_prove_table(oracle, table, cfgw, tr, bg, state)
    assert np.array_equal(words, want) and np.array_equal(st, st_want)
    # and the flag does not stick to the context: the next ordinary job runs the fused evaluator
    be2 = zk.ZkGpuBackend(ctx, cfg, labels)
    words2, _ = be2.finish(be2.begin(table, be2.commit(table, tr), bg), state)
    assert np.array_equal(words2, want)
