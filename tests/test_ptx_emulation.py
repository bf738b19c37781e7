"""CPU: the device instruction streams of the product's hot arithmetic — as nvcc emits them, inline-PTX carry chains included —
executed by the PTX interpreter in tools/ptx_emu.py (one thread, integer subset) and compared with exact big-integer arithmetic and
the oracle: Goldilocks field forms (gl.cuh), the lazy forms + the fast Poseidon permutation (poseidon_fast.cuh), the compile-time
power-of-two multiplications and the radix-16 butterfly network of the NTT (ntt_tile.cuh).  Needs nvcc (no GPU); what stays for
the GPU tests is ptxas' translation of this PTX."""
import os
import shutil
import subprocess
import sys
import numpy as np
import pytest
from tests import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
CSRC = os.path.join(ROOT, "zk_evm_b200", "csrc")
SRC = os.path.join(HERE, "native", "ptx_kernels.cu")
PTX = os.path.join(HERE, "native", "ptx_kernels.ptx")
P = oracle_lib.P
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
IN, OUT = 0x10000000, 0x20000000
EDGE = [0, 1, 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, P - 1, P - 2, 2 ** 63, 0x8000000080000000 % P]


@pytest.fixture(scope="module")
def emu():
    from ptx_emu import PtxEmu
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("gl.cuh", "poseidon.cuh", "poseidon_fast.cuh", "ntt_tile.cuh")]
    if not os.path.exists(PTX) or any(os.path.getmtime(d) > os.path.getmtime(PTX) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", PTX, SRC])
    return PtxEmu(open(PTX).read())


def run(emu, kernel, words, nout, extra=()):
    mem = {IN + 8 * i: int(w) for i, w in enumerate(words)}
    emu.run(kernel, [IN, OUT] + list(extra), mem)
    return [mem[OUT + 8 * i] for i in range(nout)]


def test_field_forms(emu):
    rng = np.random.default_rng(5)
    vals = EDGE + [int(v) for v in oracle_lib.rand_field(rng, (6,))]
    for a in vals:
        for b in vals:
            add, sub, mul, neg, red = run(emu, "k_field", [a, b], 5)
            assert (add, sub, mul, neg) == ((a + b) % P, (a - b) % P, a * b % P, (-a) % P), (a, b)
            assert red == (a + (b << 64)) % P, (a, b)


def test_lazy_forms_accept_any_u64(emu):
    rng = np.random.default_rng(6)
    vals = EDGE + [P, P + 1, 2 ** 64 - 1, 0xFFFFFFFEFFFFFFFF] + [int(v) for v in rng.integers(0, 2 ** 64, size=6, dtype=np.uint64)]
    for a in vals:
        for b in vals[:8]:
            mul, sqr, sbox, addc = run(emu, "k_lazy", [a, b], 4)
            assert (mul, sqr, sbox, addc) == (a * b % P, a * a % P, pow(a, 7, P), (a + b) % P), (a, b)


def test_power_of_two_multiplications(emu):
    shifts = [1, 12, 24, 31, 32, 36, 48, 60, 63, 64, 72, 84, 95]
    rng = np.random.default_rng(7)
    for a in EDGE + [int(v) for v in oracle_lib.rand_field(rng, (10,))]:
        got = run(emu, "k_pow2", [a], len(shifts))
        assert got == [a * pow(2, s, P) % P for s in shifts], a


def test_fast_permutation_ptx(emu):
    orc = oracle_lib.load()
    rng = np.random.default_rng(8)
    x = oracle_lib.rand_field(rng, (4, 12))
    x[1] = 0
    x[2] = P - 1
    want = orc.poseidon(x)
    for i in range(4):
        assert run(emu, "k_perm", x[i], 12) == [int(v) for v in want[i]]


def test_radix16_butterfly_network(emu):
    """ntt_dft_regs<4>: natural order in, bit-reversed order out, root w_16 = 2^12 (inverse: 2^-12)"""
    rng = np.random.default_rng(9)
    x = [int(v) for v in oracle_lib.rand_field(rng, (16,))]
    rev = [int("{:04b}".format(i)[::-1], 2) for i in range(16)]
    for inverse in (0, 1):
        w = pow(2, 12, P) if not inverse else pow(pow(2, 12, P), P - 2, P)
        got = run(emu, "k_dft16", x, 16, extra=[inverse])
        want = [sum(x[j] * pow(w, j * k, P) for j in range(16)) % P for k in range(16)]
        assert [got[rev[k]] for k in range(16)] == want


# ---- the production Merkle kernels themselves (csrc/merkle.cu compiled to PTX, one thread emulated) ---------------------------------
MERKLE_PTX = os.path.join(HERE, "native", "merkle.ptx")


@pytest.fixture(scope="module")
def merkle_emu():
    from ptx_emu import PtxEmu
    src = os.path.join(CSRC, "merkle.cu")
    deps = [src] + [os.path.join(CSRC, f) for f in ("gl.cuh", "poseidon.cuh", "poseidon_fast.cuh", "internal.h", "merkle.h")]
    if not os.path.exists(MERKLE_PTX) or any(os.path.getmtime(d) > os.path.getmtime(MERKLE_PTX) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", MERKLE_PTX, src])
    return PtxEmu(open(MERKLE_PTX).read())


@pytest.mark.parametrize("ncols", [1, 3, 4, 5, 8, 9, 16, 21])
def test_leaf_hash_kernel_one_row(merkle_emu, ncols):
    """leaf_hash_kernel: thread j hashes LDE row j = data[c * stride + j] over the columns (hash_or_noop: <= 4 columns are copied,
    zero padded; wider rows go through the overwrite-mode sponge, 8 columns per permutation)"""
    orc = oracle_lib.load()
    rng = np.random.default_rng(20 + ncols)
    nrows, stride, j = 3, 5, 1
    data = oracle_lib.rand_field(rng, (ncols, stride))
    mem = {IN + 8 * (c * stride + r): int(data[c, r]) for c in range(ncols) for r in range(stride)}
    merkle_emu.run("leaf_hash_kernel", [IN, stride, ncols, nrows, OUT], mem, tid=j, ctaid=0, ntid=128)
    got = [mem[OUT + 8 * (4 * j + k)] for k in range(4)]
    assert got == [int(v) for v in orc.hash_or_noop(data[:, j])]
    # a thread past the last row writes nothing
    mem2 = dict(mem)
    merkle_emu.run("leaf_hash_kernel", [IN, stride, ncols, nrows, OUT + 0x1000], mem2, tid=nrows, ctaid=0, ntid=128)
    assert not any(a >= OUT + 0x1000 for a in mem2)


def test_merkle_level_kernel_one_node(merkle_emu):
    orc = oracle_lib.load()
    rng = np.random.default_rng(30)
    kids = oracle_lib.rand_field(rng, (6, 4))           # three nodes: (0,1), (2,3), (4,5)
    mem = {IN + 8 * i: int(v) for i, v in enumerate(kids.ravel())}
    for node in range(3):
        merkle_emu.run("merkle_level_kernel", [IN, OUT, 3], mem, tid=node, ctaid=0, ntid=128)
        got = [mem[OUT + 8 * (4 * node + k)] for k in range(4)]
        assert got == [int(v) for v in orc.two_to_one(kids[2 * node], kids[2 * node + 1])]


def test_merkle_tail_kernel_block(merkle_emu):
    """merkle_tail_kernel (round 2: the levels under one cap entry in ONE launch): block s walks cap subtree s from its 8 input digests
    down to its root, a barrier between levels — the kernel's PTX for whole blocks of 4 threads (so that the level loop strides) against
    two_to_one of the oracle, for the second of two subtrees (the per-block offsets)"""
    import struct
    orc = oracle_lib.load()
    rng = np.random.default_rng(31)
    count0, nsub = 8, 2
    lvl = [oracle_lib.rand_field(rng, (nsub * count0, 4))]                      # level 0: 8 digests per subtree, subtree after subtree
    while len(lvl[-1]) > nsub:
        prev = lvl[-1]
        lvl.append(np.array([orc.two_to_one(prev[2 * i], prev[2 * i + 1]) for i in range(len(prev) // 2)], dtype=np.uint64))
    off, o = [], 0
    for l in lvl:                                                                # merkle_layout: level after level, 4 words per digest
        off.append(o)
        o += 4 * len(l)
    mem = {IN + 8 * i: int(v) for i, v in enumerate(lvl[0].ravel())}
    args = struct.pack("<Q12QII", IN, *(off + [0] * (12 - len(off))), count0, len(lvl) - 1)
    for block in range(nsub):
        merkle_emu.run_block("merkle_tail_kernel", [args], mem, ntid=4, ctaid=(block, 0))
    for k in range(1, len(lvl)):
        got = [mem[IN + 8 * (off[k] + i)] for i in range(4 * len(lvl[k]))]
        assert got == [int(v) for v in lvl[k].ravel()], "level %d" % k


# ---- the production NTT pass kernel: a whole 32-thread block with its barriers and shared-memory tile -----------------------------
NTT_PTX = os.path.join(HERE, "native", "ntt.ptx")


@pytest.fixture(scope="module")
def ntt_emu():
    from ptx_emu import PtxEmu
    src = os.path.join(CSRC, "ntt.cu")
    deps = [src] + [os.path.join(CSRC, f) for f in ("gl.cuh", "ntt_tile.cuh", "ntt.h", "internal.h")]
    if not os.path.exists(NTT_PTX) or any(os.path.getmtime(d) > os.path.getmtime(NTT_PTX) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", NTT_PTX, src])
    return PtxEmu(open(NTT_PTX).read())


def _pass_params(src, dst, n, log_n, m, r, t, strided, roots, interpass=0, prescale0=0):
    import struct
    # PassParams of csrc/ntt_tile.cuh: src, dst, src_stride, dst_stride, src_shift, log_n, m, r, t, strided, roots, interpass,
    # prescale0, prescale1, prescale_mask (+ padding) = 96 bytes
    return struct.pack("<QQQQIIIIIIQQQQII", src, dst, n, n, 0, log_n, m, r, t, strided, roots, interpass, prescale0, 0, 0, 0)


@pytest.mark.parametrize("lg,inverse", [(6, False), (6, True), (9, False), (5, True)])
def test_ntt_pass_kernel_block(ntt_emu, lg, inverse):
    """single-pass transforms (L <= 12: one final contiguous pass, tile = the whole transform, 32 threads): the kernel's PTX for a whole
    block — global loads fused into the first round, radix-16 rounds through the padded shared-memory tile, barriers — against the
    oracle's fft / ifft (DIF: natural order in, bit-reversed order out; the inverse leaves out the 1/n)"""
    orc = oracle_lib.load()
    rng = np.random.default_rng(40 + lg)
    n = 1 << lg
    x = oracle_lib.rand_field(rng, (n,))
    w = int(orc.lib.orc_root_of_unity(12))
    if inverse:
        w = pow(w, P - 2, P)
    ROOTS, SRC, DST = 0x30000000, IN, OUT
    mem = {ROOTS + 8 * k: pow(w, k, P) for k in range(4096)}
    mem.update({SRC + 8 * i: int(v) for i, v in enumerate(x)})
    params = _pass_params(SRC, DST, n, lg, lg, lg, 0, 0, ROOTS)
    ntt_emu.run_block("ntt_pass_kernelILi32ELb%dE" % int(inverse), [params], mem, ntid=32, ctaid=(0, 0))
    got = [mem[DST + 8 * i] for i in range(n)]
    rev = [int(("{:0%db}" % lg).format(i)[::-1], 2) for i in range(n)]
    want = orc.ntt(x, 1 if inverse else 0)[0]
    if inverse:
        want = [int(v) * n % P for v in want]                 # the kernel leaves the 1/n to the bit-reversal pass
    assert [got[rev[k]] for k in range(n)] == [int(v) for v in want]


@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_pass_tma_kernel_block(ntt_emu, inverse):
    """the production kernel for full tiles (round 2): 256 threads, the 4096-element tile staged by cp.async.bulk copies of 128 bytes per
    thread into the unpadded staging area, one mbarrier, then the three radix-16 rounds of a 12-bit final pass through the padded tile.
    The interpreter completes a bulk copy at once and treats the mbarrier wait as the block-wide phase boundary it is; everything else is
    the kernel's own PTX.  A whole 2^12 transform against the oracle."""
    orc = oracle_lib.load()
    rng = np.random.default_rng(52 + int(inverse))
    lg, n = 12, 1 << 12
    x = oracle_lib.rand_field(rng, (n,))
    w = int(orc.lib.orc_root_of_unity(12))
    if inverse:
        w = pow(w, P - 2, P)
    ROOTS, SRC, DST = 0x30000000, IN, OUT
    mem = {ROOTS + 8 * k: pow(w, k, P) for k in range(4096)}
    mem.update({SRC + 8 * i: int(v) for i, v in enumerate(x)})
    params = _pass_params(SRC, DST, n, lg, lg, lg, 0, 0, ROOTS)
    ntt_emu.run_block("ntt_pass_tma_kernelILb%dE" % int(inverse), [params], mem, ntid=256, ctaid=(0, 0))
    got = [mem[DST + 8 * i] for i in range(n)]
    rev = [int("{:012b}".format(i)[::-1], 2) for i in range(n)]
    want = orc.ntt(x, 1 if inverse else 0)[0]
    if inverse:
        want = [int(v) * n % P for v in want]
    assert [got[rev[k]] for k in range(n)] == [int(v) for v in want]


def test_ntt_two_pass_coset_transform_blocks(ntt_emu):
    """L = 13, the launcher's plan [8, 5]: a strided pass (tiles of 2^8 digit values x 16 contiguous elements, coset prescale on the
    first load, inter-pass twiddle on the store straight from the last round) and the final contiguous pass (128 sub-transforms of 32
    per tile), every block of both passes emulated with 256 threads; against the oracle's coset_fft"""
    orc = oracle_lib.load()
    lg, n = 13, 1 << 13
    G = 14293326489335486720
    rng = np.random.default_rng(50)
    x = oracle_lib.rand_field(rng, (n,))
    w12 = int(orc.lib.orc_root_of_unity(12))
    wN = int(orc.lib.orc_root_of_unity(lg))
    ROOTS, PRE, IP, BUF = 0x30000000, 0x40000000, 0x50000000, IN
    mem = {ROOTS + 8 * k: pow(w12, k, P) for k in range(4096)}
    mem.update({PRE + 8 * j: pow(G, j, P) for j in range(n)})
    mp = lg - 8
    mem.update({IP + 8 * ((kd << mp) + jp): pow(wN, (kd * jp) % n, P) for kd in range(256) for jp in range(1 << mp)})
    mem.update({BUF + 8 * i: int(v) for i, v in enumerate(x)})
    p1 = _pass_params(BUF, BUF, n, lg, lg, 8, 4, 1, ROOTS, interpass=IP, prescale0=PRE)          # in place, like ntt_dif
    for tile in range(n >> 12):
        ntt_emu.run_block("ntt_pass_kernelILi256ELb0E", [p1], mem, ntid=256, ctaid=(tile, 0))
    p2 = _pass_params(BUF, BUF, n, lg, lg - 8, lg - 8, 7, 0, ROOTS)
    for tile in range(n >> 12):
        ntt_emu.run_block("ntt_pass_kernelILi256ELb0E", [p2], mem, ntid=256, ctaid=(tile, 0))
    got = [mem[BUF + 8 * i] for i in range(n)]
    rev = [int("{:013b}".format(i)[::-1], 2) for i in range(n)]
    want = orc.ntt(x, 2, G)[0]
    assert [got[rev[k]] for k in range(n)] == [int(v) for v in want]


# ---- more production kernels through the interpreter -------------------------------------------------------------------------------
def _ptx_of(name, extra_deps=()):
    from ptx_emu import PtxEmu
    src = os.path.join(CSRC, name + ".cu")
    out = os.path.join(HERE, "native", name + ".ptx")
    deps = [src] + [os.path.join(CSRC, f) for f in ("gl.cuh", "internal.h", "stark_dev.h") + tuple(extra_deps)]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", out, src])
    return PtxEmu(open(out).read())


@pytest.fixture(scope="module")
def quotient_emu():
    return _ptx_of("quotient", ("quotient_kernel.cuh",))


@pytest.fixture(scope="module")
def fri_emu():
    return _ptx_of("fri", ("poseidon_fast.cuh", "poseidon.cuh"))


@pytest.fixture(scope="module")
def aux_emu():
    return _ptx_of("aux")


def _inv(x):
    return pow(x, P - 2, P)


def test_quotient_domain_table(quotient_emu):
    """quotient_domain_kernel: per LDE point x = g w_N^bitrev(j) the three selectors the fused evaluators multiply by — x - w_n^-1
    (vanishes on the last row), L_first(x) and L_last(x) — against their closed forms (Z_H(x) / n) / (x - 1), (Z_H(x) w_n^-1 / n) / (x - w_n^-1)"""
    import struct
    orc = oracle_lib.load()
    k = 5
    n, N, logN = 1 << k, 2 << k, k + 1
    G = 14293326489335486720
    wN, wn = int(orc.lib.orc_root_of_unity(logN)), int(orc.lib.orc_root_of_unity(k))
    last, ninv = _inv(wn), _inv(n)
    gn = pow(G, n, P)
    zh = [(gn - 1) % P, (-gn - 1) % P]
    c_first = [z * ninv % P for z in zh]
    c_last = [z * last % P * ninv % P for z in zh]
    args = struct.pack("<QQIIQQQQQQ", OUT, N, logN, 0, wN, last, c_first[0], c_first[1], c_last[0], c_last[1])
    mem = {}
    for j in range(N):
        quotient_emu.run("quotient_domain_kernel", [args], mem, tid=j % 128, ctaid=j // 128, ntid=128)
    for j in range(N):
        i = int(("{:0%db}" % logN).format(j)[::-1], 2)
        x = G * pow(wN, i, P) % P
        z = (pow(x, n, P) - 1) % P
        assert mem[OUT + 8 * j] == (x - last) % P
        assert mem[OUT + 8 * (N + j)] == z * ninv % P * _inv((x - 1) % P) % P
        assert mem[OUT + 8 * (2 * N + j)] == z * last % P * ninv % P * _inv((x - last) % P) % P


@pytest.mark.parametrize("T,nc", [(1, 1), (4, 2), (7, 2), (9, 1)])
def test_quotient_combine_kernel_point(quotient_emu, T, nc):
    """quotient_combine_kernel (round 2: the alpha-dependent half of the quotient when the constraint values were recorded ahead of the
    relay): storage position j -> out[k N + bitrev(j)] = zh_inv[i & 1] * sum_t alpha_k^(T - 1 - t) cons[t N + j], the Horner combination
    the fused evaluator's ConstraintConsumer makes; the four-at-a-time loop and its remainder"""
    rng = np.random.default_rng(70 + T)
    log_N, N = 3, 8
    cons = oracle_lib.rand_field(rng, (T, N))
    a0, a1, zh0, zh1 = (int(v) for v in oracle_lib.rand_field(rng, (4,)))
    CONS = 0x30000000
    for j in range(N):
        mem = {CONS + 8 * (t * N + jj): int(cons[t, jj]) for t in range(T) for jj in range(N)}
        quotient_emu.run("quotient_combine_kernel", [CONS, T, N, log_N, nc, a0, a1, zh0, zh1, OUT], mem, tid=j, ctaid=0, ntid=256)
        i = int("{:03b}".format(j)[::-1], 2)
        for k, a in enumerate((a0, a1)[:nc]):
            acc = 0
            for t in range(T):
                acc = (acc * a + int(cons[t, j])) % P
            assert mem[OUT + 8 * (k * N + i)] == acc * (zh1 if i & 1 else zh0) % P, (j, k)
        assert sum(1 for adr in mem if OUT <= adr < CONS) == nc           # nothing else is written
    mem = {}
    quotient_emu.run("quotient_combine_kernel", [CONS, T, N, log_N, nc, a0, a1, zh0, zh1, OUT], mem, tid=N, ctaid=0, ntid=256)
    assert not mem                                                         # a thread past the domain does nothing


def _ext_mul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def test_fri_fold_and_leaves_kernels(fri_emu):
    """fri_fold_kernel: out[i] = sum_t beta^t in[16 i + t] over the quadratic extension (coefficient-space folding of a commit-phase
    layer); fri_leaves_kernel: the column-major leaf matrix of a layer (16 extension elements = 32 words per leaf)"""
    rng = np.random.default_rng(60)
    M, ab = 64, 4
    re, im = oracle_lib.rand_field(rng, (M,)), oracle_lib.rand_field(rng, (M,))
    beta = (int(oracle_lib.rand_field(rng, (1,))[0]), int(oracle_lib.rand_field(rng, (1,))[0]))
    RE, IM, ORE, OIM = IN, IN + 0x10000, OUT, OUT + 0x10000
    mem = {RE + 8 * i: int(v) for i, v in enumerate(re)}
    mem.update({IM + 8 * i: int(v) for i, v in enumerate(im)})
    out_len = M >> ab
    import struct
    for i in range(out_len):
        fri_emu.run("fri_fold_kernel", [RE, IM, out_len, ab, struct.pack("<QQ", *beta), ORE, OIM], mem, tid=i, ctaid=0, ntid=256)
    for i in range(out_len):
        acc, bp = (0, 0), (1, 0)
        for t in range(1 << ab):
            term = _ext_mul(bp, (int(re[16 * i + t]), int(im[16 * i + t])))
            acc = ((acc[0] + term[0]) % P, (acc[1] + term[1]) % P)
            bp = _ext_mul(bp, beta)
        assert (mem[ORE + 8 * i], mem[OIM + 8 * i]) == acc
    rows, width = M >> ab, 2 << ab
    LEAVES = OUT + 0x40000
    for idx in range(rows * width):
        fri_emu.run("fri_leaves_kernel", [RE, IM, rows, ab, LEAVES], mem, tid=idx % 256, ctaid=idx // 256, ntid=256)
    for kcol in range(width):
        for r in range(rows):
            src = im if kcol & 1 else re
            assert mem[LEAVES + 8 * (kcol * rows + r)] == int(src[(r << ab) + (kcol >> 1)])


def test_pow_and_gather_kernels(fri_emu):
    """pow_kernel: candidate = base + thread; the transcript state with the candidate at the next input position goes through the
    permutation, and the candidate is a witness when the challenge popped next — state[7] — has `bits` leading zero bits
    (fri_proof_of_work; atomicMin keeps the smallest).  gather_addr_kernel: out[i] = *addrs[i] (the query answers of a table in one launch)."""
    import struct
    orc = oracle_lib.load()
    rng = np.random.default_rng(91)
    st = oracle_lib.rand_field(rng, (12,))
    pos, bits, base = 3, 2, 1000
    BEST = 0x30000000
    mem = {BEST: 2 ** 64 - 1}
    want = []
    for k in range(12):
        s = st.copy()
        s[pos] = base + k
        out = orc.poseidon(s.reshape(1, 12))[0]
        if int(out[7]) >> (64 - bits) == 0:
            want.append(base + k)
    assert 1 <= len(want) < 12                                             # the case has witnesses and non-witnesses
    for k in reversed(range(12)):                                          # any order: atomicMin
        fri_emu.run("pow_kernel", [struct.pack("<12Q", *(int(v) for v in st)), pos, bits, base, BEST], mem, tid=k, ctaid=0, ntid=128)
    assert mem[BEST] == want[0]
    mem = {BEST: 2 ** 64 - 1}
    fri_emu.run("pow_kernel", [struct.pack("<12Q", *(int(v) for v in st)), pos, bits, P - 3, BEST], mem, tid=3, ctaid=0, ntid=128)
    assert mem[BEST] == 2 ** 64 - 1                                         # candidates >= p are not field elements
    # gather
    ADDRS, DATA = 0x40000000, 0x50000000
    vals = [int(v) for v in oracle_lib.rand_field(rng, (9,))]
    order = [7, 0, 3, 3, 8]
    mem = {DATA + 8 * i: v for i, v in enumerate(vals)}
    mem.update({ADDRS + 8 * i: DATA + 8 * j for i, j in enumerate(order)})
    for i in range(6):
        fri_emu.run("gather_addr_kernel", [ADDRS, len(order), OUT], mem, tid=i, ctaid=0, ntid=128)
    assert [mem[OUT + 8 * i] for i in range(len(order))] == [vals[j] for j in order] and OUT + 8 * len(order) not in mem


def test_modular_scan_kernels(aux_emu):
    """the three scan kernels of the running-sum (Z) columns: per-tile inclusive scan (256 threads x 8 items, Hillis-Steele on the thread
    sums), exclusive scan of the tile totals, offset add — prefix sums for the logUp columns, suffix sums (reverse) for the CTL columns"""
    import struct
    rng = np.random.default_rng(61)
    n, ncols = 2500, 2
    data = oracle_lib.rand_field(rng, (ncols, n))
    DATA, JOBS, TOT = IN, IN + 0x100000, IN + 0x200000
    mem = {DATA + 8 * (c * n + i): int(data[c, i]) for c in range(ncols) for i in range(n)}
    # ScanJob {uint32 col; uint32 reverse}: job 0 = prefix sums of column 0, job 1 = suffix sums of column 1
    mem[JOBS] = 0 | (0 << 32)
    mem[JOBS + 8] = 1 | (1 << 32)
    ntiles = (n + 2047) // 2048
    for job in range(2):
        for tile in range(ntiles):
            aux_emu.run_block("scan_tiles_kernel", [JOBS, DATA, n, TOT, ntiles], mem, ntid=256, ctaid=(tile, job))
    for job in range(2):
        aux_emu.run_block("scan_totals_kernel", [TOT, ntiles], mem, ntid=256, ctaid=(job, 0))
    for job in range(2):
        for tile in range(ntiles):
            aux_emu.run_block("scan_add_kernel", [JOBS, DATA, n, TOT, ntiles], mem, ntid=256, ctaid=(tile, job))
    acc = 0
    for i in range(n):
        acc = (acc + int(data[0, i])) % P
        assert mem[DATA + 8 * i] == acc, i
    acc = 0
    for i in range(n - 1, -1, -1):
        acc = (acc + int(data[1, i])) % P
        assert mem[DATA + 8 * (n + i)] == acc, i


# ---- the fused quotient kernels, one LDE point at a time, against the oracle's evaluator -------------------------------------------------
import ctypes as C

QUOT_TABLES = {7: "MemBefore / MemAfter", 6: "Memory", 5: "Logic", 1: "BytePacking", 4: "KeccakSponge", 3: "Keccak", 0: "Arithmetic", 2: "Cpu"}
NUM_COLUMNS = (116, 71, 85, 2431, 438, 523, 30, 12, 12)
APOW_MAX = 1023


@pytest.fixture(scope="module")
def flat_host():
    src = os.path.join(HERE, "native", "flat_eval_host.cpp")
    lib = os.path.join(HERE, "native", "libflat_eval_host.so")
    oracle_dir = os.path.join(ROOT, "oracle")
    deps = [src] + [os.path.join(CSRC, "stark", f) for f in os.listdir(os.path.join(CSRC, "stark"))] + \
           [os.path.join(oracle_dir, f) for f in os.listdir(oracle_dir) if f.endswith(".h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-Wno-unknown-pragmas", "-I", CSRC, "-I", oracle_dir,
                               "-o", lib, src])
    h = C.CDLL(lib)
    h.flat_pack.restype = C.c_size_t
    return h


def _load_bytes(mem, base, data):
    data = bytes(data) + bytes(-len(data) % 8)
    for i in range(0, len(data), 8):
        mem[base + i] = int.from_bytes(data[i:i + 8], "little")


@pytest.mark.parametrize("table", sorted(QUOT_TABLES))
def test_fused_quotient_kernel_point(flat_host, table):
    """quotient_kernel<TABLE> (csrc/quotient_kernel.cuh + the table's single-source evaluator + the CTL / lookup interpreter, as PTX): one
    thread = one LDE point.  The inputs are arbitrary field elements in the kernel's own layout (column-major LDEs with bit-reversed
    rows, domain table, alpha-power table, packed descriptors), so the kernel is checked as the FUNCTION it is — rows i and i+2, the
    selectors, two challenges — against the oracle's eval_vanishing_poly on the same rows (times 1 / Z_H)."""
    import struct
    from ptx_emu import PtxEmu
    src = os.path.join(CSRC, "quotient_t%d.cu" % table)
    ptx = os.path.join(HERE, "native", "quotient_t%d.ptx" % table)
    deps = [src, os.path.join(CSRC, "quotient_kernel.cuh")] + [os.path.join(CSRC, "stark", f) for f in os.listdir(os.path.join(CSRC, "stark"))]
    if not os.path.exists(ptx) or any(os.path.getmtime(d) > os.path.getmtime(ptx) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", ptx, src])
    emu = PtxEmu(open(ptx).read())
    rng = np.random.default_rng(70 + table)
    u64p = C.POINTER(C.c_uint64)
    nc = NUM_COLUMNS[table]
    logN, N = 3, 8
    # descriptors
    offs, scal = (C.c_uint64 * 10)(), (C.c_uint32 * 7)()
    need = flat_host.flat_pack(C.c_uint32(table), C.c_uint32(2), None, C.c_size_t(0), offs, scal)
    buf = (C.c_uint8 * need)()
    flat_host.flat_pack(C.c_uint32(table), C.c_uint32(2), buf, C.c_size_t(need), offs, scal)
    naux = scal[2] + scal[3] + scal[4]
    TR, AUX, DOM, APOW, FLAT, QOUT = 0x100000000, 0x200000000, 0x300000000, 0x310000000, 0x320000000, 0x330000000
    trace = oracle_lib.rand_field(rng, (nc, N))
    aux = oracle_lib.rand_field(rng, (max(naux, 1), N))
    dom = oracle_lib.rand_field(rng, (3, N))
    al, be, ga, zh = (oracle_lib.rand_field(rng, (2,)) for _ in range(4))
    labels = np.array(oracle_lib.DEFAULT_LABELS, dtype=np.uint64)
    mem = {}
    _load_bytes(mem, TR, trace.tobytes())
    _load_bytes(mem, AUX, aux.tobytes())
    _load_bytes(mem, DOM, dom.tobytes())
    apow = np.array([[pow(int(a), e, P) for e in range(APOW_MAX + 1)] for a in al], dtype=np.uint64)
    _load_bytes(mem, APOW, apow.tobytes())
    _load_bytes(mem, FLAT, bytes(buf))
    args = struct.pack("<QQQQII", TR, AUX, QOUT, N, logN, 2) + struct.pack("<6Q", *(int(v) for v in list(al) + list(be) + list(ga))) + \
        struct.pack("<Q", DOM) + struct.pack("<2Q", int(zh[0]), int(zh[1])) + struct.pack("<Q", APOW) + \
        struct.pack("<10Q", *(FLAT + int(o) for o in offs)) + struct.pack("<7I", *scal) + b"\0" * 4 + struct.pack("<4Q", *(int(v) for v in labels))
    assert len(args) == 264
    kernel = "quotient_kernelILj%dEE" % table
    nthreads = 256 if table == 6 else 128
    p = lambda a: a.ctypes.data_as(u64p)
    fused = {}
    for j in (1, 6):
        emu.run(kernel, [args], mem, tid=j, ctaid=0, ntid=nthreads)
        i = int("{:03b}".format(j)[::-1], 2)
        jn = int("{:03b}".format((i + 2) % N)[::-1], 2)
        lv, nv = np.ascontiguousarray(trace[:, j]), np.ascontiguousarray(trace[:, jn])
        alv, anv = np.ascontiguousarray(aux[:, j]), np.ascontiguousarray(aux[:, jn])
        sel = np.ascontiguousarray(dom[:, j])
        a, b = np.zeros(2, np.uint64), np.zeros(2, np.uint64)
        na = C.c_uint32()
        r = flat_host.flat_eval_agrees(C.c_uint32(table), C.c_uint32(2), p(lv), p(nv), p(alv), p(anv), p(al), p(be), p(ga), p(sel), p(labels),
                                       p(a), p(b), C.byref(na))
        assert r == 1
        for k in range(2):
            assert mem[QOUT + 8 * (k * N + i)] == int(a[k]) * int(zh[i & 1]) % P, (table, j, k)
        fused[j] = [int(a[0]), int(a[1])]
    # constraints_record_kernel<TABLE> (round 2): the same evaluators through RecordConsumer write every constraint value of the point to
    # its own column — alpha-independent — and thread 0 reports how many there are; Horner in alpha over the recorded column of a point
    # must give what the fused kernel accumulated for it (before 1 / Z_H)
    CONS, CNT = 0x340000000, 0x350000000
    rargs = struct.pack("<QQQQQII", TR, AUX, CONS, CNT, N, logN, 0) + struct.pack("<4Q", *(int(v) for v in list(be) + list(ga))) + \
        struct.pack("<Q", DOM) + struct.pack("<10Q", *(FLAT + int(o) for o in offs)) + struct.pack("<7I", *scal) + b"\0" * 4 + \
        struct.pack("<4Q", *(int(v) for v in labels))
    assert len(rargs) == 232
    rkernel = "constraints_record_kernelILj%dEE" % table
    emu.run(rkernel, [rargs], mem, tid=0, ctaid=0, ntid=nthreads)
    T = mem[CNT] & 0xFFFFFFFF
    assert T == sum(1 for adr in mem if CONS <= adr < CNT) and T >= 1          # thread 0 wrote exactly T values, 8 N bytes apart
    emu.run(rkernel, [rargs], mem, tid=1, ctaid=0, ntid=nthreads)
    for k in range(2):
        acc = 0
        for t in range(T):
            acc = (acc * int(al[k]) + mem[CONS + 8 * (t * N + 1)]) % P
        assert acc == fused[1][k], (table, k)


# ---- device-side trace finishing: the Keccak row kernel ----------------------------------------------------------------------------------
def test_keccak_trace_kernel_rows():
    """keccak_trace_kernel (csrc/trace_gen.cu, as PTX), one thread = one trace row: first / middle / last round of a permutation, the last
    permutation and a padding row, against the Python restatement of keccak_stark.rs generate_trace_rows (tests/traces.py)."""
    from ptx_emu import PtxEmu
    from tests import traces
    src = os.path.join(CSRC, "trace_gen.cu")
    ptx = os.path.join(HERE, "native", "trace_gen.ptx")
    deps = [src] + [os.path.join(CSRC, "stark", f) for f in ("keccak_trace.h", "table_keccak.h", "logic_trace.h", "table_logic.h")]
    if not os.path.exists(ptx) or any(os.path.getmtime(d) > os.path.getmtime(ptx) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", ptx, src])
    emu = PtxEmu(open(ptx).read())
    rng = np.random.default_rng(41)
    nperm, log_n = 3, 7
    n = 1 << log_n
    inputs = rng.integers(0, 1 << 64, size=(nperm, 25), dtype=np.uint64)
    ts = rng.integers(1, 1 << 40, size=(nperm,), dtype=np.uint64)
    want, _ = traces.keccak_trace(log_n, inputs, ts)
    TSA, TR = IN + 8 * 25 * nperm, 0x100000000
    mem = {IN + 8 * i: int(w) for i, w in enumerate(inputs.ravel())}
    mem.update({TSA + 8 * i: int(w) for i, w in enumerate(ts)})
    for row in (0, 1, 11, 23, 24, 24 * 2 + 17, 24 * 3 - 1, 24 * 3, n - 1):
        emu.run("keccak_trace_kernel", [IN, TSA, nperm, n, TR], mem, tid=row % 128, ctaid=row // 128, ntid=128)
        got = [mem.get(TR + 8 * (c * n + row)) for c in range(2431)]
        assert got == [int(v) for v in want[:, row]], row
    assert len([a for a in mem if a >= TR]) == 9 * 2431      # nothing written outside the rows' own cells


def test_logic_trace_kernel_rows():
    """logic_trace_kernel (csrc/trace_gen.cu, as PTX), one thread = one row: the three operators and a padding row"""
    from ptx_emu import PtxEmu
    from tests import traces
    src = os.path.join(CSRC, "trace_gen.cu")
    ptx = os.path.join(HERE, "native", "trace_gen.ptx")
    deps = [src] + [os.path.join(CSRC, "stark", f) for f in ("keccak_trace.h", "table_keccak.h", "logic_trace.h", "table_logic.h")]
    if not os.path.exists(ptx) or any(os.path.getmtime(d) > os.path.getmtime(ptx) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", ptx, src])
    emu = PtxEmu(open(ptx).read())
    nops, log_n = 6, 3
    n = 1 << log_n
    ops = traces.logic_ops(nops, 77)
    ops[:3, 0] = [0, 1, 2]
    want = traces.logic_trace_from_ops(log_n, ops)
    TR = 0x100000000
    mem = {IN + 8 * i: int(w) for i, w in enumerate(ops.ravel())}
    for row in (0, 1, 2, 5, 6, 7):
        emu.run("logic_trace_kernel", [IN, nops, n, TR], mem, tid=row, ctaid=0, ntid=128)
        assert [mem.get(TR + 8 * (c * n + row)) for c in range(523)] == [int(v) for v in want[:, row]], row


def test_arithmetic_range_check_kernels():
    """arith_rc_init_kernel + arith_rc_hist_kernel (csrc/trace_gen.cu, as PTX): every thread of a small grid, one after the other, against
    ArithmeticStark::generate_range_checks restated with numpy (arithmetic_stark.rs:130-156); a cell >= 2^16 raises the flag."""
    from ptx_emu import PtxEmu
    src = os.path.join(CSRC, "trace_gen.cu")
    ptx = os.path.join(HERE, "native", "trace_gen.ptx")
    if not os.path.exists(ptx) or os.path.getmtime(src) > os.path.getmtime(ptx):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", ptx, src])
    emu = PtxEmu(open(ptx).read())
    rng = np.random.default_rng(3)
    cells = 300
    vals = rng.integers(0, 1 << 16, size=cells, dtype=np.uint64)
    vals[rng.integers(0, cells, size=120)] = 0                     # sparse, as real traces are
    vals[:4] = [65535, 1, 1, 65535]
    FREQ, BAD = 0x200000000, 0x300000000
    mem = {IN + 8 * i: int(v) for i, v in enumerate(vals)}
    mem[BAD] = 0
    ntid, nblocks = 8, 3
    for b in range(nblocks):
        for t in range(ntid):
            emu.run("arith_rc_hist_kernel", [IN, cells, FREQ, BAD], mem, tid=t, ctaid=b, ntid=ntid, nctaid=nblocks)
    want = np.bincount(vals.astype(np.int64), minlength=1 << 16)
    assert all(mem.get(FREQ + 8 * x, 0) == int(want[x]) for x in range(1 << 16))
    assert mem[BAD] == 0
    mem[IN + 8 * 7] = 1 << 16
    emu.run("arith_rc_hist_kernel", [IN, cells, FREQ, BAD], mem, tid=7, ctaid=0, ntid=ntid, nctaid=nblocks)
    assert mem[BAD] & 0xFFFFFFFF == 1
    # counter / zeroed frequencies
    CNT, FRQ2, n = 0x400000000, 0x500000000, 1 << 17
    for i in (0, 1, 65534, 65535, 65536, 65537, n - 1):
        mem[FRQ2 + 8 * i] = 99
        emu.run("arith_rc_init_kernel", [CNT, FRQ2, n], mem, tid=i % 256, ctaid=i // 256, ntid=256)
        assert mem[CNT + 8 * i] == min(i, 65535) and mem[FRQ2 + 8 * i] == 0


def test_memory_finish_kernels():
    """memory_stale_kernel + memory_finish_kernel (csrc/trace_gen.cu, as PTX; finish_row incl. the Fermat inversion of the timestamp and the
    histogram reductions), every thread of the grid one after the other, against the restated MemoryStark finishing (tests/traces.py)."""
    from ptx_emu import PtxEmu
    from tests import traces
    src = os.path.join(CSRC, "trace_gen.cu")
    ptx = os.path.join(HERE, "native", "trace_gen.ptx")
    deps = [src, os.path.join(CSRC, "gl.cuh")] + [os.path.join(CSRC, "stark", f) for f in ("memory_trace.h", "table_memory.h")]
    if not os.path.exists(ptx) or any(os.path.getmtime(d) > os.path.getmtime(ptx) for d in deps):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-ptx", "-o", ptx, src])
    emu = PtxEmu(open(ptx).read())
    log_n, stale = 5, (1, 2)
    n = 1 << log_n
    ops = traces.memory_sorted_ops(log_n, 21, nops=25)
    want = traces.memory_finish_reference(ops, stale)
    T, BAD, ST = 0x100000000, 0x300000000, 0x310000000
    mem = {BAD: 0}
    for k, c in enumerate(traces.MEM_OP_COLS):
        for i in range(n):
            mem[T + 8 * (c * n + i)] = int(ops[k, i])
    for c in (21, 22, 23, 29):                       # the library zero-fills the accumulated columns (cudaMemsetAsync)
        for i in range(n):
            mem[T + 8 * (c * n + i)] = 0
    for k, ctx in enumerate(stale):
        mem[ST + 8 * k] = ctx
    for k in range(len(stale)):
        emu.run("memory_stale_kernel", [ST, len(stale), n, T], mem, tid=k, ctaid=0, ntid=256)
    for i in range(n):
        emu.run("memory_finish_kernel", [T, n, BAD], mem, tid=i, ctaid=0, ntid=256)
    assert mem[BAD] == 0
    got = np.array([[mem[T + 8 * (c * n + i)] for i in range(n)] for c in range(30)], dtype=np.uint64)
    assert np.array_equal(got, want), [c for c in range(30) if not np.array_equal(got[c], want[c])]
