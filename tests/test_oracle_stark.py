"""CPU: the oracle's restated prover produces proofs the oracle's restated verifier accepts (valid traces), and rejects
tampered ones.  This is the self-consistency pin for every [UPSTREAM-RECALL] convention."""
import numpy as np
import pytest
from tests import traces, oracle_lib
from tests.oracle_lib import orc_prove_segment, orc_verify_segment, orc_prove_table, orc_verify_table, STANDARD_FAST, TEST_CONFIG, P

BG2 = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
STATE0 = np.arange(1, 13, dtype=np.uint64)


def _trace(table, lg, seed):
    if table in (traces.T_MEM_BEFORE, traces.T_MEM_AFTER):
        return traces.memcont_trace(lg, seed)
    if table == traces.T_LOGIC:
        return traces.logic_trace(lg, seed)
    if table == traces.T_MEMORY:
        return traces.memory_trace_simple(lg)
    raise ValueError(table)


@pytest.mark.parametrize("table,lg,cfg", [
    (traces.T_MEM_BEFORE, 7, TEST_CONFIG), (traces.T_MEM_AFTER, 5, TEST_CONFIG), (traces.T_MEM_BEFORE, 9, STANDARD_FAST),
    (traces.T_LOGIC, 6, TEST_CONFIG), (traces.T_LOGIC, 8, STANDARD_FAST),
    (traces.T_MEMORY, 6, TEST_CONFIG), (traces.T_MEMORY, 10, STANDARD_FAST), (traces.T_MEMORY, 4, TEST_CONFIG),
])
def test_prove_then_verify(oracle, table, lg, cfg):
    tr = _trace(table, lg, 42 + lg)
    bg = BG2[:2 * cfg[1]]
    proof, st = orc_prove_table(oracle, table, cfg, tr, bg, STATE0)
    ok, err, st2 = orc_verify_table(oracle, table, cfg, proof, bg, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)       # prover and verifier leave the shared transcript in the same state
    # tamper with the aux cap, a trace opening and the final polynomial
    for pos in (82, 212, len(proof) - 3):
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % P
        ok, err, _ = orc_verify_table(oracle, table, cfg, bad, bg, STATE0)
        assert not ok


def test_invalid_trace_is_rejected(oracle):
    tr = traces.logic_trace(6, 1)
    tr[515, 3] = (int(tr[515, 3]) + 1) % P        # wrong result limb
    proof, _ = orc_prove_table(oracle, traces.T_LOGIC, TEST_CONFIG, tr, BG2[:2], STATE0)
    ok, err, _ = orc_verify_table(oracle, traces.T_LOGIC, TEST_CONFIG, proof, BG2[:2], STATE0)
    assert not ok and "quotient" in err


def test_forced_pow_witness(oracle):
    tr = traces.memcont_trace(6, 3)
    proof, st = orc_prove_table(oracle, traces.T_MEM_AFTER, STANDARD_FAST, tr, BG2, STATE0)
    w = int(proof[-1])
    proof2, st2 = orc_prove_table(oracle, traces.T_MEM_AFTER, STANDARD_FAST, tr, BG2, STATE0, forced_pow=w)
    assert np.array_equal(proof, proof2) and np.array_equal(st, st2)
    with pytest.raises(RuntimeError):
        orc_prove_table(oracle, traces.T_MEM_AFTER, STANDARD_FAST, tr, BG2, STATE0, forced_pow=w + 1 if w + 1 != 0 else 2)


# ---- the remaining tables: valid traces from restated generators must satisfy the transcribed constraints --------------------
def test_keccak_generator_is_keccak_f():
    """the trace generator's permutation output equals Keccak-f[1600] (checked through hashlib's SHA3-256 of b'')"""
    import hashlib
    block = bytearray(200)
    block[0] = 0x06
    block[135] |= 0x80
    lanes = np.frombuffer(bytes(block), dtype="<u8").reshape(1, 25)
    _, out = traces.keccak_trace(5, lanes)
    assert out[0, :4].astype("<u8").tobytes() == hashlib.sha3_256(b"").digest()


@pytest.mark.parametrize("table,lg,cfg", [
    (traces.T_CPU, 6, TEST_CONFIG), (traces.T_CPU, 9, STANDARD_FAST),
    (traces.T_KECCAK, 5, TEST_CONFIG), (traces.T_KECCAK, 7, STANDARD_FAST),
    (traces.T_BYTE_PACKING, 8, TEST_CONFIG), (traces.T_BYTE_PACKING, 9, STANDARD_FAST),
    (traces.T_ARITHMETIC, 16, TEST_CONFIG),
])
def test_prove_then_verify_more_tables(oracle, table, lg, cfg):
    rng = np.random.default_rng(lg)
    if table == traces.T_CPU:
        tr = traces.cpu_padding_trace(lg)
    elif table == traces.T_KECCAK:
        nperm = (1 << lg) // 24 - (1 if lg == 7 else 0)      # leave padding rows in one case
        tr, _ = traces.keccak_trace(lg, rng.integers(0, 2 ** 63, size=(max(nperm, 1), 25), dtype=np.uint64))
    elif table == traces.T_BYTE_PACKING:
        tr = traces.byte_packing_trace(lg, 5)
    else:
        tr = traces.arithmetic_addcy_trace(lg, 6)
    bg = BG2[:2 * cfg[1]]
    proof, st = orc_prove_table(oracle, table, cfg, tr, bg, STATE0)
    ok, err, st2 = orc_verify_table(oracle, table, cfg, proof, bg, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)


@pytest.mark.parametrize("table", [traces.T_KECCAK, traces.T_BYTE_PACKING])
def test_corrupted_valid_trace_is_rejected(oracle, table):
    rng = np.random.default_rng(1)
    if table == traces.T_KECCAK:
        tr, _ = traces.keccak_trace(5, rng.integers(0, 2 ** 63, size=(1, 25), dtype=np.uint64))
        tr[traces.K_START_A_PP + 3, 5] ^= np.uint64(4)
    else:
        tr = traces.byte_packing_trace(8, 2)
        tr[37 + 31, 0] = 9 if tr[1 + 31, 0] == 0 else tr[37 + 31, 0]     # a byte past the sequence length
        tr[1, 3] = 2                                                      # non-boolean index flag
    proof, _ = orc_prove_table(oracle, table, TEST_CONFIG, tr, BG2[:2], STATE0)
    ok, err, _ = orc_verify_table(oracle, table, TEST_CONFIG, proof, BG2[:2], STATE0)
    assert not ok


# ---- whole segment: prove_with_traces + verify_proof incl. the cross-table-lookup sums ----------------------------------------
PUBLIC_VALUES = np.arange(1000, 1000 + 37, dtype=np.uint64)     # stand-in for the flattened PublicValues


def test_segment_prove_verify_with_ctl_sums(oracle):
    from tests.oracle_lib import orc_prove_segment, orc_verify_segment
    tr = traces.valid_segment(seed=3)
    proofs, bg, caps = orc_prove_segment(oracle, TEST_CONFIG, tr, PUBLIC_VALUES)
    assert [p is not None for p in proofs] == [True, False, True, False, False, False, True, True, True]
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PUBLIC_VALUES)
    assert ok, err
    # the memory <-> mem_before / mem_after lookups really carry something: dropping one MemAfter row breaks the sum check only
    bad = [None if t is None else t.copy() for t in tr]
    bad[traces.T_MEM_AFTER][0, 5] = 0                       # filter off: the row is no longer looked
    proofs2, _, _ = orc_prove_segment(oracle, TEST_CONFIG, bad, PUBLIC_VALUES)
    ok2, err2 = orc_verify_segment(oracle, TEST_CONFIG, proofs2, PUBLIC_VALUES)
    assert not ok2 and "Cross-table lookup 8" in err2, err2
    # other public values -> other challenges -> the old proofs no longer verify
    ok3, _ = orc_verify_segment(oracle, TEST_CONFIG, proofs, PUBLIC_VALUES + np.uint64(1))
    assert not ok3


def test_keccak_blocked_evaluator_equals_reference_order():
    """the device's reordered Keccak evaluator (index-addressed alpha powers) == the reference emission order on arbitrary rows"""
    import ctypes as C
    from tests import oracle_lib
    orc = oracle_lib.load()
    rng = np.random.default_rng(11)
    u64p = C.POINTER(C.c_uint64)
    for trial in range(3):
        lv = oracle_lib.rand_field(rng, (2431,))
        nv = oracle_lib.rand_field(rng, (2431,))
        if trial == 2:   # bit-valued rows as in a real trace
            lv = (lv & np.uint64(1)).astype(np.uint64)
            nv = (nv & np.uint64(1)).astype(np.uint64)
        al = oracle_lib.rand_field(rng, (2,))
        sel = oracle_lib.rand_field(rng, (3,))
        a, b = np.zeros(2, np.uint64), np.zeros(2, np.uint64)
        r = orc.lib.orc_eval_forms_agree(3, lv.ctypes.data_as(u64p), nv.ctypes.data_as(u64p), al.ctypes.data_as(u64p),
                                         sel.ctypes.data_as(u64p), a.ctypes.data_as(u64p), b.ctypes.data_as(u64p))
        assert r == 1, (a, b)
        assert a[0] != 0 or trial == 2


# ---- KeccakSpongeStark on VALID traces (the other sponge cases in this repo are random traces: GPU == oracle only) ------------------------
def _sponge_ops(seed):
    rng = np.random.default_rng(seed)
    return [(1, 2, 100, 7, b""), (0, 3, 5, 9, rng.bytes(135)), (2, 1, 0, 11, rng.bytes(136)), (0, 0, 50, 13, rng.bytes(300)),
            (3, 4, 9, 20, rng.bytes(1)), (0, 2, 7, 21, b"abc")]


def test_keccak_sponge_generator_is_keccak256():
    _, dg = traces.keccak_sponge_trace(8, _sponge_ops(1))
    assert dg[0].hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"      # keccak256("")
    assert dg[5].hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"      # keccak256("abc")


def test_keccak_sponge_test_generation_of_the_reference(oracle):
    """keccak_sponge_stark.rs:993-1021 `test_generation`: the operation it builds (input [1, 2, 3] at (0, Segment::Code, 0), timestamp 0)
    gives ONE row whose updated_digest_state_bytes are keccak256([1, 2, 3]) (the well-known digest, written out here)"""
    tr, dg = traces.keccak_sponge_trace(8, [(0, 0, 0, 0, bytes([1, 2, 3]))])
    assert int(tr[0, 0]) == 0 and [int(v) for v in tr[6:142, 0]].index(1) == 3      # rows.len() == 1: the first row is already the final one (3 input bytes)
    assert bytes(int(v) for v in tr[404:436, 0]).hex() == "f1885eda54b7a053318cd41e2093220dab15d65381b1157a3633a83bfd5c9239"
    assert dg[0].hex() == "f1885eda54b7a053318cd41e2093220dab15d65381b1157a3633a83bfd5c9239"
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_KECCAK_SPONGE, tr) == []


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_valid_keccak_sponge_trace_verifies(oracle, cfg):
    """generated rows => all 706 KeccakSponge constraints + the byte range-check lookup hold (the reference's test_generation pattern,
    keccak_sponge_stark.rs:995): the oracle proves the trace and its verifier accepts"""
    tr, _ = traces.keccak_sponge_trace(8, _sponge_ops(2))
    bg = BG2[:2 * cfg[1]]
    proof, st = orc_prove_table(oracle, traces.T_KECCAK_SPONGE, cfg, tr, bg, STATE0)
    ok, err, st2 = orc_verify_table(oracle, traces.T_KECCAK_SPONGE, cfg, proof, bg, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)


@pytest.mark.parametrize("col,row", [(0, 3), (192 + 7, 2), (404 + 5, 2), (362 + 3, 4), (5, 3), (437, 0)])
def test_corrupted_keccak_sponge_trace_is_rejected(oracle, col, row):
    """rows 2, 4, 5 are full input blocks (their updated state must reappear as the next row's original state); a final row's digest is
    tied to the Keccak table by a cross-table lookup only, so it is not among the cases"""
    tr, _ = traces.keccak_sponge_trace(8, _sponge_ops(3))
    tr[col, row] ^= np.uint64(1)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_KECCAK_SPONGE, TEST_CONFIG, tr, bg, STATE0)
    ok, _, _ = orc_verify_table(oracle, traces.T_KECCAK_SPONGE, TEST_CONFIG, proof, bg, STATE0)
    assert not ok


# ---- ArithmeticStark MUL rows (the addcy generator covers ADD / SUB / LT / GT only) -----------------------------------------------------
def test_valid_arithmetic_mul_shl_byte_trace_verifies_and_corruptions_are_rejected(oracle):
    """MUL (mul.rs), SHL (shift.rs, incl. shifts >= 256) and BYTE (byte.rs, incl. indices >= 32 and with high limbs) rows"""
    tr = traces.arithmetic_mul_trace(16, 3, nops=100)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, tr, bg, STATE0)
    ok, err, _ = orc_verify_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, proof, bg, STATE0)
    assert ok, err
    byte_row = 106 + 7 + 8                  # a BYTE row with a random index / value
    assert tr[13, byte_row] == 1 and tr[1, 7] == 1
    for col, row in ((66, byte_row),):                  # a wrong selected byte, range-check frequencies kept consistent
        t2 = tr.copy()
        t2[col, row] ^= np.uint64(1)
        t2[115, :65536] = np.bincount(t2[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
        proof, _ = orc_prove_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, t2, bg, STATE0)
        assert not orc_verify_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, proof, bg, STATE0)[0], (col, row)


def test_valid_arithmetic_modular_trace_verifies_and_corruptions_are_rejected(oracle):
    """ADDMOD / MULMOD / SUBMOD / the three FP254 operations / DIV / MOD / SHR (two-row operations: modular.rs, divmod.rs, shift.rs), incl.
    modulus / denominator 0 and 1, negative SUBMOD quotients, shifts >= 256, maximal operands"""
    tr = traces.arithmetic_modular_trace(16, 3, nops=40)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, tr, bg, STATE0)
    ok, err, _ = orc_verify_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, proof, bg, STATE0)
    assert ok, err
    # DIV remainder   (each case is a 2^16-row proof; ADDMOD / MOD cells were rejected the same way when this test was written)
    for col, row in ((82, 20),):
        t2 = tr.copy()
        t2[col, row] ^= np.uint64(1)
        t2[115, :65536] = np.bincount(t2[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
        proof, _ = orc_prove_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, t2, bg, STATE0)
        assert not orc_verify_table(oracle, traces.T_ARITHMETIC, TEST_CONFIG, proof, bg, STATE0)[0], (col, row)


# ---- CpuStark with ACTIVE rows (the other valid Cpu traces of this repo are all-padding rows) -------------------------------------------
@pytest.mark.parametrize("program", ["J", "P", "0", "JP0PJ00PPJ", "PN", "PX", "PPXJ", "PZ", "0Z", "PPE", "P0E", "PPA", "PPM", "PPNEZ0PAXNJPPPMXXXJ",
                                     "PPPa", "PPPm", "PPPPPPPPSDOLGPPB&|^PPaPPmXJ",
                                     "Pu", "PPv", "PPs", "PPPt", "P" * 17 + "qyJ", "PPPuvwstNXJ",
                                     "PPSuAuAPAjNJ", "0PPSuAuAPAjNJ", "0PPSuAuAPAiJJ", "PPPSuAuAPAiNJ", "00PPSuAuAPAiJJ", "0PPPSuAuAPAiNJ",
                                     "I", "IIf", "IIg", "IIh", "IIK", "IIE", "INZ", "IIIIAm", "IIIfNKXJ"])
def test_cpu_program_rows_verify(oracle, program):
    """a straight-line kernel program (JUMPDEST, PC, PUSH0, NOT, POP, ISZERO, EQ, the eight binary arithmetic instructions, AND / OR / XOR,
    ADDMOD, MULMOD, DUP1/2/3/16, SWAP1/2/16, JUMP and JUMPI taken and not taken — the destinations are computed on the stack from PC values —, the FP254 operations, KECCAK_GENERAL,
    PROVER_INPUT pushing random 256-bit words)
    running into halt_final: decode, control flow,
    gas, clock, every StackBehavior shape (cached top, partial-channel write of the old top, second-operand and new-top reads, stack_inv*),
    pc.rs, push0.rs, simple_logic/{not,eq_iszero}.rs, halt.rs with operation flags set"""
    tr = traces.cpu_program_trace(6, program)
    bg = BG2[:2]
    proof, st = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    ok, err, st2 = orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)


@pytest.mark.parametrize("what,col,row,delta", [
    ("gas after JUMPDEST", 5, 1, 1), ("pc skips an instruction", 2, 3, 1), ("PC pushes another value", 46, 2, 1), ("PUSH0 pushes non-zero", 46, 3, 1),
    ("old top not written to memory", 80, 3, -1), ("partial channel address", 84, 3, 1), ("opcode bit", 26, 1, 1), ("stack_len after a push", 3, 2, 1),
    ("top changes over JUMPDEST", 46, 5, 1), ("two operation flags", 13, 1, 1), ("clock", 40, 4, 1), ("kernel mode dropped", 4, 2, -1),
    ("stack_inv_aux", 37, 3, -1)])
def test_cpu_program_corruptions_are_rejected(oracle, what, col, row, delta):
    tr = traces.cpu_program_trace(6, "JP0PJ00PPJ")
    tr[col, row] = np.uint64(int(tr[col, row]) + delta)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


PROGRAM2 = "PPNEZ0PAXNJPPPMXXXJ"      # rows: 0 P, 1 P, 2 N, 3 E, 4 Z, 5 0, 6 P, 7 A, 8 X, 9 N, 10 J, 11 P, 12 P, 13 P, 14 M, 15 X, 16 X, 17 X, 18 J


@pytest.mark.parametrize("what,col,row,delta", [
    ("NOT result limb", 46 + 3, 3, 1), ("EQ result", 46, 4, 1), ("EQ diff_pinv", 32, 3, 1), ("EQ second operand address", 58, 3, 1),
    ("EQ second operand not read", 54, 3, -1), ("ISZERO result", 46, 5, 1), ("ADD gas", 5, 8, 1), ("MUL gas", 5, 15, -2),
    ("stack_len after ADD", 3, 8, 1), ("POP new-top read missing", 41, 16, -1), ("POP new-top address", 45, 16, 1), ("stack_inv_aux_2", 38, 15, -1)])
def test_cpu_program_corruptions_of_pops_and_logic_are_rejected(oracle, what, col, row, delta):
    tr = traces.cpu_program_trace(6, PROGRAM2)
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


PROGRAM3 = "PPPPPPPPSDOLGPPB&|^PPaPPmXJ"     # rows 8 S, 9 D, 10 O, 11 L, 12 G, 15 B, 16 &, 17 |, 18 ^, 21 a, 24 m


@pytest.mark.parametrize("what,col,row,delta", [
    ("DIV gas", 5, 10, -2), ("LT gas", 5, 12, 2), ("AND gas", 5, 17, 1), ("ADDMOD gas", 5, 22, -8), ("ADDMOD third operand not read", 67, 21, -1),
    ("MULMOD third operand address", 71, 24, 1), ("stack_len after MULMOD", 3, 25, 1), ("ADDMOD second operand address", 58, 21, 1), ("XOR with the OR flag too", 6, 18, 1)])
def test_cpu_program_corruptions_of_arithmetic_and_logic_rows_are_rejected(oracle, what, col, row, delta):
    """(which of the arithmetic opcodes a binary_op row carries is NOT an in-table constraint: decode.rs leaves it to the cross-table
    lookup with the Arithmetic table, so an opcode-bit flip on such a row is deliberately not among the cases)"""
    tr = traces.cpu_program_trace(6, PROGRAM3)
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


PROGRAM4 = "PPPuvwstNXJ"     # rows 3 u DUP1, 4 v DUP2, 5 w DUP3, 6 s SWAP1, 7 t SWAP2


@pytest.mark.parametrize("what,col,row,delta", [
    ("DUP2 reads another element", 71, 4, 1), ("DUP2 new top differs from the element read", 46, 5, 1), ("DUP1 old top not written", 54, 3, -1),
    ("DUP3 written value differs from the top", 59, 5, 1), ("SWAP1 write address", 71, 6, 1), ("SWAP2 value written differs from the top", 72, 7, 1),
    ("SWAP2 new top differs from the element read", 46, 8, 1), ("stack_len after DUP", 3, 4, 1), ("DUP gas", 5, 4, 1), ("SWAP marked as a read", 68, 6, 1)])
def test_cpu_program_corruptions_of_dup_swap_rows_are_rejected(oracle, what, col, row, delta):
    tr = traces.cpu_program_trace(6, PROGRAM4)
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


PROGRAM5 = "0PPPSuAuAPAiNJ"     # rows 0..10 build [0, cond, dst]; row 11 = JUMPI taken to the J at +13, skipping the N at +12; one element stays


@pytest.mark.parametrize("what,col,row,delta", [
    ("JUMPI lands elsewhere", 2, 12, -1), ("should_jump cleared", 32, 11, -1), ("cond_sum_pinv", 33, 11, 1),
    ("JUMPDEST-bit channel segment", 70, 11, 1), ("JUMPDEST-bit channel address", 71, 11, 1), ("JUMPDEST-bit channel used in kernel mode", 67, 11, 1),
    ("new top not read after the jump", 41, 12, -1), ("stack_len after JUMPI", 3, 12, 1), ("JUMPI gas", 5, 12, -2)])
def test_cpu_program_corruptions_of_jump_rows_are_rejected(oracle, what, col, row, delta):
    """(the `used` flag and address of the channel JUMPI reads its condition through are not constrained in this table by the reference
    either — stack.rs defines JUMPI_OP but STACK_BEHAVIORS.jumps is None and jumps.rs only disables the channel for JUMP — so clearing
    that flag is deliberately not among the cases)"""
    tr = traces.cpu_program_trace(6, PROGRAM5)
    assert tr[14, 11] == 1 and tr[24, 11] == 1          # row 11 is the JUMPI
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


@pytest.mark.parametrize("what,col,row,delta", [
    ("FP254 modulus limb", 72 + 2, 3, 1), ("FP254 operation charged gas", 5, 4, 1), ("PROVER_INPUT charged gas", 5, 1, 3),
    ("KECCAK_GENERAL stack_len", 3, 6, 1), ("NOT of a random word", 46 + 5, 5, 1), ("PROVER_INPUT old top not written", 80, 1, -1)])
def test_cpu_program_corruptions_of_kernel_only_rows_are_rejected(oracle, what, col, row, delta):
    tr = traces.cpu_program_trace(6, "IIIfNKXJ")      # rows 0-2 I, 3 f ADDFP254, 4 N, 5 K, 6 X, 7 J
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


def _addr(ctx, seg, virt):
    return virt | (seg << 32) | (ctx << 64)


MEMIO_PROGRAM, MEMIO_INPUTS = "IIIIrlXJ", [_addr(9, 9, 9), 1, _addr(2, 5, 77), 123]      # rows 0-3 I, 4 r MSTORE_GENERAL, 5 l MLOAD_GENERAL, 6 X, 7 J


@pytest.mark.parametrize("program,inputs", [("Il", [_addr(3, 7, 100)]), ("IIl", [5, _addr(0, 2, 9)]), ("IIIl", [5, 6, _addr(1, 1, 0)]),
                                            ("IIr", [_addr(2, 5, 77), 123]), ("IIIrJ", [9, _addr(2, 5, 77), 123]), (MEMIO_PROGRAM, MEMIO_INPUTS)])
def test_cpu_memio_rows_verify(oracle, program, inputs):
    """MLOAD_GENERAL / MSTORE_GENERAL (memio.rs): the address word's limbs drive the load channel / the partial store channel"""
    tr = traces.cpu_program_trace(6, program, inputs=inputs)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    ok, err, _ = orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)
    assert ok, err


@pytest.mark.parametrize("what,col,row,delta", [
    ("store segment differs from the address word", 83, 4, 1), ("store marked as a read", 81, 4, 1), ("store address word not read", 54, 4, -1),
    ("new top not read after the store", 41, 5, -1), ("load context differs from the address word", 56, 5, 1), ("loaded word is not the new top", 46 + 1, 6, 1),
    ("load through a second channel too", 67, 5, 1), ("stack_inv_aux_2 on a load", 38, 5, 1)])
def test_cpu_memio_corruptions_are_rejected(oracle, what, col, row, delta):
    tr = traces.cpu_program_trace(6, MEMIO_PROGRAM, inputs=MEMIO_INPUTS)
    assert tr[20, 4] == 1 and tr[20, 5] == 1 and tr[24, 5] == 1
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


@pytest.mark.parametrize("program,inputs", [("II<", [0xABCDEF, 5]), ("II>", [1 << 200, 199]), ("II<", [7, 256]), ("II>", [7, 1 << 40]),
                                            ("II<", [7, (1 << 255) + 3]), ("III<>XJ", [3, 1 << 100, 64])])
def test_cpu_shift_rows_verify(oracle, program, inputs):
    """SHL / SHR (cpu/shift.rs): 2^d read from the kernel's shift table, not read when the displacement has high limbs"""
    tr = traces.cpu_program_trace(6, program, inputs=inputs)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    ok, err, _ = orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)
    assert ok, err


@pytest.mark.parametrize("what,inputs,col,delta", [
    ("shift-table address", [7, 5], 71, 1), ("shift-table segment", [7, 5], 70, 1), ("shift table not read", [7, 5], 67, -1),
    ("table read although the displacement has high limbs", [7, 1 << 40], 67, 1), ("high_limb_sum_inv", [7, 1 << 40], 32, 1), ("shift gas", [7, 5], 5, 1)])
def test_cpu_shift_corruptions_are_rejected(oracle, what, inputs, col, delta):
    tr = traces.cpu_program_trace(6, "II<J", inputs=inputs)
    row = 3 if col == 5 else 2
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


@pytest.mark.parametrize("program", ["C", "IC", "IICZXJ"])
def test_cpu_get_context_rows_verify(oracle, program):
    tr = traces.cpu_program_trace(6, program)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    ok, err, _ = orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)
    assert ok, err


@pytest.mark.parametrize("what,col,row,delta", [("old top not written", 67, 2, -1), ("write address", 71, 2, 1), ("pushed context", 46 + 2, 3, 1),
                                                ("stack_len after GET_CONTEXT", 3, 3, 1), ("write marked as a read", 68, 2, 1)])
def test_cpu_get_context_corruptions_are_rejected(oracle, what, col, row, delta):
    tr = traces.cpu_program_trace(6, "IICZXJ")          # row 2 = GET_CONTEXT (contextops.rs)
    tr[col, row] = np.uint64((int(tr[col, row]) + delta) % traces.P)
    bg = BG2[:2]
    proof, _ = orc_prove_table(oracle, traces.T_CPU, TEST_CONFIG, tr, bg, STATE0)
    assert not orc_verify_table(oracle, traces.T_CPU, TEST_CONFIG, proof, bg, STATE0)[0], what


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_cpu_demo_program_verifies(oracle, cfg):
    """72 executed rows mixing every instruction family cpu_program_trace knows (the GPU parity test proves the same trace)"""
    program, inputs = traces.cpu_demo_program()
    tr = traces.cpu_program_trace(7, program, inputs=inputs)
    assert int((tr[6:24].sum(axis=0) > 0).sum()) == 72
    bg = BG2[:2 * cfg[1]]
    proof, st = orc_prove_table(oracle, traces.T_CPU, cfg, tr, bg, STATE0)
    ok, err, st2 = orc_verify_table(oracle, traces.T_CPU, cfg, proof, bg, STATE0)
    assert ok, err
    assert np.array_equal(st, st2)


def test_arithmetic_basic_trace_of_the_reference(oracle):
    """arithmetic_stark.rs:374-451 `basic_trace`, the reference's own expected values: the ten operations it generates, the rows it
    expects them on (two-row operations included) and the single-word answers it reads from OUTPUT_REGISTER — reproduced by the restated
    generator; and the transcribed constraints vanish on every row of that trace."""
    ops = [("add", 123, 456), ("mulmod", 123, 456, 1007), ("addmod", 1234, 567, 1007), ("mul", 123, 456), ("mod", 128, 13),
           ("lt", 128, 13), ("lt", 13, 128), ("lt", 128, 128), ("div", 128, 13), ("byte", 30, 0xABCD)]
    t, first = traces.arithmetic_trace_from_operations(ops)
    assert t.shape == (116, 1 << 16)                                     # NUM_ARITH_COLUMNS x RANGE_MAX
    expected_output = [(0, 579), (1, 703), (3, 794), (5, 56088), (6, 11), (8, 0), (9, 1), (10, 0), (11, 9), (13, 0xAB)]
    assert first == [row for row, _ in expected_output]
    for row, expected in expected_output:
        assert int(t[66, row]) == expected                               # OUTPUT_REGISTER.start
        assert not t[67:82, row].any()                                   # "...other registers should be zero"
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_ARITHMETIC, t) == []


# ---- a VALID multi-table segment around an executing Cpu program: the cross-table lookups Cpu -> Memory / Arithmetic / Logic ----------
PV37 = np.arange(1000, 1037, dtype=np.uint64)


def _segment_verifies(oracle, program, cfg=TEST_CONFIG, **kw):
    tr, labels = traces.cpu_segment(program, **kw)
    proofs, _, _ = orc_prove_segment(oracle, cfg, tr, PV37, labels=labels)
    return orc_verify_segment(oracle, cfg, proofs, PV37, labels=labels)


@pytest.mark.parametrize("program", ["PP|PP^PPaXXJ", "0PPPSuAuAPAiNJ"])       # (each case proves a 2^16-row Arithmetic table)
def test_cpu_segment_cross_table_lookups_verify(oracle, program):
    """every memory operation the Cpu rows send (opcode fetch, general-purpose channels, partial channel; timestamps clock * 5 + channel - 4)
    is found by the Memory table, every arithmetic / logic instruction by the Arithmetic / Logic tables, MemBefore / MemAfter close the
    memory: verify_cross_table_lookups accepts all ten lookups (the generator follows the reference's definitions, cpu_stark.rs:324-379)"""
    ok, err = _segment_verifies(oracle, program)
    assert ok, err


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_cpu_segment_long_program_verifies(oracle, cfg):
    ok, err = _segment_verifies(oracle, traces.CPU_SEGMENT_PROGRAM, cfg)
    assert ok, err


def test_memory_lookup_with_the_public_value_writes_needs_the_extra_looking_sum(oracle):
    """verify_proof (verifier.rs:172-313) adds get_memory_extra_looking_sum(public_values, challenge) (:319-512) to the looking side of the
    Memory lookup: the kernel writes the block metadata, trie roots, bloom filter, 256 block hashes and the registers to memory at
    timestamp 2 without a Cpu row sending them.  A segment whose Memory table holds those 301 writes verifies WITH the sum computed from
    the PublicValues and the segment's (beta, gamma), and fails on lookup 6 without it / with the sum of other public values."""
    from zk_evm_b200.public_values import (PublicValues, RegistersData, flatten_public_values, memory_extra_looking_values,
                                           memory_extra_looking_sum)
    rng = np.random.default_rng(11)
    pv = PublicValues()
    pv.block_metadata.block_beneficiary = rng.bytes(20)
    pv.block_metadata.block_timestamp, pv.block_metadata.block_number, pv.block_metadata.block_gas_used = 1700000000, 19807080, 12345678
    pv.block_metadata.block_random, pv.block_metadata.block_base_fee = rng.bytes(32), 9_000_000_007
    pv.block_metadata.block_bloom = [int.from_bytes(rng.bytes(32), "big") for _ in range(8)]
    pv.block_hashes.prev_hashes = [rng.bytes(32) for _ in range(256)]
    pv.block_hashes.cur_hash = rng.bytes(32)
    pv.trie_roots_before.state_root, pv.trie_roots_after.receipts_root = rng.bytes(32), rng.bytes(32)
    pv.extra_block_data.txn_number_after, pv.extra_block_data.gas_used_after = 7, 654321
    pv.registers_before = RegistersData(program_counter=3, is_kernel=1, stack_len=0, stack_top=0, context=0, gas_used=0)
    pv.registers_after = RegistersData(program_counter=0x1234, is_kernel=1, stack_len=2, stack_top=1 << 200, context=0, gas_used=77)
    kernel_hash, kernel_len = rng.bytes(32), 60123
    rows = memory_extra_looking_values(pv, kernel_hash, kernel_len)
    assert len(rows) == 25 + 8 + 256 + 12 and all(len(r) == 13 and r[0] == 0 and r[1] == 0 and r[12] == 2 for r in rows)
    assert len({(r[2], r[3]) for r in rows}) == len(rows)                     # one write per address
    flat = flatten_public_values(pv)
    cfg = TEST_CONFIG
    tr, labels = traces.cpu_segment("PPMXJ", log_mem=10, extra_memory_rows=rows)
    proofs, bg, _ = orc_prove_segment(oracle, cfg, tr, flat, labels=labels)
    nc = cfg[1]
    sums = np.zeros(10 * nc, dtype=np.uint64)
    for c in range(nc):
        sums[6 * nc + c] = memory_extra_looking_sum(pv, int(bg[2 * c]), int(bg[2 * c + 1]), kernel_hash, kernel_len)
    ok, err = orc_verify_segment(oracle, cfg, proofs, flat, labels=labels, extra_looking_sums=sums)
    assert ok, err
    ok, err = orc_verify_segment(oracle, cfg, proofs, flat, labels=labels)
    assert not ok and "lookup 6" in err
    pv.block_metadata.block_number += 1
    for c in range(nc):
        sums[6 * nc + c] = memory_extra_looking_sum(pv, int(bg[2 * c]), int(bg[2 * c + 1]), kernel_hash, kernel_len)
    ok, err = orc_verify_segment(oracle, cfg, proofs, flat, labels=labels, extra_looking_sums=sums)
    assert not ok and "lookup 6" in err


def test_cpu_segment_with_four_channel_timestamps_is_rejected(oracle):
    """the check that found the NUM_CHANNELS transcription error: memory timestamps computed with 4 channels do not match the lookups"""
    ok, err = _segment_verifies(oracle, "PPMXJ", num_channels=4)
    assert not ok and "lookup 6" in err


def test_keccak_sponge_lookups_into_keccak_logic_and_memory_balance(oracle):
    """KeccakSponge operations next to the executing Cpu program: the sponge rows' permutations are found in the Keccak table (lookups 3 and
    4: inputs and outputs), their xors in the Logic table (5), their input bytes in Memory (6).  No Cpu row asked for these operations
    (a KECCAK_GENERAL row needs PROVER_INPUT pushes, which look up Arithmetic range-check rows this generator does not build), so lookup 2
    (Cpu -> KeccakSponge) is the ONLY one that does not balance — the oracle's verifier lists every failing lookup."""
    rng = np.random.default_rng(3)
    sponge_ops = [(1, 0, 10, 7, rng.bytes(5)), (1, 0, 40, 9, rng.bytes(140)), (1, 0, 200, 11, b"")]
    tr, labels = traces.cpu_segment("PPMXJ", sponge_ops=sponge_ops, log_mem=10)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok and err.endswith("failing lookups: 2"), err
    # and the lookups do depend on the sponge's data: another timestamp in the Keccak table unbalances 3 and 4 as well
    tr[traces.T_KECCAK][24, :24] += np.uint64(1)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok and err.endswith("failing lookups: 2 3 4"), err


def test_byte_packing_lookups_into_memory_balance(oracle):
    """BytePacking operations (three unpacking writes, one packing read of a pre-initialised segment) next to the executing Cpu program:
    their bytes are found in Memory at virt + len - 1 - i (lookup 6); only Cpu -> BytePacking (1), which nobody asked for, stays open"""
    rng = np.random.default_rng(4)
    packing_ops = [(0, 2, 3, 0, 5, rng.bytes(32)), (0, 2, 3, 40, 6, rng.bytes(1)), (1, 2, 0, 10, 7, rng.bytes(17)), (0, 2, 3, 80, 9, rng.bytes(5))]
    tr, labels = traces.cpu_segment("PPMXJ", packing_ops=packing_ops, log_mem=10)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok and err.endswith("failing lookups: 1"), err


def test_cpu_segment_with_keccak_general_and_prover_input_verifies(oracle):
    """eight tables in use: the program pushes (length, address) pairs with PROVER_INPUT (range-checked by Arithmetic rows, IS_RANGE_CHECK)
    and hashes two memory ranges with KECCAK_GENERAL; the KeccakSponge table holds exactly the operations the Cpu rows ask for (timestamp
    (clock - 1) * 5 + 1, digest as the pushed word), Keccak their permutations, Logic their xors, Memory their byte reads: all ten lookups
    balance, incl. Cpu -> KeccakSponge (2), which the hand-placed sponge operations of the test above leave open"""
    rng = np.random.default_rng(5)
    addr = lambda c, sg, v: v | (sg << 32) | (c << 64)
    data = {(1, 0, 10): rng.bytes(150), (1, 0, 300): b"abc"}
    tr, labels = traces.cpu_segment("IIKIIKXXJ", inputs=[150, addr(1, 0, 10), 3, addr(1, 0, 300)], keccak_inputs=data, log_mem=10)
    assert [t is not None for t in tr] == [True, False, True, True, True, True, True, True, True]
    # the second digest is keccak256("abc"), pushed as a big-endian word: row 6 holds it as the cached top of the stack
    top = sum(int(tr[traces.T_CPU][46 + l, 6]) << (32 * l) for l in range(8))
    assert top == 0x4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert ok, err


# ---- 32-byte memory operations: MLOAD_32BYTES / MSTORE_32BYTES rows, byte_unpacking.rs, and the Cpu -> BytePacking lookup with non-zero sums ----
PACK_PROGRAM = "IIR" "IIW" "IIV" "IIT" "IIR" "XXXXXJ"
_ADDR = lambda c, sg, v: v | (sg << 32) | (c << 64)
PACK_INPUTS = [7, _ADDR(1, 0, 100), 0x1234567890abcdef << 64, _ADDR(0, 2, 50), 0xaabbccddee, _ADDR(0, 2, 200), 0x7f, _ADDR(3, 4, 0), 32, _ADDR(1, 0, 95)]


def test_cpu_rows_of_32_byte_memory_operations_satisfy_every_constraint(oracle):
    """the reference's generator-test pattern (all constraints vanish on consecutive trace rows) on a program with MLOAD_32BYTES and
    MSTORE_32BYTES_32 / _5 / _1 rows: decode.rs:202-211, stack.rs (two pops, one push), byte_unpacking.rs:11-44 with its filter on"""
    t = traces.cpu_program_trace(6, PACK_PROGRAM, inputs=PACK_INPUTS)
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, t) == []
    # MSTORE_32BYTES_32 is row 5: the address word it pushes (row 6's cached top) must be the old one + 32 in its virtual limb only
    for col, what in ((46, "virt advanced by len"), (47, "segment kept"), (48, "context kept"), (49, "upper limbs zero")):
        bad = t.copy()
        bad[col, 6] += np.uint64(1)
        hits = oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad)
        assert hits and hits[0][0] == 5 and hits[0][1] < 8, what      # the byte_unpacking constraints are the first eight of the dispatcher
    # an MSTORE_32BYTES opcode whose length bits disagree with the advance
    bad = t.copy()
    bad[24, 5] = 0                                                   # MSTORE_32BYTES_31 instead of _32
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad)
    # the flag on a row whose opcode is neither 0xf8 nor 0xc0..0xdf
    bad = t.copy()
    bad[18, 0], bad[15, 0] = 1, 0
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad)


@pytest.mark.parametrize("cfg", [TEST_CONFIG, STANDARD_FAST])
def test_segment_with_32_byte_memory_operations_verifies(oracle, cfg):
    """Cpu -> BytePacking (lookup 1) with NON-ZERO sums: two MLOAD_32BYTES (7 and 32 bytes of a code segment, overlapping) and three
    MSTORE_32BYTES rows ask the BytePacking table for exactly the operations it holds (is_read, address, length, timestamp
    (clock - 1) * 5 + 1, packed value: cpu_stark.rs:150-223), whose bytes Memory holds (lookup 6): all ten lookups balance"""
    tr, labels = traces.cpu_segment(PACK_PROGRAM, inputs=PACK_INPUTS, log_mem=10)
    assert tr[traces.T_BYTE_PACKING] is not None and int(tr[traces.T_BYTE_PACKING][1:33].sum()) == 5
    for t in (traces.T_CPU, traces.T_BYTE_PACKING, traces.T_MEMORY, traces.T_ARITHMETIC):
        assert oracle_lib.orc_check_table_rows(oracle, t, tr[t], labels=labels) == []
    proofs, _, _ = orc_prove_segment(oracle, cfg, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, cfg, proofs, PV37, labels=labels)
    assert ok, err


@pytest.mark.parametrize("what", ["byte", "length", "timestamp", "is_read", "pushed"])
def test_segment_with_tampered_32_byte_operation_is_rejected(oracle, what):
    """each field the Cpu row sends to BytePacking is bound by lookup 1"""
    tr, labels = traces.cpu_segment(PACK_PROGRAM, inputs=PACK_INPUTS, log_mem=10)
    bp, cpu = tr[traces.T_BYTE_PACKING], tr[traces.T_CPU]
    if what == "byte":                 # a byte of the first operation in the BytePacking table (and consistently in Memory would still differ from the Cpu's word)
        bp[37, 0] ^= np.uint64(1)
    elif what == "length":             # the 7-byte load recorded as 8 bytes
        bp[7, 0], bp[8, 0] = 0, 1
    elif what == "timestamp":
        bp[36, 1] += np.uint64(5)
    elif what == "is_read":
        bp[0, 1] = 1
    else:                              # the word MLOAD_32BYTES pushed (row 3's cached top)
        cpu[46, 3] ^= np.uint64(2)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok, what


# ---- syscall, exception and EXIT_KERNEL rows (syscalls_exceptions.rs:23-134, jumps.rs:13-65) ----------------------------------------
SYS_PROGRAM = "P" "YNJ" "X" "xNNJ" "X" "X" "I" "eNNJ" "J"


def _sys_inputs(halt_final):
    base = halt_final - len(SYS_PROGRAM)
    return [(base + SYS_PROGRAM.index("e") + 3) | (1 << 32) | (77 << 192)]      # kexit_info: continue at the J after "eNN", kernel mode, gas 77


def test_cpu_rows_of_syscall_exception_and_exit_kernel_satisfy_every_constraint(oracle):
    t = traces.cpu_program_trace(6, SYS_PROGRAM, inputs=_sys_inputs(0x1234))
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, t) == []
    rows = {c: SYS_PROGRAM.index(c) - (0 if c == "Y" else 1 if c == "x" else 3) for c in "Yxe"}   # executed-row index: skipped instructions do not run
    ry, rx, re = 1, 4, 9
    assert [int(t[22, ry]), int(t[23, rx]), int(t[19, re])] == [1, 1, 1], rows
    # after the syscall: kernel mode, gas 0, pc = handler, kexit_info on top = (pc + 1, kernel flag, 0.., gas, 0)
    assert int(t[5, ry + 1]) == 0 and int(t[47, ry + 1]) == 1 and int(t[52, ry + 1]) == int(t[5, ry]) and int(t[46, ry + 1]) == int(t[2, ry]) + 1
    # the exception records pc, not pc + 1
    assert int(t[46, rx + 1]) == int(t[2, rx])
    # EXIT_KERNEL restores the gas of its kexit_info
    assert int(t[5, re + 1]) == 77
    for (row, col, what) in ((ry + 1, 2, "pc after the syscall is the handler"), (ry + 1, 5, "gas after the syscall is 0"), (ry + 1, 46, "kexit_info.pc"),
                             (ry + 1, 52, "kexit_info.gas"), (ry, 58, "jump-table address"), (rx, 58, "exception jump-table address"),
                             (rx, 32, "exception code bit"), (re + 1, 2, "pc after EXIT_KERNEL"), (re + 1, 5, "gas after EXIT_KERNEL"),
                             (ry, 54, "the jump-table channel is not a bus access")):
        bad = t.copy()
        bad[col, row] += np.uint64(1)
        assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad), what


def test_segment_with_syscall_exception_and_exit_kernel_verifies(oracle):
    """the handler addresses are read from the kernel's jump tables through BytePacking (Cpu -> BytePacking's jumptable entries,
    cpu_stark.rs:225-262, carry non-zero sums), the pushed kexit_info words are range-checked by Arithmetic rows (operation.rs:777-790)"""
    hf = len(SYS_PROGRAM) + 8
    tr, labels = traces.cpu_segment(SYS_PROGRAM, inputs=_sys_inputs(hf), log_mem=15)
    assert labels[0] == hf and int(tr[traces.T_BYTE_PACKING][3].sum()) == 2                  # two 3-byte reads
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert ok, err
    # the handler the Cpu row jumps to is bound to the jump-table bytes: another handler byte in BytePacking (consistently in Memory) breaks lookup 1
    bad = [None if t is None else t.copy() for t in tr]
    bad[traces.T_BYTE_PACKING][37, 0] ^= np.uint64(1)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, bad, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok
    # a proof made for other jump-table labels does not verify against these
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=(labels[0], labels[1], labels[2] + 3, labels[3]))
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok


# ---- SET_CONTEXT and context pruning (contextops.rs:150-215, cpu_stark.rs:391-431, all_stark.rs ctl_context_pruning) ---------------------
CTX_PROGRAM = "PP" "Ic" "PCP" "Ic" "XXJ"          # context 0 -> fresh context 5 -> back to 0, pruning 5
CTX_INPUTS = [5 << 64, (0 << 64) | 1]


def test_cpu_rows_of_set_context_satisfy_every_constraint(oracle):
    stale = []
    t = traces.cpu_program_trace(6, CTX_PROGRAM, inputs=CTX_INPUTS, stale=stale)
    assert stale == [5] and [int(x) for x in t[0, :11]] == [0, 0, 0, 0, 5, 5, 5, 5, 5, 0, 0]
    assert [int(x) for x in t[3, :11]] == [0, 1, 2, 3, 0, 1, 2, 3, 4, 2, 1]        # every context resumes with the stack it was left with
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, t) == []
    # GET_CONTEXT in context 5 pushed 5 << 64 (row 6 -> cached top of row 7)
    assert int(t[48, 6]) == 5 and int(t[46, 6]) == 0
    for row, col, what in ((4, 0, "the context after SET_CONTEXT is the popped one"), (8, 32, "pruning flag = low limb of the popped word"),
                           (9, 46, "the top handed over by channel 2 is the next row's cached top"), (8, 71, "new top read at new_sp - 1"),
                           (8, 69, "new top read in the new context"), (3, 38, "stack_inv_aux_2"), (6, 48, "GET_CONTEXT pushes the context")):
        bad = t.copy()
        bad[col, row] += np.uint64(1)
        assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad), what


def test_segment_with_set_context_and_pruning_verifies(oracle):
    """the stack pointers saved in / restored from the ContextMetadata::StackSize cells are Memory operations that only the SET_CONTEXT
    lookups send (cpu_stark.rs:391-431), and the pruned context is one the Memory table lists as stale (context pruning lookup with a
    non-zero sum; its cells do not reach MemAfter)"""
    tr, labels = traces.cpu_segment(CTX_PROGRAM, inputs=CTX_INPUTS, log_mem=10)
    m = tr[traces.T_MEMORY]
    assert int(m[22].sum()) == 1 and int(m[21, 5]) == 6 and int(m[24].sum()) > 0       # one stale context: 5
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert ok, err
    # the Cpu prunes a context the Memory table does not list (and the other way round): the pruning lookup fails
    unlisted = [None if t is None else t.copy() for t in tr]
    unlisted[traces.T_CPU][32, 8] = 0
    unlisted[traces.T_CPU][46, 8] = 0                       # keep the Cpu table itself consistent: flag = low limb of the popped word
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, unlisted[traces.T_CPU], labels=labels) == []
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, unlisted, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    # lookup 9 is the context pruning one; 0 fails too: the changed word is also what PROVER_INPUT's range-check row recorded
    assert not ok and err.endswith("failing lookups: 0 9"), err
    # without pruning the same program verifies with an empty stale list
    tr0, labels0 = traces.cpu_segment(CTX_PROGRAM, inputs=[5 << 64, 0], log_mem=10)
    assert int(tr0[traces.T_MEMORY][22].sum()) == 0
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr0, PV37, labels=labels0)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels0)
    assert ok, err
    # a restored stack pointer that is not the saved one
    wrong = [None if t is None else t.copy() for t in tr]
    wrong[traces.T_CPU][3, 9:] += np.uint64(1)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, wrong, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok


# ---- a user-mode excursion: EXIT_KERNEL to user mode, PUSH read through BytePacking, syscall back (cpu_stark.rs:264-304) -----------------
USER_PROGRAM = "P" "I" "e" "p.." "X" "Y" "N" "J" "X" "X" "J"


def _user_inputs(halt_final):
    base = halt_final - len(USER_PROGRAM)
    return [(base + USER_PROGRAM.index("p")) | (0 << 32) | (1000 << 192)]       # kexit_info: pc of the PUSH2, USER mode, gas 1000


def test_cpu_rows_of_a_user_mode_excursion_satisfy_every_constraint(oracle):
    t = traces.cpu_program_trace(6, USER_PROGRAM, inputs=_user_inputs(0x1234))
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, t) == []
    assert [int(x) for x in t[4, :8]] == [1, 1, 1, 0, 0, 0, 1, 1]               # kernel, kernel, EXIT_KERNEL | PUSH2, POP, syscall | handler ...
    assert int(t[32, 3]) == 1 and int(t[5, 4]) - int(t[5, 3]) == 3              # is_not_kernel on the user-mode PUSH; it costs G_VERYLOW
    assert int(t[2, 4]) - int(t[2, 3]) == 3                                      # and skips its two immediate bytes
    assert int(t[47, 6]) == 0 and int(t[52, 6]) == 1005                          # the syscall's kexit_info: user mode, the gas used so far
    for row, col, what in ((3, 32, "is_not_kernel = 1 - is_kernel_mode"), (3, 39, "stack bound check of a user-mode push"), (2, 39, "... and of EXIT_KERNEL"),
                           (3, 4, "kernel flag after EXIT_KERNEL comes from kexit_info"), (3, 1, "code context = context in user mode"), (6, 4, "a syscall enters kernel mode"),
                           (6, 47, "kexit_info records the mode the syscall came from")):
        bad = t.copy()
        bad[col, row] += np.uint64(1)
        assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad), what


def test_segment_with_a_user_mode_push_verifies(oracle):
    """the last Cpu -> BytePacking entry with a non-zero sum: the word a user-mode PUSH2 pushes is the big-endian packing of the two code
    bytes after it, read by the BytePacking table at (code context, Segment::Code, pc + 1)"""
    hf = len(USER_PROGRAM) + 8
    tr, labels = traces.cpu_segment(USER_PROGRAM, inputs=_user_inputs(hf), log_mem=15)
    bp = tr[traces.T_BYTE_PACKING]
    assert int(bp[2].sum()) == 1 and int(bp[3].sum()) == 1                       # one 2-byte read (the PUSH), one 3-byte read (the jump table)
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, tr, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert ok, err
    # the pushed word is bound to the code bytes: change it on the Cpu side only (row 4's cached top; POP discards it, nothing else sees it)
    bad = [None if t is None else t.copy() for t in tr]
    bad[traces.T_CPU][46, 4] ^= np.uint64(1)
    assert oracle_lib.orc_check_table_rows(oracle, traces.T_CPU, bad[traces.T_CPU], labels=labels) == []
    proofs, _, _ = orc_prove_segment(oracle, TEST_CONFIG, bad, PV37, labels=labels)
    ok, err = orc_verify_segment(oracle, TEST_CONFIG, proofs, PV37, labels=labels)
    assert not ok and err.endswith("failing lookups: 1"), err
