"""starky's debug `check_ctls` (cross_table_lookup.rs debug_utils), which the reference runs over the raw traces of every segment when
debug assertions are on (prover.rs:165-184, ci.yml:93), restated in the oracle (orc_check_ctls): for each of the ten cross-table
lookups the multiset of rows sent must equal the multiset held.  Unlike the proof's running sums it needs no challenges and no proving,
and it names the offending lookup of a broken trace generator directly.  Here it runs over every executing segment the STARK tests
prove, and its verdicts are compared with the lookups the restated verifier reports as failing for the same tampering."""
import numpy as np
import pytest
from tests import oracle_lib, traces
from tests.test_oracle_stark import (PACK_PROGRAM, PACK_INPUTS, SYS_PROGRAM, _sys_inputs, CTX_PROGRAM, CTX_INPUTS, USER_PROGRAM, _user_inputs)

ZERO = [0] * 10


def _segments():
    rng = np.random.default_rng(5)
    addr = lambda c, sg, v: v | (sg << 32) | (c << 64)
    yield "long program", traces.cpu_segment(traces.CPU_SEGMENT_PROGRAM)
    yield "keccak_general", traces.cpu_segment("IIKIIKXXJ", inputs=[150, addr(1, 0, 10), 3, addr(1, 0, 300)],
                                               keccak_inputs={(1, 0, 10): rng.bytes(150), (1, 0, 300): b"abc"}, log_mem=10)
    yield "32-byte operations", traces.cpu_segment(PACK_PROGRAM, inputs=PACK_INPUTS, log_mem=10)
    yield "syscall / exception / exit_kernel", traces.cpu_segment(SYS_PROGRAM, inputs=_sys_inputs(len(SYS_PROGRAM) + 8), log_mem=15)
    yield "set_context + pruning", traces.cpu_segment(CTX_PROGRAM, inputs=CTX_INPUTS, log_mem=10)
    yield "user-mode push", traces.cpu_segment(USER_PROGRAM, inputs=_user_inputs(len(USER_PROGRAM) + 8), log_mem=15)


def test_every_executing_segment_passes_check_ctls(oracle):
    for name, (tr, _) in _segments():
        assert oracle_lib.orc_check_ctls(oracle, tr) == ZERO, name
    tr = traces.valid_segment(seed=3)
    assert oracle_lib.orc_check_ctls(oracle, tr[0] if isinstance(tr, tuple) else tr) == ZERO


def test_check_ctls_names_the_lookup_the_verifier_rejects(oracle):
    # memory timestamps computed with 4 channels (the NUM_CHANNELS transcription error of round 1): the Memory lookup, 6
    tr, _ = traces.cpu_segment("PPMXJ", num_channels=4)
    bad = oracle_lib.orc_check_ctls(oracle, tr)
    assert bad[6] > 0 and sum(bad) == bad[6]
    # a BytePacking operation recorded with another timestamp: Cpu -> BytePacking, 1, and its bytes in Memory, 6
    tr, _ = traces.cpu_segment(PACK_PROGRAM, inputs=PACK_INPUTS, log_mem=10)
    tr[traces.T_BYTE_PACKING][36, 1] += np.uint64(5)
    bad = oracle_lib.orc_check_ctls(oracle, tr)
    assert bad[1] == 2 and bad[6] > 0 and sum(bad) == bad[1] + bad[6]           # the row sent and the row held are both unmatched
    # the Cpu prunes a context the Memory table does not list: the verifier reports lookups 0 and 9 (test_oracle_stark), so does this
    tr, _ = traces.cpu_segment(CTX_PROGRAM, inputs=CTX_INPUTS, log_mem=10)
    tr[traces.T_CPU][32, 8] = 0
    tr[traces.T_CPU][46, 8] = 0
    bad = oracle_lib.orc_check_ctls(oracle, tr)
    assert [i for i, b in enumerate(bad) if b] == [0, 9]
    # an operation the Logic table does not hold
    tr, _ = traces.cpu_segment("PP|PP^PPaXXJ")
    tr[traces.T_LOGIC][3, 0] ^= np.uint64(1)             # an input bit of the first operation
    bad = oracle_lib.orc_check_ctls(oracle, tr)
    assert [i for i, b in enumerate(bad) if b] == [5]


def test_check_ctls_with_the_public_value_writes(oracle):
    from zk_evm_b200.public_values import PublicValues, memory_extra_looking_values
    pv = PublicValues()
    pv.block_metadata.block_number = 19807080
    rows = memory_extra_looking_values(pv, bytes(range(32)), 777)
    tr, _ = traces.cpu_segment("PPMXJ", log_mem=10, extra_memory_rows=rows)
    assert oracle_lib.orc_check_ctls(oracle, tr, rows) == ZERO
    bad = oracle_lib.orc_check_ctls(oracle, tr)
    assert bad[6] == len(rows) and sum(bad) == bad[6]


def test_check_ctls_rejects_a_non_binary_filter(oracle):
    tr, _ = traces.cpu_segment("PPMXJ")
    tr[traces.T_MEMORY][0, 3] = 2                        # the Memory table's filter column
    with pytest.raises(RuntimeError, match="Non-binary filter"):
        oracle_lib.orc_check_ctls(oracle, tr)
