"""The reference's per-table constraint sanity test (first row): starky 1.0.0 `test_stark_low_degree`, called from
cpu_stark.rs:679-703, memory_stark.rs:902-925, logic.rs:400-424, keccak_stark.rs:631-655, keccak_sponge_stark.rs:973-993,
byte_packing_stark.rs:453-476, memory_continuation_stark.rs:160-180 and arithmetic_stark.rs:334+ — on the single-source constraint
templates (csrc/stark/table_*.h) that both the oracle and the CUDA quotient kernels are compiled from.  A transcription slip that
multiplies one column too many shows up here as a degree above 3 n - 1 without needing a valid trace."""
import pytest
from tests import oracle_lib, traces

WITNESS_LOG = 5     # starky's WITNESS_SIZE = 1 << 5
# what the transcription gives today (regression guard; the reference asserts the bound only): degree-3 terms under a transition
# selector reach 3 (n - 1) + 1, Logic has no transition constraints, BytePacking's and the continuation tables' terms are quadratic
EXPECTED = {traces.T_ARITHMETIC: 94, traces.T_BYTE_PACKING: 63, traces.T_CPU: 94, traces.T_KECCAK: 94, traces.T_KECCAK_SPONGE: 94,
            traces.T_LOGIC: 93, traces.T_MEMORY: 94, traces.T_MEM_BEFORE: 62, traces.T_MEM_AFTER: 62}


@pytest.mark.parametrize("table", sorted(EXPECTED))
@pytest.mark.parametrize("seed", [1, 2])
def test_stark_low_degree(oracle, table, seed):
    n = 1 << WITNESS_LOG
    d = oracle_lib.orc_table_constraint_degree(oracle, table, WITNESS_LOG, seed)
    assert d != -2, "a random trace satisfied every constraint"
    assert d <= 3 * n - 1                     # the reference's assertion: constraint_degree() = 3 for every table
    assert d == EXPECTED[table]


@pytest.mark.parametrize("table", [traces.T_CPU, traces.T_MEMORY, traces.T_KECCAK])
def test_stark_low_degree_other_sizes(oracle, table):
    for lg in (3, 7):
        d = oracle_lib.orc_table_constraint_degree(oracle, table, lg, 3)
        assert (1 << lg) <= d <= 3 * (1 << lg) - 1


@pytest.mark.parametrize("table", sorted(EXPECTED))
def test_base_and_extension_evaluations_agree(oracle, table):
    """the analogue of starky's `test_stark_circuit_constraints` (same call sites as above) for this path's two evaluation types: the
    prover evaluates the templates over the base field, the verifier over the quadratic extension; they agree on base-field frames and
    the extension evaluation commutes with conjugation (the constraints have base-field coefficients)"""
    for seed in (1, 2, 3):
        assert oracle_lib.orc_table_eval_consistency(oracle, table, seed) == 0
