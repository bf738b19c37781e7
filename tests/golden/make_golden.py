#!/usr/bin/env python3
"""Regenerates tests/golden/golden_r1.json.

Two kinds of entries:
  "reference_kats"  constants copied from the reference's own sources/tests (file:line given) — the pins of the oracle;
  "oracle_vectors"  outputs of the CPU oracle on seeded inputs (caps, aux/quotient fingerprints, proof fingerprints): the
                    reference has no golden vectors for these stages and cannot be run here (no Rust toolchain), so these are
                    regression vectors of the restatement, NOT reference outputs.  They let the GPU tests check the device
                    against committed numbers without the oracle in the loop, and detect silent drift of the oracle itself.
Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from tests import oracle_lib, traces  # noqa: E402
from tests.oracle_lib import STANDARD_FAST, TEST_CONFIG, orc_prove_table, orc_prove_segment  # noqa: E402

BG2 = np.array([0x1111111111, 0x2222222222, 0x3333333333, 0x4444444444], dtype=np.uint64)
STATE0 = np.arange(1, 13, dtype=np.uint64)
PUBLIC_VALUES = np.arange(1000, 1000 + 37, dtype=np.uint64)


def fp(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


TABLE_CASES = [("mem_after", traces.T_MEM_AFTER, 7, "std", "valid"), ("logic", traces.T_LOGIC, 6, "test", "valid"),
               ("memory", traces.T_MEMORY, 8, "std", "valid"), ("cpu", traces.T_CPU, 9, "std", "random"),
               ("keccak", traces.T_KECCAK, 5, "test", "valid"), ("byte_packing", traces.T_BYTE_PACKING, 8, "std", "random"),
               ("arithmetic", traces.T_ARITHMETIC, 8, "std", "random"), ("keccak_sponge", traces.T_KECCAK_SPONGE, 5, "test", "random")]


def table_trace(table, lg, kind):
    if kind == "random":
        return traces.random_trace(table, lg, lg)
    if table in (traces.T_MEM_BEFORE, traces.T_MEM_AFTER):
        return traces.memcont_trace(lg, lg)
    if table == traces.T_LOGIC:
        return traces.logic_trace(lg, lg)
    if table == traces.T_MEMORY:
        return traces.memory_trace_simple(lg)
    if table == traces.T_KECCAK:
        rng = np.random.default_rng(lg)
        return traces.keccak_trace(lg, rng.integers(0, 2 ** 63, size=(1, 25), dtype=np.uint64))[0]
    raise ValueError


def main():
    orc = oracle_lib.load()
    g = {"reference_kats": {
        "poseidon_hash_zeros": {"source": "smt_trie/src/keys.rs:10-15", "input": "poseidon([0;12])[0..4]",
                                "value": [4330397376401421145, 14124799381142128323, 8742572140681234676, 14345658006221440202]},
        "empty_consolidated_blockhash": {"source": "evm_arithmetization/src/proof.rs:505-510", "input": "hash_no_pad([0;2048])",
                                         "value": [5498946765822202150, 10724662260254836878, 9161393967331872654, 5704373722058976135]},
        "hash_contract_bytecode_empty": {"source": "smt_trie/src/code.rs:57-67", "input": "hash_contract_bytecode([])",
                                         "value": [10052403398432742521, 15195891732843337299, 2019258788108304834, 4300613462594703212]},
        "goldilocks_inverse_65536": {"source": "evm_arithmetization/src/arithmetic/addcy.rs:67", "value": 18446462594437939201}},
        "oracle_vectors": {}}
    ov = g["oracle_vectors"]
    rng = np.random.default_rng(1234)
    cols = oracle_lib.rand_field(rng, (12, 1 << 10))
    co, le, di, cap = orc.commit(cols, 1, 4)
    ov["commit_12x1024_seed1234"] = {"cap": [int(x) for x in cap.ravel()], "coeffs": fp(co), "leaves": fp(le), "digests": fp(di)}
    x = oracle_lib.rand_field(np.random.default_rng(1), 1 << 16)
    ov["ntt_65536_seed1"] = {"fft": fp(orc.ntt(x, 0)), "coset_fft": fp(orc.ntt(x, 2, oracle_lib.GENERATOR)), "first4_fft": [int(v) for v in orc.ntt(x, 0)[0, :4]]}
    st = oracle_lib.rand_field(np.random.default_rng(2), (64, 12))
    ov["poseidon_64_states_seed2"] = fp(orc.poseidon(st))
    for name, table, lg, cfgname, kind in TABLE_CASES:
        cfg = STANDARD_FAST if cfgname == "std" else TEST_CONFIG
        tr = table_trace(table, lg, kind)
        proof, st2, aux, quot, fri = orc_prove_table(orc, table, cfg, tr, BG2[:2 * cfg[1]], STATE0, debug=True)
        ov["prove_table_%s_2^%d_%s_%s" % (name, lg, cfgname, kind)] = {
            "proof_words": len(proof), "proof": fp(proof), "state_after": [int(v) for v in st2], "aux_values": fp(aux), "quotient_chunks": fp(quot),
            "fri_values": fp(fri), "pow_witness": int(proof[-1])}
    seg = traces.valid_segment(seed=11)
    proofs, bg, caps = orc_prove_segment(orc, TEST_CONFIG, seg, PUBLIC_VALUES)
    ov["prove_segment_valid_seed11_test"] = {"ctl_challenges": [int(v) for v in bg], "caps": fp(caps),
                                             "proofs": [None if p is None else fp(p) for p in proofs]}
    with open(os.path.join(ROOT, "tests", "golden", "golden_r1.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden_r1.json with", len(ov), "oracle vectors")


if __name__ == "__main__":
    main()
