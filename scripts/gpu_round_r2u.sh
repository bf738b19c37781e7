# r2u: compute-sanitizer over the parity suite's small cases — memcheck (out-of-bounds / misaligned global, shared and local accesses,
# leaks of device allocations are not tracked) over every kernel family, then racecheck (shared-memory hazards) and synccheck
# (barrier / mbarrier misuse) over the kernels that stage through shared memory (NTT tiles incl. the TMA pass, Merkle tail, quotient blocks).
set -x
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
MEM_CASES="tests/test_gpu_primitives.py::test_poseidon_permutation_kats_and_random tests/test_gpu_primitives.py::test_hash_rows \
 tests/test_gpu_primitives.py::test_commit_values_matches_oracle tests/test_gpu_primitives.py::test_commit_values_other_blowups_match_oracle \
 tests/test_gpu_primitives.py::test_commit_coeffs_matches_oracle tests/test_gpu_stark.py tests/test_gpu_shard.py::test_split_commit_equals_one_device_commit \
 tests/test_gpu_trace_gen.py::test_keccak_trace_matches_reference_restatement tests/test_gpu_trace_gen.py::test_logic_trace_matches_reference_restatement \
 tests/test_gpu_trace_gen.py::test_memory_finish_matches_reference_restatement tests/test_gpu_segment.py::test_valid_segment_matches_oracle_and_verifies"
( time timeout 300 $SAN --tool memcheck --error-exitcode 86 --log-file gpurun_out/r2u_memcheck.log \
    python -m pytest $MEM_CASES -q -x -p no:cacheprovider > gpurun_out/r2u_memcheck_pytest.log 2>&1 ) 2>&1 | grep real
echo "memcheck rc=$?"; tail -3 gpurun_out/r2u_memcheck_pytest.log; grep -c "Invalid\|misaligned" gpurun_out/r2u_memcheck.log; tail -3 gpurun_out/r2u_memcheck.log
( time timeout 100 $SAN --tool memcheck --error-exitcode 86 --log-file gpurun_out/r2u_memcheck_ntt.log \
    python -m pytest tests/test_gpu_primitives.py::test_ntt_all_kinds -q -x -p no:cacheprovider -k "not 20 and not 21 and not 22" > gpurun_out/r2u_memcheck_ntt_pytest.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/r2u_memcheck_ntt_pytest.log; tail -2 gpurun_out/r2u_memcheck_ntt.log
RACE_CASES="tests/test_gpu_primitives.py::test_ntt_all_kinds tests/test_gpu_primitives.py::test_commit_values_matches_oracle"
( time timeout 150 $SAN --tool racecheck --racecheck-report all --error-exitcode 86 --log-file gpurun_out/r2u_racecheck.log \
    python -m pytest $RACE_CASES -q -x -p no:cacheprovider -k "not 16 and not 17 and not 20 and not 21 and not 22" > gpurun_out/r2u_racecheck_pytest.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/r2u_racecheck_pytest.log; tail -3 gpurun_out/r2u_racecheck.log
( time timeout 100 $SAN --tool synccheck --error-exitcode 86 --log-file gpurun_out/r2u_synccheck.log \
    python -m pytest "tests/test_gpu_stark.py" -q -x -p no:cacheprovider > gpurun_out/r2u_synccheck_pytest.log 2>&1 ) 2>&1 | grep real
tail -2 gpurun_out/r2u_synccheck_pytest.log; tail -3 gpurun_out/r2u_synccheck.log
