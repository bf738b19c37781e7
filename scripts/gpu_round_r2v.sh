# r2v: the wider compute-sanitizer pass (r2u took 40 s): memcheck over the whole GPU suite but the full-size cases, initcheck
# (reads of device memory nobody wrote) over primitives / single tables / segments, racecheck over every kernel that stages through
# shared memory (quotient blocks, Merkle tail, trace generation, split commits).
set -x
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL="tests/test_gpu_primitives.py tests/test_gpu_stark.py tests/test_gpu_shard.py tests/test_gpu_trace_gen.py tests/test_gpu_segment.py tests/test_gpu_zz_scheduler.py tests/test_golden.py tests/test_host_mirror.py"
( time timeout 500 $SAN --tool memcheck --error-exitcode 86 --log-file gpurun_out/r2v_memcheck.log \
    python -m pytest $SMALL -m gpu -q -p no:cacheprovider -k "not 20 and not 21 and not 22" > gpurun_out/r2v_memcheck_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r2v_memcheck_pytest.log; tail -2 gpurun_out/r2v_memcheck.log
( time timeout 250 $SAN --tool initcheck --error-exitcode 86 --log-file gpurun_out/r2v_initcheck.log \
    python -m pytest tests/test_gpu_primitives.py tests/test_gpu_stark.py tests/test_gpu_segment.py tests/test_gpu_shard.py -m gpu -q -p no:cacheprovider -k "not 20 and not 21 and not 22" > gpurun_out/r2v_initcheck_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r2v_initcheck_pytest.log; grep -c "Uninitialized" gpurun_out/r2v_initcheck.log; tail -2 gpurun_out/r2v_initcheck.log
( time timeout 250 $SAN --tool racecheck --racecheck-report all --error-exitcode 86 --log-file gpurun_out/r2v_racecheck.log \
    python -m pytest tests/test_gpu_stark.py tests/test_gpu_shard.py tests/test_gpu_trace_gen.py tests/test_gpu_segment.py -m gpu -q -p no:cacheprovider > gpurun_out/r2v_racecheck_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r2v_racecheck_pytest.log; tail -3 gpurun_out/r2v_racecheck.log
head -c 3000 gpurun_out/r2v_initcheck.log
