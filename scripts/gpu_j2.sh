#!/bin/bash
# 2-GPU pass: incremental THREADS fork-join (tests + bench, on / off), MPI C-API arm after the lock removal
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== c++ gpu suite"; FAABRIC_TEST_WATCHDOG_SECS=60 timeout 300 build/bin/faabric_tests --tag gpu > gpurun_out/j_cpp_gpu.log 2>&1; echo "rc=$?"; grep -E "OK|FAIL|====" gpurun_out/j_cpp_gpu.log | tail -14; grep -B2 -A12 "FAIL" gpurun_out/j_cpp_gpu.log | head -40
for inc in 1 0; do
echo "== threads fork-join, 2 GPUs, incremental=$inc"
FAABRIC_THREADS_INCREMENTAL=$inc timeout 200 build/bin/threads_bench --memory device --hosts 2 --iters 20 --warmup 3 2> gpurun_out/j_threads.err | tee -a gpurun_out/j_threads.json | cut -c1-420; tail -2 gpurun_out/j_threads.err
FAABRIC_THREADS_INCREMENTAL=$inc timeout 200 build/bin/threads_bench --memory device --hosts 2 --dirty-pct 10 --iters 20 --warmup 3 2>/dev/null | tee -a gpurun_out/j_threads.json | cut -c1-420
FAABRIC_THREADS_INCREMENTAL=$inc timeout 200 build/bin/threads_bench --memory device --hosts 8 --iters 20 --warmup 3 2>/dev/null | tee -a gpurun_out/j_threads.json | cut -c1-420
done
echo "== MPI C-API arm, 2 ranks"; timeout 300 python bench.py --impl mpi-symmetric-nb --gpus 2 --steps 20 --warmup 5 2> gpurun_out/j_mpi_nb.err | tail -1 | tee gpurun_out/j_mpi_nb.json | cut -c1-420
timeout 300 python bench.py --impl mpi-symmetric --gpus 2 --steps 10 --warmup 3 2> gpurun_out/j_mpi_b.err | tail -1 | tee gpurun_out/j_mpi_b.json | cut -c1-420
echo "== pytest gpu: snapshot/state/runtime"; timeout 400 python -m pytest tests/test_gpu_snapshot.py tests/test_gpu_state.py tests/test_gpu_runtime.py -x -q > gpurun_out/j_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/j_pytest.log
echo "== planner (128 cores)"; for i in 1 2; do timeout 200 build/bin/planner_bench --mode native --iters 30 2>/dev/null | tail -1 | tee -a gpurun_out/j_planner.jsonl | cut -c1-330; done
