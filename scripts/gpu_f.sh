#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== c++ gpu tests (direct, line-buffered)"; timeout 400 stdbuf -oL -eL build/bin/faabric_tests --tag gpu > gpurun_out/f_cpp_gpu.log 2>&1; echo "cpp rc=$?"; grep -E "OK|FAIL|SKIP|====|fatal|what" gpurun_out/f_cpp_gpu.log | cut -c1-200 | tail -20
echo "== same, CUDA_DEVICE_MAX_CONNECTIONS=32"; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 400 stdbuf -oL -eL build/bin/faabric_tests --tag gpu > gpurun_out/f_cpp_gpu32.log 2>&1; echo "cpp rc=$?"; grep -E "FAIL|====|fatal|what" gpurun_out/f_cpp_gpu32.log | cut -c1-200 | tail -8
echo "== planner fan-out (CPU, $(nproc) cores)"
for m in native refcpu; do timeout 300 build/bin/planner_bench --mode $m --iters 30 2>/dev/null | tail -1 | tee -a gpurun_out/f_planner.jsonl | cut -c1-260; done
timeout 120 build/bin/planner_bench --functions 128 --iters 50 2>/dev/null | tail -1 | tee -a gpurun_out/f_planner.jsonl | cut -c1-200
timeout 200 build/bin/planner_bench --profile --iters 40 2> gpurun_out/f_planner_profile.txt | tail -1 | cut -c1-200; head -14 gpurun_out/f_planner_profile.txt
echo "== remaining pytest files"; timeout 600 python -m pytest tests/test_gpu_runtime.py tests/test_gpu_snapshot.py tests/test_gpu_state.py -x -q > gpurun_out/f_pytest_rest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/f_pytest_rest.log
