#!/bin/bash
# third GPU pass (1 GPU): C++ gpu tests, full pytest (driver shape), sanitizer on single-rank paths
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== c++ gpu tests"; timeout 900 build/bin/faabric_tests --tag gpu > gpurun_out/c_cpp_gpu.log 2>&1; echo "cpp rc=$?"; grep -E "FAIL|====|fatal|what" gpurun_out/c_cpp_gpu.log | head -20
echo "== pytest gpu (driver command)"; timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c_pytest.log
echo "== sanitizer memcheck: snapshot + state + single-rank smoke"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/c_memcheck_snapshot.log python -m pytest tests/test_gpu_snapshot.py tests/test_gpu_state.py -x -q > gpurun_out/c_memcheck_snapshot.out 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/c_memcheck_snapshot.out; tail -4 gpurun_out/c_memcheck_snapshot.log
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/c_racecheck_state.log python -m pytest tests/test_gpu_state.py -x -q -k "4096 or 127" > gpurun_out/c_racecheck_state.out 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/c_racecheck_state.log
echo "== bench n1"; timeout 300 python bench.py --gpus 1 > gpurun_out/c_bench1.json 2> gpurun_out/c_bench1.err; echo "rc=$?"; cut -c1-300 gpurun_out/c_bench1.json
