#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== c++ gpu tests with watchdog"; FAABRIC_TEST_WATCHDOG_SECS=45 timeout 300 stdbuf -oL -eL build/bin/faabric_tests --tag gpu > gpurun_out/g_cpp_gpu.log 2>&1; echo "cpp rc=$?"; grep -E "OK|FAIL|SKIP|====|fatal|what|watchdog" gpurun_out/g_cpp_gpu.log | cut -c1-200 | tail -16
grep -A14 "^---- thread" gpurun_out/g_cpp_gpu.log | grep -v "libcuda\|^--$" | cut -c1-190 | head -120
echo "== individually: ptp tests alone"; FAABRIC_TEST_WATCHDOG_SECS=45 timeout 200 stdbuf -oL build/bin/faabric_tests "ptp " > gpurun_out/g_cpp_ptp.log 2>&1; echo "rc=$?"; grep -E "OK|FAIL|====|watchdog" gpurun_out/g_cpp_ptp.log | tail -5
echo "== pytest snapshot/state"; timeout 300 python -m pytest tests/test_gpu_snapshot.py tests/test_gpu_state.py -q > gpurun_out/g_pytest_rest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/g_pytest_rest.log
echo "== bench n1"; timeout 200 python bench.py --gpus 1 > gpurun_out/g_bench1.json 2> gpurun_out/g_bench1.err; echo "rc=$?"; cut -c1-330 gpurun_out/g_bench1.json
echo "== planner fan-out variants ($(nproc) cores)"
for v in workers direct; do FAABRIC_PLANNER_RESULTS=$v timeout 200 build/bin/planner_bench --mode native --iters 30 2>/dev/null | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/g_planner.jsonl | cut -c1-250; done
timeout 200 build/bin/planner_bench --mode refcpu --iters 30 2>/dev/null | tail -1 | tee -a gpurun_out/g_planner.jsonl | cut -c1-250
echo "== threads fork-join (2 virtual hosts on one GPU, 1 GiB)"; timeout 200 build/bin/threads_bench --memory device --hosts 2 --iters 10 --warmup 2 2> gpurun_out/g_threads.err | tee gpurun_out/g_threads.json | cut -c1-420; tail -3 gpurun_out/g_threads.err
