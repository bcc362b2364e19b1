#!/bin/bash
# 2-GPU pass: C++ gpu suite across two real GPUs, THREADS fork-join over NVLink, MPI C-API arm, multi tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== c++ gpu suite on 2 GPUs"; FAABRIC_TEST_WATCHDOG_SECS=60 timeout 400 build/bin/faabric_tests --tag gpu > gpurun_out/h_cpp_gpu.log 2>&1; echo "rc=$?"; grep -E "OK|FAIL|====" gpurun_out/h_cpp_gpu.log | tail -14
echo "== threads fork-join, 2 GPUs"; timeout 200 build/bin/threads_bench --memory device --hosts 2 --iters 10 --warmup 2 2> gpurun_out/h_threads.err | tee gpurun_out/h_threads.json | cut -c1-420; tail -3 gpurun_out/h_threads.err
timeout 200 build/bin/threads_bench --memory device --hosts 2 --dirty-pct 10 --iters 10 --warmup 2 2>/dev/null | tee -a gpurun_out/h_threads.json | cut -c1-420
timeout 300 build/bin/threads_bench --memory host --hosts 2 --iters 5 --warmup 1 2>/dev/null | tee -a gpurun_out/h_threads.json | cut -c1-420
echo "== bench n2 with the MPI C-API arm"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/h_bench2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("ms_per_step","busbw_GBps","vs_nccl","mpi_api","e2e")})
PY
echo "== multi-gpu pytest"; timeout 500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/h_multi.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/h_multi.log
echo "== planner fan-out variants (128 cores)"
for v in combine workers; do FAABRIC_PLANNER_RESULTS=$v timeout 200 build/bin/planner_bench --mode native --iters 30 2>/dev/null | tail -1 | sed "s/^/$v: /" | tee -a gpurun_out/h_planner.jsonl | cut -c1-330; done
timeout 200 build/bin/planner_bench --mode refcpu --iters 30 2>/dev/null | tail -1 | tee -a gpurun_out/h_planner.jsonl | cut -c1-330
PROFILE_ROOT=callFunctions timeout 200 build/bin/planner_bench --profile --iters 100 2> gpurun_out/h_planner_profile.txt | tail -1 | cut -c1-200
