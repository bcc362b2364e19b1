#!/bin/bash
# Final 1-GPU pass: the driver's own commands
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== python -m pytest tests/ -q -m gpu"; timeout 400 python -m pytest tests/ -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/final_pytest_gpu.log | cut -c1-300
echo "== smoke()"; timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench.py (default flags)"; timeout 200 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; cut -c1-2400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
echo "== bench.py --impl reference"; timeout 60 python bench.py --impl reference 2>/dev/null | tail -1 | cut -c1-300
echo "== threads fork-join, 1 GPU (2 virtual hosts)"; timeout 100 build/bin/threads_bench --memory device --hosts 2 --iters 20 --warmup 3 2>/dev/null | tee gpurun_out/final_threads.json | cut -c1-400
echo "== planner variants (128 cores)"
timeout 60 build/bin/planner_bench --mode native --iters 30 2>/dev/null | tail -1 | sed "s/^/default: /" | tee -a gpurun_out/final_planner.jsonl | cut -c1-330
FAABRIC_EXECUTOR_DEQUEUE_SPIN=0 timeout 60 build/bin/planner_bench --mode native --iters 30 2>/dev/null | tail -1 | sed "s/^/dequeue-nospin: /" | tee -a gpurun_out/final_planner.jsonl | cut -c1-330
FAABRIC_SCHED_IDLE_LOCK=spin timeout 60 build/bin/planner_bench --mode native --iters 30 2>/dev/null | tail -1 | sed "s/^/idle-spinlock: /" | tee -a gpurun_out/final_planner.jsonl | cut -c1-330
timeout 60 build/bin/planner_bench --mode refcpu --iters 30 2>/dev/null | tail -1 | sed "s/^/refcpu: /" | tee -a gpurun_out/final_planner.jsonl | cut -c1-330
