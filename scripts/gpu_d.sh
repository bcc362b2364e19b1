#!/bin/bash
# fourth GPU pass (1 GPU): the driver's own commands, planner fan-out on the big box, ncu full capture of the flagship kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu (driver command, no env)"; timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/d_pytest.log
echo "== smoke (driver)"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/d_smoke.log
echo "== planner fan-out (CPU, $(nproc) cores)"
for m in native refcpu; do timeout 300 build/bin/planner_bench --mode $m --iters 30 2>/dev/null | tail -1 | tee -a gpurun_out/d_planner.jsonl | cut -c1-260; done
timeout 120 build/bin/planner_bench --functions 128 --iters 50 2>/dev/null | tail -1 | tee -a gpurun_out/d_planner.jsonl | cut -c1-200
timeout 200 build/bin/planner_bench --profile --iters 40 2> gpurun_out/d_planner_profile.txt | tail -1 | cut -c1-200; head -16 gpurun_out/d_planner_profile.txt
echo "== ncu full: groupAllReduceKernel (N=1, 97.6 MiB)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:groupAllReduceKernel -s 3 -c 2 -o gpurun_out/prof_group python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/d_ncu_group.log 2>&1; echo "ncu rc=$?"
echo "== ncu full: statePushDirtyKernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:statePushDirtyKernel -c 1 -o gpurun_out/prof_state python -m pytest tests/test_gpu_state.py -x -q -k all_blocks > gpurun_out/d_ncu_state.log 2>&1; echo "ncu rc=$?"
echo "== mpi C api numbers (N=2 ranks on one GPU is not representative; skipped)"
