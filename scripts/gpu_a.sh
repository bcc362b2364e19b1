#!/bin/bash
# first GPU pass of round 2: smoke, new tests, full suite, bench N=1, ncu launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/a_smoke.log
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_collectives.py -x -q -k "send or grouped or in_kernel or aborted or graph or watchdog" > gpurun_out/a_newtests.log 2>&1; echo "new rc=$?"; tail -15 gpurun_out/a_newtests.log
echo "== full gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/a_fullsuite.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/a_fullsuite.log
echo "== bench n1"; timeout 300 python bench.py --gpus 1 > gpurun_out/a_bench1.json 2> gpurun_out/a_bench1.err; echo "bench rc=$?"; cat gpurun_out/a_bench1.json; tail -3 gpurun_out/a_bench1.err
echo "== ncu smoke launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_ncu_smoke.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/a_ncu_smoke.log
grep -o '"[a-zA-Z_:0-9<>, ]*Kernel[^"]*"' gpurun_out/a_smoke_launches.csv | sed 's/<.*//' | sort | uniq -c | sort -rn | head -30
