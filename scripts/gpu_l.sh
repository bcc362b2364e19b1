#!/bin/bash
# ncu capture of the flagship kernel at N=1 (NR=1 variant of groupAllReduceKernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:groupAllReduceKernel -s 5 -c 2 -o gpurun_out/prof_group_r2 python bench.py --gpus 1 --steps 4 --warmup 3 --no-mpi-api > gpurun_out/l_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/l_ncu.log | cut -c1-200
ls -la gpurun_out/prof_group_r2.ncu-rep
