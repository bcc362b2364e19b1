#!/bin/bash
# ncu capture of the incremental fork-join kernels (1 GPU, 2 virtual hosts)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'pageSync|pagePull|snapshotDiffPush' -s 9 -c 6 -o gpurun_out/prof_forkjoin build/bin/threads_bench --memory device --hosts 2 --iters 6 --warmup 2 > gpurun_out/k_ncu.log 2>&1; echo "ncu rc=$?"; tail -4 gpurun_out/k_ncu.log | cut -c1-300
ls -la gpurun_out/prof_forkjoin.ncu-rep
