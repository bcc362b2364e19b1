#!/bin/bash
# 8-GPU pass: headline at N=8/4/2, NVLink byte counters, multi-GPU tests, all-reduce sweeps (tuning tables)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e_gpus.txt 2>&1
run() { # n port args...
  local n=$1; local port=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"
}
echo "== bench n8 (grouped)"; nvidia-smi nvlink -gt d -i 0 > gpurun_out/e_nvlink_before.txt 2>&1
run 8 29601 --steps 50 --warmup 10 > gpurun_out/e_bench8.json 2> gpurun_out/e_bench8.err; echo "rc=$?"; cat gpurun_out/e_bench8.json | cut -c1-900; tail -2 gpurun_out/e_bench8.err
nvidia-smi nvlink -gt d -i 0 > gpurun_out/e_nvlink_after.txt 2>&1
for gb in 148 64; do echo "== bench n8 groupBlocks=$gb"; FAABRIC_GROUP_BLOCKS=$gb run 8 2961$((gb % 10)) --steps 50 --warmup 10 --no-nccl --no-mpi-api > gpurun_out/e_bench8_gb$gb.json 2> gpurun_out/e_bench8_gb$gb.err; echo "rc=$?"; cut -c1-260 gpurun_out/e_bench8_gb$gb.json; done
echo "== bench n4"; run 4 29602 --steps 50 --warmup 10 --no-mpi-api > gpurun_out/e_bench4.json 2> gpurun_out/e_bench4.err; echo "rc=$?"; cut -c1-420 gpurun_out/e_bench4.json
echo "== bench n2"; run 2 29603 --steps 50 --warmup 10 --no-mpi-api > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err; echo "rc=$?"; cut -c1-420 gpurun_out/e_bench2.json
echo "== multi-gpu tests (8 GPUs, one process)"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/e_multi.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/e_multi.log
echo "== sweep n8"; run 8 29604 --mode sweep --max-bytes 268435456 --out gpurun_out/e_sweep8.json > /dev/null 2> gpurun_out/e_sweep8.err; echo "rc=$?"; grep "\[sweep\]" gpurun_out/e_sweep8.err | cut -c1-330
cp gpurun_out/tuning_N8.json gpurun_out/e_tuning_N8.json 2>/dev/null; cp gpurun_out/tuning_N8.txt gpurun_out/e_tuning_N8.txt 2>/dev/null
echo "== sweep n4"; run 4 29605 --mode sweep --max-bytes 268435456 --out gpurun_out/e_sweep4.json > /dev/null 2> gpurun_out/e_sweep4.err; echo "rc=$?"; grep "\[sweep\]" gpurun_out/e_sweep4.err | cut -c1-330 | tail -12
cp gpurun_out/tuning_N4.json gpurun_out/e_tuning_N4.json 2>/dev/null; cp gpurun_out/tuning_N4.txt gpurun_out/e_tuning_N4.txt 2>/dev/null
echo "== bench n8 lanes (round-1 path, for comparison)"; run 8 29606 --sync-mode lanes --no-nccl --no-mpi-api --steps 20 > gpurun_out/e_bench8_lanes.json 2> gpurun_out/e_bench8_lanes.err; echo "rc=$?"; cut -c1-330 gpurun_out/e_bench8_lanes.json
echo "== threads fork-join through the runtime (8 virtual GPU hosts, 1 GiB)"; timeout 300 build/bin/threads_bench --memory device --hosts 8 --iters 10 --warmup 2 2> gpurun_out/e_threads.err | tee gpurun_out/e_threads.json | cut -c1-420; tail -3 gpurun_out/e_threads.err
timeout 300 build/bin/threads_bench --memory device --hosts 8 --dirty-pct 10 --iters 10 --warmup 2 2>/dev/null | tee -a gpurun_out/e_threads.json | cut -c1-420
timeout 300 build/bin/threads_bench --memory host --hosts 8 --iters 5 --warmup 1 2>/dev/null | tee -a gpurun_out/e_threads.json | cut -c1-420
echo "== MPI C API arms (8 ranks in one worker process)"; for impl in mpi-symmetric; do timeout 300 python bench.py --impl $impl --gpus 8 --steps 10 --warmup 3 2> gpurun_out/e_$impl.err | tail -1 | tee -a gpurun_out/e_mpi_api.jsonl | cut -c1-330; done
