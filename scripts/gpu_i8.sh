#!/bin/bash
# short 8-GPU pass: THREADS fork-join over 8 GPUs, headline with the MPI C-API arm, blocking MPI arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== threads fork-join through the runtime (8 GPUs, 1 GiB)"; timeout 200 build/bin/threads_bench --memory device --hosts 8 --iters 10 --warmup 2 2> gpurun_out/i_threads.err | tee gpurun_out/i_threads.json | cut -c1-420; tail -3 gpurun_out/i_threads.err
timeout 200 build/bin/threads_bench --memory device --hosts 8 --dirty-pct 10 --iters 10 --warmup 2 2>/dev/null | tee -a gpurun_out/i_threads.json | cut -c1-420
echo "== bench n8 (headline + mpi_api)"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/i_bench8.json 2> gpurun_out/i_bench8.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/i_bench8.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("ms_per_step","busbw_GBps","vs_nccl","mpi_api","e2e")})
PY
echo "== blocking MPI_Allreduce arm"; timeout 300 python bench.py --impl mpi-symmetric --gpus 8 --steps 10 --warmup 3 2> gpurun_out/i_mpi_blocking.err | tail -1 | tee gpurun_out/i_mpi_blocking.json | cut -c1-400
