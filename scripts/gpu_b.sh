#!/bin/bash
# second GPU pass (2 GPUs): C++ gpu tests, full pytest, smoke under ncu, bench N=1/2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi -L > gpurun_out/b_gpus.txt 2>&1
echo "== c++ gpu tests"; timeout 900 build/bin/faabric_tests --tag gpu > gpurun_out/b_cpp_gpu.log 2>&1; echo "cpp rc=$?"; grep -E "FAIL|====|fatal" gpurun_out/b_cpp_gpu.log | head -20
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/b_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/b_smoke.log
echo "== ncu smoke"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_ncu_smoke.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/b_ncu_smoke.log
grep -o '"[a-zA-Z_:0-9<>, ]*Kernel[^"]*"' gpurun_out/b_smoke_launches.csv | sed 's/<.*//;s/(.*//' | sort | uniq -c | sort -rn | head -30
echo "== bench n1"; timeout 300 python bench.py --gpus 1 > gpurun_out/b_bench1.json 2> gpurun_out/b_bench1.err; echo "rc=$?"; cut -c1-400 gpurun_out/b_bench1.json
echo "== bench n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/b_bench2.json 2> gpurun_out/b_bench2.err; echo "rc=$?"; cat gpurun_out/b_bench2.json; tail -5 gpurun_out/b_bench2.err
echo "== bench n2 lanes"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --sync-mode lanes --no-nccl > gpurun_out/b_bench2_lanes.json 2> gpurun_out/b_bench2_lanes.err; echo "rc=$?"; cut -c1-300 gpurun_out/b_bench2_lanes.json
