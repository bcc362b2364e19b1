#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for b in off on; do for c in 8 16; do timeout 60 python bench.py --no-mpi-api --steps 100 --warmup 10 --bind-numa $b --e2e-chunks $c 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bind=$b chunks=$c', d['ms_per_step'], d['e2e'])"; done; done
nvidia-smi topo -m 2>/dev/null | head -6
