"""Single-GPU driver for ncu captures of the hot kernels (one rank, so ncu's
kernel serialisation cannot deadlock a cross-rank barrier)."""
import sys
import torch
sys.path.insert(0, ".")
from faabric_b200.ops import snapshot as snap
from faabric_b200.parallel import LocalGroup

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda:0"
if which in ("all", "snapshot"):
    size = 1 << 30
    base = torch.zeros(size, dtype=torch.uint8, device=dev)
    mem = base.clone()
    n_pages = size // 4096
    perm = torch.randperm(n_pages, device=dev)[: n_pages // 100]
    mem.view(n_pages, 4096)[perm] = 7
    main = base.clone()
    regs = snap.prepare_regions([], size, dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    for _ in range(4):
        snap.diff_push(mem, base, main, regs, stats=stats)
    for _ in range(4):
        snap.dirty_scan(mem, base)
    torch.cuda.synchronize()
if which in ("all", "allreduce"):
    g = LocalGroup(1, devices=[0], heapBytes=640 << 20, stageBytes=4 << 20, maxBlocks=64, channels=1)
    c = g.comms[0]
    a = c.empty(64 << 20, torch.float32)
    b = c.empty(64 << 20, torch.float32)
    a.fill_(1.0)
    for algo in ("twoshot", "oneshot"):
        for _ in range(4):
            c.all_reduce(a, b, algo=algo)
    small = c.empty(4096, torch.float32)
    out = c.empty(4096, torch.float32)
    for _ in range(4):
        c.all_reduce(small, out, algo="ll")
    torch.cuda.synchronize()
    g.close()
print("done")
