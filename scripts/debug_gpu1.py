import sys, time, torch
sys.path.insert(0, '.')
from faabric_b200.parallel import LocalGroup

def mk(n=2, **kw):
    return LocalGroup(n, devices=[0]*n, heapBytes=64 << 20, stageBytes=4 << 20, maxBlocks=8, timeoutMs=2000, **kw)

def seq_test(name, algos_sizes, n=2):
    g = mk(n)
    bufs = {}
    for r, c in enumerate(g.comms):
        bufs[r] = [(c.empty(s, torch.int32), c.empty(s, torch.int32)) for _, s in algos_sizes]
        for a, b in bufs[r]:
            a.fill_(r + 1)
    torch.cuda.synchronize()
    t0 = time.time()
    def issue(c, r, st):
        for (algo, s), (a, b) in zip(algos_sizes, bufs[r]):
            c.all_reduce(a, b, algo=algo)
    g.run(issue)
    g.synchronize()
    errs = g.check_errors()
    ok = all(bool((b == n*(n+1)//2).all()) for r in range(n) for a, b in bufs[r])
    print(f"{name}: errs={errs} ok={ok} t={time.time()-t0:.2f}s", flush=True)
    g.close()

seq_test("ll once", [("ll", 8)])
seq_test("ll twice", [("ll", 8), ("ll", 8)])
seq_test("ll x4", [("ll", 8), ("ll", 1000), ("ll", 8), ("ll", 1000)])
seq_test("oneshot x3", [("oneshot", 1000)] * 3)
seq_test("twoshot x3", [("twoshot", 70000)] * 3)
seq_test("ll+twoshot", [("ll", 8), ("twoshot", 70000)])
seq_test("twoshot+ll", [("twoshot", 70000), ("ll", 8)])
seq_test("mix", [("auto", 8), ("auto", 1000), ("auto", 70000), ("auto", 300000)])

# p2p
g = mk(2)
a, b = g.comms
for nbytes in (0, 16, 4096, 1 << 20):
    src = torch.arange(nbytes, dtype=torch.int64).to(torch.uint8).cuda()
    dst = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    t0 = time.time()
    b.recv(dst, 0, stream=g.streams[1])
    a.send(src, 1, stream=g.streams[0])
    g.synchronize()
    print(f"p2p {nbytes}: errs={g.check_errors()} ok={torch.equal(src, dst)} t={time.time()-t0:.2f}s", flush=True)
g.close()
