"""Real multi-GPU checks (skipped with < 2 devices): ranks on distinct GPUs in
one process (peer access / VMM / NVLS multicast when the box exposes it)."""

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ndev():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture(scope="module")
def group():
    from faabric_b200.parallel import LocalGroup

    n = _ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    g = LocalGroup(n, devices=list(range(n)), heapBytes=256 << 20, stageBytes=16 << 20, timeoutMs=8000)
    yield g
    g.close()


@pytest.mark.parametrize("algo", ["ll", "oneshot", "twoshot", "nvls", "auto"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.int32, torch.bfloat16])
def test_allreduce_multi_gpu(group, algo, dtype):
    g = group
    n = g.size
    if algo == "nvls" and not g.comms[0].has_multicast:
        pytest.skip("no NVLS multicast on this box")
    for numel in (4, 1024, 262144, 4 << 20):
        if algo == "ll" and numel * 4 > 65536:
            continue
        sends, recvs = [], []
        for r, c in enumerate(g.comms):
            s = c.empty(numel, dtype)
            s.copy_(((torch.arange(numel) % 13) + r).to(dtype))
            sends.append(s)
            recvs.append(c.empty(numel, dtype))
        for d in range(n):
            torch.cuda.synchronize(d)
        g.run(lambda c, r, st: c.all_reduce(sends[r], recvs[r], algo=algo))
        g.synchronize()
        assert g.check_errors() == [0] * n
        ref = ((torch.arange(numel) % 13).to(torch.float64) * n + n * (n - 1) / 2)
        for r in range(n):
            got = recvs[r].cpu().to(torch.float64)
            tol = 0.0 if dtype != torch.bfloat16 else 0.05 * float(ref.max())
            assert torch.allclose(got, ref, atol=tol, rtol=0), (algo, dtype, numel, r, g.comms[r].last_algo)
        for c, s, o in zip(g.comms, sends, recvs):
            c.free(s)
            c.free(o)


def test_moves_and_p2p_multi_gpu(group):
    g = group
    n = g.size
    numel = 100000
    sends = [c.empty(numel, torch.int32) for c in g.comms]
    for r, s in enumerate(sends):
        s.copy_(torch.arange(numel, dtype=torch.int32) + 7 * r)
    outs = [c.empty(numel * n, torch.int32) for c in g.comms]
    for d in range(n):
        torch.cuda.synchronize(d)
    g.run(lambda c, r, st: c.all_gather(sends[r], outs[r]))
    g.synchronize()
    assert g.check_errors() == [0] * n
    ref = torch.cat([torch.arange(numel, dtype=torch.int32) + 7 * r for r in range(n)])
    for r in range(n):
        assert torch.equal(outs[r].cpu(), ref)
    a2a = [c.empty(numel * n, torch.int32) for c in g.comms]
    g.run(lambda c, r, st: c.all_to_all(outs[r], a2a[r]))
    g.synchronize()
    assert g.check_errors() == [0] * n
    for r in range(n):
        exp = ref[r * numel : (r + 1) * numel].repeat(n)
        assert torch.equal(a2a[r].cpu(), exp)
    # ring send/recv across GPUs
    dsts = [torch.zeros(numel, dtype=torch.int32, device=f"cuda:{c.device}") for c in g.comms]
    side = [torch.cuda.Stream(device=c.device) for c in g.comms]
    for r, c in enumerate(g.comms):
        with torch.cuda.device(c.device):
            c.recv(dsts[r], (r - 1) % n, stream=side[r])
    g.run(lambda c, r, st: c.send(sends[r], (r + 1) % n))
    g.synchronize()
    for s in side:
        s.synchronize()
    assert g.check_errors() == [0] * n
    for r in range(n):
        assert torch.equal(dsts[r].cpu(), sends[(r - 1) % n].cpu())
    # snapshot push into GPU 0's image from every other GPU
    from faabric_b200.ops import snapshot as snap

    size = 1 << 22
    main = g.comms[0].zeros(size, torch.uint8)
    others = [c.zeros(size, torch.uint8) for c in g.comms[1:]]  # keep heaps symmetric
    main_ptr = main.data_ptr()
    for r, c in enumerate(g.comms):
        with torch.cuda.device(c.device):
            base = torch.zeros(size, dtype=torch.uint8, device=f"cuda:{c.device}")
            mem = base.clone()
            mem[4096 * (r + 1) : 4096 * (r + 1) + 100] = r + 1
            snap.diff_push(mem, base, main_ptr)
            torch.cuda.synchronize(c.device)
    got = main.cpu()
    for r in range(n):
        assert bool((got[4096 * (r + 1) : 4096 * (r + 1) + 100] == r + 1).all())
    assert int(got.sum()) == sum(100 * (r + 1) for r in range(n))
    del others


def test_grouped_allreduce_multi_gpu(group):
    """The flagship path on real NVLink: 214 ResNet-50 gradient tensors in ONE
    launch per rank, in-kernel barriers, checked tensor by tensor."""
    from faabric_b200.models import GradientSync, resnet50_grad_sizes

    g = group
    n = g.size
    assert not g.comms[0].stream_sync
    sizes = resnet50_grad_sizes()[:60] + [1, 3, 17]
    syncs = [GradientSync(c, sizes, dtype=torch.int32, mode="grouped") for c in g.comms]
    for r, s in enumerate(syncs):
        with torch.cuda.device(s.comm.device):
            s.send.copy_(torch.arange(s.send.numel(), dtype=torch.int32) % 1000 + r)
    for d in range(n):
        torch.cuda.synchronize(d)
    for _ in range(3):
        g.run(lambda c, r, st: syncs[r].step(st))
    g.synchronize()
    assert g.check_errors() == [0] * n
    pos = torch.arange(syncs[0].send.numel(), dtype=torch.int64) % 1000
    exp = (pos * n + n * (n - 1) // 2).to(torch.int32)
    for s in syncs:
        got = s.recv.cpu()
        for o, sz in zip(s.offsets, s.sizes):
            assert torch.equal(got[o : o + sz], exp[o : o + sz])
    for s in syncs:
        s.close()


def test_sendrecv_multi_gpu(group):
    g = group
    n = g.size
    nbytes = 20 << 20
    srcs = [torch.full((nbytes,), r + 1, dtype=torch.uint8, device=f"cuda:{c.device}") for r, c in enumerate(g.comms)]
    dsts = [torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{c.device}") for c in g.comms]
    for d in range(n):
        torch.cuda.synchronize(d)
    g.run(lambda c, r, st: c.send_recv(srcs[r], (r + 1) % n, dsts[r], (r - 1) % n))
    g.synchronize()
    assert g.check_errors() == [0] * n
    for r in range(n):
        assert int(dsts[r][0]) == (r - 1) % n + 1 and int(dsts[r][-1]) == (r - 1) % n + 1
