"""CPU-only checks of the native library: it must load without a GPU, the
merge-region preparation must match the reference's gap filling, and the
Unix-socket bootstrap (allgather / barrier / fd passing) must work between
processes."""

import ctypes as C
import multiprocessing as mp
import os

from faabric_b200 import _lib
from faabric_b200._lib import FbMergeRegion


def test_library_loads_without_gpu(native_lib):
    assert native_lib.fb_cuda_device_count() >= 0
    assert _lib.lib_path().exists()


def prep(regions, size, fill=0):
    lib = _lib.load()
    n_in = len(regions)
    arr = (FbMergeRegion * max(n_in, 1))()
    for i, r in enumerate(regions):
        arr[i] = FbMergeRegion(*r)
    cap = 2 * n_in + 2
    out = (FbMergeRegion * cap)()
    typed = (C.c_int32 * cap)()
    nt = C.c_int(0)
    n = lib.fb_snapshot_prepare_regions(arr, n_in, fill, size, out, cap, typed, C.byref(nt))
    return [(out[i].offset, out[i].length, out[i].dataType, out[i].op) for i in range(n)], list(typed)[: nt.value]


def test_gap_fill_matches_reference_semantics():
    # no regions => one (0, 0, Raw, fill) region
    assert prep([], 8192) == ([(0, 0, 0, 0)], [])
    assert prep([], 8192, fill=7) == ([(0, 0, 0, 7)], [])
    # gaps before / between / after; trailing gap has length 0
    regs, typed = prep([(500, 4, 2, 1), (100, 8, 3, 4)], 10000)
    assert regs == [
        (0, 100, 0, 0),
        (100, 8, 3, 4),
        (108, 392, 0, 0),
        (500, 4, 2, 1),
        (504, 0, 0, 0),
    ]
    assert typed == [1, 3]
    # a region that runs to the end suppresses the trailing gap
    regs, _ = prep([(4096, 0, 0, 6)], 10000)
    assert regs == [(0, 4096, 0, 0), (4096, 0, 0, 6)]


def _bootstrap_worker(rank, n, job, q):
    lib = C.CDLL(str(_lib.lib_path()))
    lib.fb_test_bootstrap.restype = C.c_int
    lib.fb_test_bootstrap.argtypes = [C.c_int, C.c_int, C.c_char_p]
    q.put((rank, lib.fb_test_bootstrap(rank, n, job.encode())))


def test_bootstrap_allgather_barrier_fds():
    n = 3
    job = f"pytest-{os.getpid()}"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_bootstrap_worker, args=(r, n, job, q)) for r in range(n)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=60) for _ in range(n))
    for p in ps:
        p.join(timeout=30)
    assert res == {0: 0, 1: 0, 2: 0}
