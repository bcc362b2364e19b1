"""Device snapshot kernels vs a NumPy oracle that mirrors the reference's
diffWithDirtyRegions / addDiffs / applyDiff semantics
(src/util/snapshot.cpp:402-492,524-578,652-824)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from faabric_b200.ops import snapshot as snap  # noqa: E402

PAGE = 4096


def oracle(orig, mem, main, regions, dirty_pages=None):
    """Apply the reference semantics on host arrays (uint8). Returns new main
    and the number of differing bytes merged."""
    size = min(len(orig), len(mem))
    main = main.copy()
    diff_bytes = 0
    n_pages = (size + PAGE - 1) // PAGE
    dirty = np.ones(n_pages, dtype=bool) if dirty_pages is None else dirty_pages.astype(bool)
    for r in regions:
        beg = r.offset
        end = size if r.length == 0 else min(size, r.offset + r.length)
        if beg >= size or r.op == snap.IGNORE:
            continue
        if r.op in (snap.BYTEWISE, snap.XOR):
            for p in range(beg // PAGE, (end + PAGE - 1) // PAGE):
                if not dirty[p]:
                    continue
                b = max(beg, p * PAGE)
                e = min(end, (p + 1) * PAGE)
                o = orig[b:e]
                m = mem[b:e]
                d = o != m
                diff_bytes += int(d.sum())
                if r.op == snap.BYTEWISE:
                    main[b:e][d] = m[d]
                else:
                    main[b:e] ^= o ^ m
        else:
            np_t = {snap.INT: np.int32, snap.LONG: np.int64, snap.FLOAT: np.float32, snap.DOUBLE: np.float64}[
                r.data_type
            ]
            sz = np.dtype(np_t).itemsize
            count = (end - beg) // sz
            for k in range(count):
                off = beg + k * sz
                if not (dirty[off // PAGE] or dirty[(off + sz - 1) // PAGE]):
                    continue
                o = orig[off : off + sz].view(np_t)[0]
                m = mem[off : off + sz].view(np_t)[0]
                if o == m:
                    continue
                diff_bytes += sz
                c = main[off : off + sz].view(np_t)[0]
                with np.errstate(all="ignore"):
                    if r.op == snap.SUM:
                        v = np_t(c + (m - o))
                    elif r.op == snap.SUBTRACT:
                        v = np_t(c - (o - m))
                    elif r.op == snap.PRODUCT:
                        q = np_t(0) if o == 0 else np_t(m / o) if np_t in (np.float32, np.float64) else np_t(int(m) // int(o) if (int(m) % int(o) == 0 or (int(m) ^ int(o)) >= 0) else -(-int(m) // int(o)))
                        v = np_t(c * q)
                    elif r.op == snap.MAX:
                        v = max(c, m)
                    elif r.op == snap.MIN:
                        v = min(c, m)
                main[off : off + sz] = np.array([v], dtype=np_t).view(np.uint8)
    return main, diff_bytes


def dev(a):
    return torch.from_numpy(a.copy()).cuda()


@pytest.mark.parametrize("size", [4096 * 8, 4096 * 33 + 777, 1000])
@pytest.mark.parametrize("fill", [snap.BYTEWISE, snap.XOR])
def test_diff_push_bytewise_and_xor(size, fill):
    rng = np.random.default_rng(size + fill)
    orig = rng.integers(0, 256, size, dtype=np.uint8)
    mem = orig.copy()
    # sparse byte edits, one dense page, an edit at the very end
    idx = rng.integers(0, size, 200)
    mem[idx] = rng.integers(0, 256, len(idx), dtype=np.uint8)
    if size > PAGE * 3:
        mem[PAGE * 2 : PAGE * 3] = rng.integers(0, 256, PAGE, dtype=np.uint8)
    mem[-1] ^= 0x5A
    main = orig.copy()
    # another writer already changed some other bytes of main
    other = rng.integers(0, size, 100)
    main[other] ^= 0xFF

    regs = snap.prepare_regions([], size, "cuda", fill_op=fill)
    assert regs.n == 1 and regs.host[0].length == 0
    d_mem, d_orig, d_main = dev(mem), dev(orig), dev(main)
    chunk_flags = torch.zeros((size + 127) // 128, dtype=torch.uint8, device="cuda")
    page_flags = torch.zeros((size + PAGE - 1) // PAGE, dtype=torch.uint8, device="cuda")
    stats = snap.diff_push(d_mem, d_orig, d_main, regs, chunk_flags=chunk_flags, page_flags_out=page_flags)
    torch.cuda.synchronize()
    exp_main, exp_bytes = oracle(orig, mem, main, regs.host)
    assert np.array_equal(d_main.cpu().numpy(), exp_main)
    assert int(stats[0].item()) == exp_bytes
    # page flags agree with a direct comparison
    n_pages = (size + PAGE - 1) // PAGE
    exp_pages = np.array([(orig[p * PAGE : (p + 1) * PAGE] != mem[p * PAGE : (p + 1) * PAGE]).any() for p in range(n_pages)])
    assert np.array_equal(page_flags.cpu().numpy().astype(bool), exp_pages)
    assert int(stats[1].item()) == int(exp_pages.sum())
    # the runs derived from chunk flags cover EVERY byte that produced a diff:
    # unaligned heads / tails of regions, the tail of the image and the typed
    # scalars included (ignored regions produce no diff)
    runs = snap.chunk_runs(chunk_flags, size)
    covered = np.zeros(size, dtype=bool)
    for off, ln in runs:
        covered[off : off + ln] = True
    differs = orig != mem
    for r in regs.host:
        if r.op == snap.IGNORE:
            differs[r.offset : (r.offset + r.length) if r.length else size] = False
    assert covered[differs].all()


def test_dirty_page_hint_skips_clean_pages():
    size = PAGE * 16
    rng = np.random.default_rng(1)
    orig = rng.integers(0, 256, size, dtype=np.uint8)
    mem = orig.copy()
    mem[PAGE * 3 + 5] ^= 1
    mem[PAGE * 9 + 100] ^= 1  # changed but NOT flagged dirty -> must be ignored
    dirty = np.zeros(16, dtype=np.uint8)
    dirty[3] = 1
    main = orig.copy()
    regs = snap.prepare_regions([], size, "cuda")
    d_main = dev(main)
    stats = snap.diff_push(dev(mem), dev(orig), d_main, regs, dirty_pages=dev(dirty))
    torch.cuda.synchronize()
    exp, nbytes = oracle(orig, mem, main, regs.host, dirty)
    assert np.array_equal(d_main.cpu().numpy(), exp)
    assert int(stats[0].item()) == nbytes == 1


def test_typed_regions_and_update_base():
    size = PAGE * 4
    rng = np.random.default_rng(7)
    orig = rng.integers(0, 256, size, dtype=np.uint8)
    mem = orig.copy()
    main = orig.copy()

    def put(arr, off, val, t):
        arr[off : off + np.dtype(t).itemsize] = np.array([val], dtype=t).view(np.uint8)

    R = snap.MergeRegion
    regions = [
        R(64, 4, snap.INT, snap.SUM),
        R(128, 8, snap.LONG, snap.MAX),
        R(256, 4, snap.FLOAT, snap.MIN),
        R(512, 8, snap.DOUBLE, snap.SUM),
        R(1024, 4, snap.INT, snap.SUBTRACT),
        R(2048, 16, snap.INT, snap.SUM),  # array of 4 ints
        R(3000, 100, snap.RAW, snap.IGNORE),
        R(PAGE + 3, 8, snap.DOUBLE, snap.PRODUCT),  # unaligned scalar
        R(PAGE * 2, 4, snap.FLOAT, snap.PRODUCT),
    ]
    put(orig, 64, 10, np.int32), put(mem, 64, 17, np.int32), put(main, 64, 100, np.int32)
    put(orig, 128, 5, np.int64), put(mem, 128, 99, np.int64), put(main, 128, 50, np.int64)
    put(orig, 256, 2.0, np.float32), put(mem, 256, -3.5, np.float32), put(main, 256, 1.0, np.float32)
    put(orig, 512, 1.5, np.float64), put(mem, 512, 4.0, np.float64), put(main, 512, 10.0, np.float64)
    put(orig, 1024, 30, np.int32), put(mem, 1024, 20, np.int32), put(main, 1024, 7, np.int32)
    for k in range(4):
        put(orig, 2048 + 4 * k, k, np.int32), put(mem, 2048 + 4 * k, k * 3, np.int32), put(main, 2048 + 4 * k, 1000, np.int32)
    mem[3000:3100] ^= 0xFF  # ignored
    put(orig, PAGE + 3, 2.0, np.float64), put(mem, PAGE + 3, 6.0, np.float64), put(main, PAGE + 3, 5.0, np.float64)
    put(orig, PAGE * 2, 4.0, np.float32), put(mem, PAGE * 2, 2.0, np.float32), put(main, PAGE * 2, 8.0, np.float32)
    mem[PAGE * 3 + 10] ^= 0x11  # bytewise gap

    regs = snap.prepare_regions(regions, size, "cuda")
    d_mem, d_orig, d_main = dev(mem), dev(orig), dev(main)
    snap.diff_push(d_mem, d_orig, d_main, regs, update_base=True)
    torch.cuda.synchronize()
    exp, _ = oracle(orig, mem, main, regs.host)
    got = d_main.cpu().numpy()
    assert np.array_equal(got, exp)
    assert got[64:68].view(np.int32)[0] == 107
    assert got[128:136].view(np.int64)[0] == 99
    assert got[512:520].view(np.float64)[0] == 12.5
    assert got[1024:1028].view(np.int32)[0] == -3
    assert got[PAGE + 3 : PAGE + 11].view(np.float64)[0] == 15.0
    # update_base folded the changes into the base everywhere except Ignore
    base = d_orig.cpu().numpy()
    keep = np.ones(size, dtype=bool)
    keep[3000:3100] = False
    assert np.array_equal(base[keep], mem[keep])
    assert np.array_equal(base[3000:3100], orig[3000:3100])
    # second pass: nothing left to push
    stats = snap.diff_push(d_mem, d_orig, d_main, regs)
    torch.cuda.synchronize()
    assert int(stats[0].item()) == 0


def test_concurrent_writers_merge():
    """Two executors push disjoint byte edits + the same Sum scalar into one
    main image concurrently (byte-exact stores + atomics make this safe)."""
    size = PAGE * 64
    rng = np.random.default_rng(11)
    orig = rng.integers(0, 256, size, dtype=np.uint8)
    main = orig.copy()
    mems = [orig.copy(), orig.copy()]
    # interleaved single-byte edits inside the same 16-byte vectors
    for i in range(1000, size - 16, 37):
        mems[0][i] ^= 0x0F
        mems[1][i + 1] ^= 0xF0
    for w, v in enumerate((5, 11)):
        mems[w][0:4] = np.array([orig[0:4].view(np.int32)[0] + v], dtype=np.int32).view(np.uint8)
    regs = snap.prepare_regions([snap.MergeRegion(0, 4, snap.INT, snap.SUM)], size, "cuda")
    d_main = dev(main)
    d_orig = dev(orig)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    d_mems = [dev(m) for m in mems]
    torch.cuda.synchronize()
    for w in range(2):
        with torch.cuda.stream(streams[w]):
            snap.diff_push(d_mems[w], d_orig, d_main, regs)
    torch.cuda.synchronize()
    exp = orig.copy()
    for i in range(1000, size - 16, 37):
        exp[i] ^= 0x0F
        exp[i + 1] ^= 0xF0
    exp[0:4] = np.array([orig[0:4].view(np.int32)[0] + 16], dtype=np.int32).view(np.uint8)
    assert np.array_equal(d_main.cpu().numpy(), exp)


def test_concurrent_writers_merge_unaligned_typed_regions():
    """Typed merge regions at odd byte offsets (an int at +6, a long at +3 and
    a double at +5 of their 16-byte blocks): many writers add into the same
    scalars at once; the 128-bit CAS keeps every contribution."""
    size = PAGE * 4
    orig = np.zeros(size, dtype=np.uint8)
    offs = {"int": 64 + 6, "long": 256 + 3, "double": 1024 + 5}
    orig[offs["int"] : offs["int"] + 4] = np.array([1000], dtype=np.int32).view(np.uint8)
    orig[offs["long"] : offs["long"] + 8] = np.array([1 << 40], dtype=np.int64).view(np.uint8)
    orig[offs["double"] : offs["double"] + 8] = np.array([2.5], dtype=np.float64).view(np.uint8)
    regs = snap.prepare_regions(
        [
            snap.MergeRegion(offs["int"], 4, snap.INT, snap.SUM),
            snap.MergeRegion(offs["long"], 8, snap.LONG, snap.SUM),
            snap.MergeRegion(offs["double"], 8, snap.DOUBLE, snap.SUM),
        ],
        size,
        "cuda",
    )
    n_writers = 8
    d_main = dev(orig)
    d_orig = dev(orig)
    d_mems = []
    for w in range(n_writers):
        m = orig.copy()
        m[offs["int"] : offs["int"] + 4] = np.array([1000 + (w + 1)], dtype=np.int32).view(np.uint8)
        m[offs["long"] : offs["long"] + 8] = np.array([(1 << 40) + 10 * (w + 1)], dtype=np.int64).view(np.uint8)
        m[offs["double"] : offs["double"] + 8] = np.array([2.5 + 0.25 * (w + 1)], dtype=np.float64).view(np.uint8)
        d_mems.append(dev(m))
    streams = [torch.cuda.Stream() for _ in range(n_writers)]
    torch.cuda.synchronize()
    for rep in range(3):
        d_main.copy_(d_orig)
        torch.cuda.synchronize()
        for w in range(n_writers):
            with torch.cuda.stream(streams[w]):
                snap.diff_push(d_mems[w], d_orig, d_main, regs)
        torch.cuda.synchronize()
        got = d_main.cpu().numpy()
        tot = n_writers * (n_writers + 1) // 2
        assert int(got[offs["int"] : offs["int"] + 4].view(np.int32)[0]) == 1000 + tot
        assert int(got[offs["long"] : offs["long"] + 8].view(np.int64)[0]) == (1 << 40) + 10 * tot
        assert float(got[offs["double"] : offs["double"] + 8].view(np.float64)[0]) == 2.5 + 0.25 * tot
        untouched = np.ones(size, dtype=bool)
        for k, ln in (("int", 4), ("long", 8), ("double", 8)):
            untouched[offs[k] : offs[k] + ln] = False
        assert not got[untouched].any()


def test_dirty_scan_and_flags_or():
    size = PAGE * 40 + 100
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, size, dtype=np.uint8)
    mem = base.copy()
    for p in (0, 7, 39, 40):
        mem[min(p * PAGE + 17, size - 1)] ^= 1
    flags, count = snap.dirty_scan(dev(mem), dev(base))
    torch.cuda.synchronize()
    f = flags.cpu().numpy()
    assert set(np.nonzero(f)[0].tolist()) == {0, 7, 39, 40}
    assert int(count.item()) == 4
    other = torch.zeros_like(flags)
    other[5] = 1
    snap.flags_or(flags, other)
    torch.cuda.synchronize()
    assert set(np.nonzero(flags.cpu().numpy())[0].tolist()) == {0, 5, 7, 39, 40}


def test_apply_diffs():
    size = 10000
    img = np.zeros(size, dtype=np.uint8)
    img[100:104] = np.array([10], dtype=np.int32).view(np.uint8)
    img[200:208] = np.array([2.5], dtype=np.float64).view(np.uint8)
    d_img = dev(img)
    diffs = [
        (10, snap.RAW, snap.BYTEWISE, bytes([1, 2, 3])),
        (50, snap.RAW, snap.XOR, bytes([0xFF, 0x0F])),
        (100, snap.INT, snap.SUM, np.array([5], dtype=np.int32).tobytes()),
        (200, snap.DOUBLE, snap.PRODUCT, np.array([4.0], dtype=np.float64).tobytes()),
        (300, snap.INT, snap.MAX, np.array([-3], dtype=np.int32).tobytes()),
        (400, snap.RAW, snap.IGNORE, bytes([9, 9])),
    ]
    snap.apply_diffs(d_img, diffs)
    torch.cuda.synchronize()
    got = d_img.cpu().numpy()
    assert got[10:13].tolist() == [1, 2, 3]
    assert got[50:52].tolist() == [0xFF, 0x0F]
    assert got[100:104].view(np.int32)[0] == 15
    assert got[200:208].view(np.float64)[0] == 10.0
    assert got[300:304].view(np.int32)[0] == 0
    assert got[400:402].tolist() == [0, 0]
