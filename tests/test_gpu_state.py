"""GPU tests of the device-resident state kernels against a NumPy oracle
(reference semantics: StateKeyValue dirty mask -> chunks -> push,
src/state/StateKeyValue.cpp:441-543,592-629)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from faabric_b200.ops import state as dstate  # noqa: E402


@pytest.mark.parametrize("nbytes", [1, 127, 128, 129, 4096, 1 << 20, (3 << 20) + 777])
def test_push_dirty_copies_exactly_the_flagged_blocks(nbytes):
    dev = "cuda:0"
    rng = np.random.default_rng(nbytes)
    src_h = rng.integers(0, 255, nbytes, dtype=np.uint8)
    dst_h = rng.integers(0, 255, nbytes, dtype=np.uint8)
    src = torch.from_numpy(src_h).to(dev)
    dst = torch.from_numpy(dst_h.copy()).to(dev)
    mask = dstate.new_mask(src)
    n_blocks = mask.numel()
    # random ranges, including the first and the last byte
    ranges = [(0, 1), (nbytes - 1, 1)]
    for _ in range(12):
        off = int(rng.integers(0, nbytes))
        ln = int(rng.integers(1, max(2, min(5000, nbytes - off + 1))))
        ranges.append((off, min(ln, nbytes - off)))
    expect_blocks = np.zeros(n_blocks, dtype=bool)
    for off, ln in ranges:
        dstate.flag_range(mask, off, ln)
        expect_blocks[off // dstate.BLOCK : (off + ln - 1) // dstate.BLOCK + 1] = True
    torch.cuda.synchronize()
    assert np.array_equal(mask.cpu().numpy().astype(bool), expect_blocks)
    stats = dstate.push_dirty(mask, src, dst)
    torch.cuda.synchronize()
    assert int(stats[0]) == int(expect_blocks.sum())
    oracle = dst_h.copy()
    for b in np.nonzero(expect_blocks)[0]:
        lo, hi = b * dstate.BLOCK, min(nbytes, (b + 1) * dstate.BLOCK)
        oracle[lo:hi] = src_h[lo:hi]
    assert np.array_equal(dst.cpu().numpy(), oracle)
    assert int(mask.sum()) == 0
    # nothing left to push
    stats2 = dstate.push_dirty(mask, src, dst)
    torch.cuda.synchronize()
    assert int(stats2[0]) == 0


def test_push_dirty_all_blocks_runs_at_copy_speed():
    dev = "cuda:0"
    n = 256 << 20
    src = torch.full((n,), 7, dtype=torch.uint8, device=dev)
    dst = torch.zeros(n, dtype=torch.uint8, device=dev)
    mask = dstate.new_mask(src)
    mask.fill_(1)
    st = dstate.push_dirty(mask, src, dst)
    torch.cuda.synchronize()
    assert int(st[0]) == n // dstate.BLOCK
    assert bool((dst == 7).all())
