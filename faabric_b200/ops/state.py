"""Device-resident state values: the dirty-block mask and the fused
scan + push + clear kernel (csrc/kernels/state_kernels.cu).

The C++ runtime uses these through ``faabric::state::DeviceStateKeyValue``
(csrc/src/state/device_state.cpp); this module exposes the same kernels on
torch tensors.  Reference counterpart: ``StateKeyValue::getDirtyChunks`` +
``pushPartial`` (src/state/StateKeyValue.cpp:441-543,592-629), a byte-mask scan
on the CPU followed by one RPC per 64 KiB chunk.
"""

from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

BLOCK = 128  # FB_STATE_BLOCK_BYTES: bytes of value covered by one mask byte


def _stream(device, stream):
    if stream is None:
        stream = torch.cuda.current_stream(device)
    return C.c_void_p(stream.cuda_stream)


def new_mask(value: torch.Tensor) -> torch.Tensor:
    nbytes = value.numel() * value.element_size()
    return torch.zeros((nbytes + BLOCK - 1) // BLOCK, dtype=torch.uint8, device=value.device)


def flag_range(mask: torch.Tensor, offset: int, length: int, stream=None):
    """Mark bytes [offset, offset+length) of the value as written."""
    rc = _lib.load().fb_state_flag_range(C.c_void_p(mask.data_ptr()), int(offset), int(length), _stream(mask.device, stream))
    if rc != 0:
        raise RuntimeError("fb_state_flag_range failed")


def push_dirty(mask: torch.Tensor, src: torch.Tensor, dst, stats: torch.Tensor | None = None, blocks: int = 0, stream=None):
    """Copy every dirty 128-byte block of ``src`` to ``dst`` (tensor or raw,
    possibly peer-mapped, device pointer) and clear the mask.  Returns the
    int64[2] stats tensor (stats[0] = dirty blocks)."""
    dev = src.device
    if stats is None:
        stats = torch.zeros(2, dtype=torch.int64, device=dev)
    dptr = dst.data_ptr() if isinstance(dst, torch.Tensor) else int(dst)
    rc = _lib.load().fb_state_push_dirty(
        C.c_void_p(mask.data_ptr()),
        C.c_void_p(src.data_ptr()),
        C.c_void_p(dptr),
        src.numel() * src.element_size(),
        C.c_void_p(stats.data_ptr()),
        int(blocks),
        _stream(dev, stream),
    )
    if rc != 0:
        raise RuntimeError("fb_state_push_dirty failed")
    return stats
