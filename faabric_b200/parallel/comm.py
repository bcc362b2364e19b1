"""Python face of the device Communicator (csrc/src/device/communicator.cpp).

One :class:`Communicator` per MPI rank / GPU.  Two ways to get one:

* :func:`init_from_env` — one process per GPU (torchrun): rank, world size and
  local rank come from RANK / WORLD_SIZE / LOCAL_RANK; peer memory is wired
  through the native Unix-socket bootstrap (VMM fds or CUDA IPC handles).
* :class:`LocalGroup` — all ranks inside this process (the reference's
  rank-thread model); several ranks may share one GPU, which is how the
  single-GPU test-suite exercises the cross-rank flag protocol.

All collectives are stream-ordered kernel launches (no host sync) and fuse the
reduce op; see csrc/kernels.  Tensors allocated with :meth:`Communicator.empty`
live in the symmetric heap and take the zero-copy paths.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Sequence

import torch

from .. import _lib
from .._lib import FbConfig

# FbDtype
_DTYPES = {
    torch.int8: 0,
    torch.uint8: 1,
    torch.int16: 2,
    torch.int32: 4,
    torch.int64: 6,
    torch.float32: 8,
    torch.float64: 9,
    torch.float16: 10,
    torch.bfloat16: 11,
    torch.bool: 1,
}
for _name, _code in (("uint16", 3), ("uint32", 5), ("uint64", 7)):
    if hasattr(torch, _name):
        _DTYPES[getattr(torch, _name)] = _code

OPS = {
    "max": 0,
    "min": 1,
    "sum": 2,
    "prod": 3,
    "land": 4,
    "lor": 5,
    "band": 6,
    "bor": 7,
    "maxloc": 8,
    "minloc": 9,
    "lxor": 10,
    "bxor": 11,
}
ALGOS = {"auto": 0, "oneshot": 1, "twoshot": 2, "nvls": 3, "ll": 4}
ALGO_NAMES = {v: k for k, v in ALGOS.items()}
FLAG_SYMMETRIC = 1
FLAG_NOSYNC = 2


class CommError(RuntimeError):
    pass


class _CudaBuf:
    """Minimal __cuda_array_interface__ carrier for heap memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 3,
        }


def make_config(**kw) -> FbConfig:
    lib = _lib.load()
    cfg = FbConfig()
    lib.fb_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, int(v))
    return cfg


class Communicator:
    def __init__(self, handle: int, owner=None):
        self._lib = _lib.load()
        self._h = C.c_void_p(handle)
        self._owner = owner  # keeps a LocalGroup alive
        self.rank = self._lib.fb_comm_rank(self._h)
        self.size = self._lib.fb_comm_size(self._h)
        self.device = self._lib.fb_comm_device(self._h)
        self.backing = self._lib.fb_comm_backing(self._h).decode()
        self.has_multicast = bool(self._lib.fb_comm_has_multicast(self._h))
        # cross-rank synchronisation through stream memory operations instead
        # of in-kernel spins (ranks sharing a GPU): not CUDA-graph capturable
        self.stream_sync = bool(self._lib.fb_comm_stream_sync(self._h))
        self.stream_wait_supported = bool(self._lib.fb_comm_stream_wait_supported(self._h))
        self._allocs: dict[int, int] = {}

    # ------------------------------------------------------------ lifecycle
    def close(self):
        if self._h:
            self._lib.fb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------- helpers
    def _stream(self, stream) -> C.c_void_p:
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        return C.c_void_p(stream.cuda_stream)

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise CommError(
                f"{what} failed: {self._lib.fb_error_string(rc).decode()} "
                f"[{_lib.last_error()}]"
            )

    def _sym(self, *tensors) -> int:
        for t in tensors:
            if t is None:
                continue
            if not self._lib.fb_comm_in_heap(
                self._h, C.c_void_p(t.data_ptr()), max(t.numel() * t.element_size(), 1)
            ):
                return 0
        return FLAG_SYMMETRIC

    @staticmethod
    def _dtype(t: torch.Tensor) -> int:
        try:
            return _DTYPES[t.dtype]
        except KeyError:
            raise CommError(f"unsupported dtype {t.dtype}")

    # ------------------------------------------------------------ heap
    def empty(self, shape, dtype=torch.float32) -> torch.Tensor:
        """Allocate a tensor in the symmetric heap.  Call collectively (same
        order and sizes on every rank) so offsets match across ranks."""
        if isinstance(shape, int):
            shape = (shape,)
        numel = 1
        for s in shape:
            numel *= int(s)
        esize = torch.empty((), dtype=dtype).element_size()
        nbytes = max(numel * esize, 16)
        off = self._lib.fb_comm_alloc(self._h, nbytes)
        if off < 0:
            raise CommError(f"symmetric heap exhausted ({nbytes} bytes)")
        ptr = self._lib.fb_comm_heap_ptr(self._h, off, -1)
        with torch.cuda.device(self.device):
            flat = torch.as_tensor(_CudaBuf(ptr, nbytes), device=f"cuda:{self.device}")
        t = flat[: numel * esize].view(dtype).view(*shape)
        self._allocs[t.data_ptr()] = off
        return t

    def zeros(self, shape, dtype=torch.float32) -> torch.Tensor:
        t = self.empty(shape, dtype)
        t.zero_()
        return t

    def free(self, t: torch.Tensor):
        off = self._allocs.pop(t.data_ptr(), None)
        if off is not None:
            self._lib.fb_comm_free(self._h, off)

    def heap_offset(self, t: torch.Tensor) -> int:
        base = self._lib.fb_comm_heap_ptr(self._h, 0, -1)
        return t.data_ptr() - base

    def set_allreduce_table(self, table):
        """table: [(max_bytes, algo_name), ...] measured by the autotuner."""
        n = len(table)
        mb = (C.c_uint64 * n)(*[int(t[0]) for t in table])
        al = (C.c_int * n)(*[ALGOS[t[1]] for t in table])
        self._check(self._lib.fb_comm_set_allreduce_table(self._h, n, mb, al), "set table")

    def load_tuning(self, path):
        """Apply a tuning file (see ``faabric_b200.parallel.autotune``)."""
        self._check(self._lib.fb_comm_load_tuning(self._h, str(path).encode()), f"load tuning {path}")

    def configure(self, **kw):
        keys = {
            "llMaxBytes": 0,
            "oneShotMaxBytes": 1,
            "nvlsMinBytes": 2,
            "bcast2StepMinBytes": 3,
            "maxBlocks": 4,
            "threads": 5,
            "tmaMinBytes": 6,
            "nvlsScalarMinBytes": 7,
            "groupBlocks": 8,
        }
        for k, v in kw.items():
            self._check(self._lib.fb_comm_configure(self._h, keys[k], int(v)), k)

    def stats(self, reset: bool = False) -> dict:
        out = (C.c_uint64 * 16)()
        self._lib.fb_comm_stats(self._h, out, 1 if reset else 0)
        d = {"launches": out[0], "bytes": out[1], "staged_copies": out[2], "tma_launches": out[15]}
        for code, name in ALGO_NAMES.items():
            d[f"algo_{name}"] = out[3 + code]
        return d

    @property
    def last_algo(self) -> str:
        return ALGO_NAMES.get(self._lib.fb_comm_last_algo(self._h), "?")

    def check_error(self, stream=None) -> int:
        return int(self._lib.fb_comm_check_error(self._h, self._stream(stream)))

    def host_barrier(self):
        self._lib.fb_comm_host_barrier(self._h)

    # ------------------------------------------------------------ collectives
    def all_reduce(self, send, recv=None, op="sum", algo="auto", stream=None, flags=None, channel=0):
        recv = send if recv is None else recv
        # only the source needs to be symmetric; the native side stages the
        # destination when a push algorithm needs it
        f = self._sym(send) if flags is None else flags
        f |= (channel & 0xF) << 8
        rc = self._lib.fb_allreduce(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(recv.data_ptr()),
            send.numel(),
            self._dtype(send),
            OPS[op],
            ALGOS[algo],
            f,
            self._stream(stream),
        )
        self._check(rc, "all_reduce")
        return recv

    def reduce(self, send, recv, root=0, op="sum", stream=None, flags=None):
        f = self._sym(send) if flags is None else flags
        rp = recv.data_ptr() if recv is not None else 0
        rc = self._lib.fb_reduce(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(rp),
            send.numel(),
            self._dtype(send),
            OPS[op],
            root,
            f,
            self._stream(stream),
        )
        self._check(rc, "reduce")
        return recv

    def reduce_scatter(self, send, recv, op="sum", stream=None, flags=None):
        f = self._sym(send) if flags is None else flags
        rc = self._lib.fb_reduce_scatter(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(recv.data_ptr()),
            recv.numel(),
            self._dtype(send),
            OPS[op],
            f,
            self._stream(stream),
        )
        self._check(rc, "reduce_scatter")
        return recv

    def scan(self, send, recv, op="sum", stream=None, flags=None):
        f = self._sym(send) if flags is None else flags
        rc = self._lib.fb_scan(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(recv.data_ptr()),
            send.numel(),
            self._dtype(send),
            OPS[op],
            f,
            self._stream(stream),
        )
        self._check(rc, "scan")
        return recv

    def broadcast(self, buf, root=0, stream=None, flags=None):
        f = self._sym(buf) if flags is None else flags
        rc = self._lib.fb_broadcast(
            self._h,
            C.c_void_p(buf.data_ptr()),
            buf.numel() * buf.element_size(),
            root,
            f,
            self._stream(stream),
        )
        self._check(rc, "broadcast")
        return buf

    def all_gather(self, send, recv, stream=None, flags=None):
        # pull only needs the source symmetric; the native side additionally
        # checks the destination before taking the NVLS path
        f = self._sym(send) if flags is None else flags
        rc = self._lib.fb_allgather(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(recv.data_ptr()),
            send.numel() * send.element_size(),
            f,
            self._stream(stream),
        )
        self._check(rc, "all_gather")
        return recv

    def gather(self, send, recv, root=0, stream=None, flags=None):
        f = self._sym(send) if flags is None else flags
        rp = recv.data_ptr() if recv is not None else 0
        rc = self._lib.fb_gather(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(rp),
            send.numel() * send.element_size(),
            root,
            f,
            self._stream(stream),
        )
        self._check(rc, "gather")
        return recv

    def scatter(self, send, recv, root=0, stream=None, flags=None):
        # `send` is only significant on the root.  The zero-copy path is taken
        # when EVERY rank passes a symmetric `send` (SPMD style); otherwise the
        # root stages through the fixed staging area.
        f = flags
        if f is None:
            f = self._sym(send) if send is not None else 0
        sp = send.data_ptr() if send is not None else 0
        rc = self._lib.fb_scatter(
            self._h,
            C.c_void_p(sp),
            C.c_void_p(recv.data_ptr()),
            recv.numel() * recv.element_size(),
            root,
            f,
            self._stream(stream),
        )
        self._check(rc, "scatter")
        return recv

    def all_to_all(self, send, recv, stream=None, flags=None):
        f = self._sym(send) if flags is None else flags
        rc = self._lib.fb_alltoall(
            self._h,
            C.c_void_p(send.data_ptr()),
            C.c_void_p(recv.data_ptr()),
            send.numel() * send.element_size() // self.size,
            f,
            self._stream(stream),
        )
        self._check(rc, "all_to_all")
        return recv

    # ------------------------------------------------- grouped all-reduce
    def _group_arrays(self, sends, recvs):
        n = len(sends)
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in sends])
        rp = (C.c_void_p * n)(*[t.data_ptr() for t in recvs])
        cnt = (C.c_uint64 * n)(*[t.numel() for t in sends])
        return n, sp, rp, cnt

    def prepare_group(self, sends, recvs=None) -> "GroupPlan":
        """Plan ONE launch that all-reduces every tensor of ``sends`` (each with
        per-tensor semantics).  Tensors must live in the symmetric heap at
        16-byte aligned addresses; call collectively with the same list."""
        recvs = sends if recvs is None else recvs
        if len(sends) != len(recvs) or not sends:
            raise CommError("prepare_group: need equally long, non-empty lists")
        dt = self._dtype(sends[0])
        n, sp, rp, cnt = self._group_arrays(sends, recvs)
        h = self._lib.fb_group_prepare(self._h, n, sp, rp, cnt, dt)
        if not h:
            raise CommError(f"prepare_group failed [{_lib.last_error()}]")
        return GroupPlan(self, h, n, sum(t.numel() * t.element_size() for t in sends))

    def all_reduce_group(self, plan: "GroupPlan", op="sum", stream=None, channel=0, flags=FLAG_SYMMETRIC):
        rc = self._lib.fb_group_allreduce(
            self._h, plan._h, OPS[op], flags | ((channel & 0xF) << 8), self._stream(stream)
        )
        self._check(rc, "all_reduce_group")

    def all_reduce_many(self, sends, recvs=None, op="sum", stream=None, channel=0):
        """Transient variant of :meth:`prepare_group` + :meth:`all_reduce_group`
        (the table is rebuilt and uploaded in stream order on every call)."""
        recvs = sends if recvs is None else recvs
        dt = self._dtype(sends[0])
        n, sp, rp, cnt = self._group_arrays(sends, recvs)
        f = self._sym(*sends) | ((channel & 0xF) << 8)
        rc = self._lib.fb_allreduce_many(self._h, n, sp, rp, cnt, dt, OPS[op], f, self._stream(stream))
        self._check(rc, "all_reduce_many")

    def send_recv(self, send_buf, dst, recv_buf, src, stream=None):
        rc = self._lib.fb_sendrecv(
            self._h,
            C.c_void_p(send_buf.data_ptr()),
            send_buf.numel() * send_buf.element_size(),
            dst,
            C.c_void_p(recv_buf.data_ptr()),
            recv_buf.numel() * recv_buf.element_size(),
            src,
            self._stream(stream),
        )
        self._check(rc, "send_recv")

    def synchronize(self, stream=None, timeout_ms: int = 30000) -> bool:
        """Bounded wait for ``stream``; False if it had to be aborted."""
        return bool(self._lib.fb_comm_sync_bounded(self._h, self._stream(stream), int(timeout_ms)))

    def barrier(self, stream=None):
        self._check(self._lib.fb_barrier(self._h, self._stream(stream)), "barrier")

    def send(self, buf, peer, stream=None):
        rc = self._lib.fb_send(
            self._h,
            C.c_void_p(buf.data_ptr()),
            buf.numel() * buf.element_size(),
            peer,
            self._stream(stream),
        )
        self._check(rc, "send")

    def recv(self, buf, peer, stream=None):
        rc = self._lib.fb_recv(
            self._h,
            C.c_void_p(buf.data_ptr()),
            buf.numel() * buf.element_size(),
            peer,
            self._stream(stream),
        )
        self._check(rc, "recv")

    def put_signal(self, local, dst_sym, peer, signal=0, blocks=8, stream=None):
        """Copy `local` into the peer's copy of symmetric tensor `dst_sym` and
        bump its user signal `signal` (once per CTA)."""
        rc = self._lib.fb_put_signal(
            self._h,
            C.c_void_p(local.data_ptr()),
            self.heap_offset(dst_sym),
            local.numel() * local.element_size(),
            peer,
            signal,
            blocks,
            self._stream(stream),
        )
        self._check(rc, "put_signal")

    def wait_signal(self, signal=0, count=8, stream=None):
        self._check(
            self._lib.fb_wait_signal(self._h, signal, count, self._stream(stream)),
            "wait_signal",
        )


class GroupPlan:
    """Device-resident segment tables of a grouped all-reduce."""

    def __init__(self, comm: Communicator, handle, n_tensors: int, nbytes: int):
        self._comm = comm
        self._h = C.c_void_p(handle)
        self.n_tensors = n_tensors
        self.nbytes = nbytes
        self.launches = int(comm._lib.fb_group_plan_launches(self._h))

    def close(self):
        if self._h:
            self._comm._lib.fb_group_plan_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalGroup:
    """N ranks inside this process.  ``devices[i]`` is rank i's GPU; repeating a
    device id puts several ranks on one GPU (each rank then needs its own
    stream so the per-rank kernels are co-resident)."""

    def __init__(self, nranks: int, devices: Optional[Sequence[int]] = None, **cfg):
        lib = _lib.load()
        ndev = lib.fb_cuda_device_count()
        if ndev <= 0:
            raise CommError("no CUDA device")
        if devices is None:
            devices = [i % ndev for i in range(nranks)]
        self.devices = list(devices)
        arr = (C.c_int * nranks)(*self.devices)
        c = make_config(**cfg)
        h = lib.fb_group_create_local(nranks, arr, C.byref(c))
        if not h:
            raise CommError(f"group creation failed: {_lib.last_error()}")
        self._lib = lib
        self._h = C.c_void_p(h)
        self.comms = [
            Communicator(lib.fb_group_comm(self._h, r), owner=self)
            for r in range(nranks)
        ]
        self.streams = []
        for d in self.devices:
            with torch.cuda.device(d):
                self.streams.append(torch.cuda.Stream(device=d))
        self.size = nranks

    def run(self, fn: Callable[[Communicator, int, "torch.cuda.Stream"], object]):
        """Issue ``fn(comm, rank, stream)`` for every rank, each on its own
        stream, without synchronising in between (launches are asynchronous, so
        the per-rank kernels overlap on the device(s))."""
        out = []
        for r, c in enumerate(self.comms):
            with torch.cuda.device(c.device):
                # Rank streams are non-blocking: order them after whatever the
                # caller has queued on the device's current stream (e.g. the
                # copies that filled the input buffers)
                self.streams[r].wait_stream(torch.cuda.current_stream(c.device))
                with torch.cuda.stream(self.streams[r]):
                    out.append(fn(c, r, self.streams[r]))
        return out

    def synchronize(self):
        for s in self.streams:
            s.synchronize()

    def check_errors(self):
        return [c.check_error(self.streams[r]) for r, c in enumerate(self.comms)]

    @property
    def shares_devices(self) -> bool:
        return len(set(self.devices)) < len(self.devices)

    def coresident(self) -> bool:
        """Probe for groups that synchronise INSIDE kernels while several ranks
        share a GPU: one barrier kernel per rank must meet on the device.  False
        (and a poisoned group: close it) when the ranks' kernels do not overlap,
        e.g. their streams alias one hardware queue or a tool serialises them."""
        self.run(lambda c, r, st: c.barrier())
        return self.check_errors() == [0] * self.size

    def close(self):
        for c in self.comms:
            c.close()
        self.comms = []
        if self._h:
            self._lib.fb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def init_from_env(**cfg) -> Communicator:
    """One-process-per-GPU initialisation (torchrun environment)."""
    lib = _lib.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    job = os.environ.get("FAABRIC_JOB_ID") or (
        os.environ.get("MASTER_PORT", "0") + "-" + os.environ.get("TORCHELASTIC_RUN_ID", "x")
    )
    torch.cuda.set_device(local)
    c = make_config(**cfg)
    h = lib.fb_comm_create_ipc(rank, world, local, job.encode(), C.byref(c))
    if not h:
        raise CommError(f"ipc communicator creation failed: {_lib.last_error()}")
    return Communicator(h)
