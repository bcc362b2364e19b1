"""Data-parallel gradient synchronisation — the flagship workload.

``GradientSync`` owns one flat symmetric-heap buffer holding every gradient
tensor of the model and all-reduces them with the semantics of **one
MPI_Allreduce per tensor** (the reference benchmark issues 214 of them per
pass).  Two execution modes:

* ``mode="grouped"`` (default): ONE fused peer-memory kernel per step walks a
  device-resident table of all tensors (``Communicator.prepare_group``): the
  cross-rank barriers are paid once per step instead of once per tensor.
* ``mode="lanes"``: one kernel per tensor, spread over `channels` concurrent
  lanes and replayed from a CUDA graph (the round-1 design, kept for comparison).

Public API used by bench.py / users:

    sync = GradientSync(comm, sizes, dtype=torch.int32)
    sync.step()                       # device-resident gradients
    loss = sync.step_from_host(pinned_host_grads)   # H2D + allreduce + D2H
"""

from __future__ import annotations

from typing import Optional, Sequence

import torch

from ..parallel import Communicator


# Candidate policies for algo="tuned": message bytes -> algorithm.  "auto" defers
# to the communicator's table/thresholds (measured one collective at a time);
# the others exist because the best choice under 8 concurrent lanes with few
# CTAs per kernel differs from the isolated sweep (NVLS is latency-bound with
# few CTAs, two-shot P2P is not).
def _policy_auto(nbytes, has_nvls):
    return "auto"


def _policy_twoshot(nbytes, has_nvls):
    if nbytes <= 32 * 1024:
        return "ll"
    return "twoshot"


def _policy_nvls_large(nbytes, has_nvls):
    if nbytes <= 32 * 1024:
        return "ll"
    if has_nvls and nbytes >= (4 << 20):
        return "nvls"
    return "twoshot"


POLICIES = {"auto": _policy_auto, "twoshot": _policy_twoshot, "nvls-large": _policy_nvls_large}


class GradientSync:
    def __init__(
        self,
        comm: Communicator,
        sizes: Sequence[int],
        dtype=torch.int32,
        op: str = "sum",
        algo: str = "auto",
        use_graph: bool = True,
        in_place: bool = False,
        channels: int = 8,
        bucket_bytes: int = 0,
        mode: str = "grouped",
    ):
        self.comm = comm
        self.sizes = [int(s) for s in sizes]
        self.dtype = dtype
        self.op = op
        self.algo = algo
        # stream-ordered synchronisation (ranks sharing a GPU) is driven from
        # the host and cannot be replayed from a graph
        self.use_graph = use_graph and not comm.stream_sync
        self.mode = mode
        esize = torch.empty((), dtype=dtype).element_size()
        # every tensor starts on a 256-byte boundary inside the flat buffers
        self.offsets = []
        cur = 0
        for s in self.sizes:
            self.offsets.append(cur)
            cur += (s * esize + 255) // 256 * 256 // esize
        self.total_padded = cur
        self.total_elems = sum(self.sizes)
        self.nbytes = self.total_elems * esize
        self.send = comm.empty(cur, dtype)
        self.recv = self.send if in_place else comm.empty(cur, dtype)
        self.send.zero_()
        if not in_place:
            self.recv.zero_()
        self.send_views = [self.send[o : o + s] for o, s in zip(self.offsets, self.sizes)]
        self.recv_views = [self.recv[o : o + s] for o, s in zip(self.offsets, self.sizes)]
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._stream = torch.cuda.Stream(device=comm.device)
        # independent tensors are all-reduced concurrently on `channels` lanes
        # (each lane = its own stream + its own slice of the barrier flags)
        self.channels = max(1, int(channels))
        self._lanes = [torch.cuda.Stream(device=comm.device) for _ in range(self.channels - 1)]
        self._fork = torch.cuda.Event()
        self._joins = [torch.cuda.Event() for _ in self._lanes]
        self._result = torch.zeros(2, dtype=torch.int64, device=f"cuda:{comm.device}")
        self._result_host = torch.zeros(2, dtype=torch.int64).pin_memory()
        # Optional DDP-style bucketing: neighbouring tensors of the flat buffer
        # are all-reduced together (one kernel per bucket, padding included).
        # bucket_bytes == 0 keeps the reference semantics: one call per tensor.
        self.bucket_bytes = int(bucket_bytes)
        if self.bucket_bytes > 0:
            spans = []
            start = 0
            end = 0
            for o, sz in zip(self.offsets, self.sizes):
                pad_end = o + (sz * esize + 255) // 256 * 256 // esize
                if end > start and (pad_end - start) * esize > self.bucket_bytes:
                    spans.append((start, end))
                    start = o
                end = pad_end
            spans.append((start, end))
            self._jobs = [(self.send[a:b], self.recv[a:b], b - a) for a, b in spans]
        else:
            self._jobs = [(s, r, n) for s, r, n in zip(self.send_views, self.recv_views, self.sizes)]
        self.launches_per_step = len(self._jobs)
        self._plan = None
        if self.mode == "grouped":
            self._plan = comm.prepare_group([j[0] for j in self._jobs], [j[1] for j in self._jobs])
            self.launches_per_step = self._plan.launches
        self._e2e = None
        # algo="tuned": measure the candidate policies in place (same lanes,
        # same CTA budget, same tensor mix) on first use and keep the fastest
        self.policy = None
        self.policy_timings = {}
        self._small_lanes = 0
        self.policy_name = None
        self._esize = esize
        self._chunks = None
        self._chunk_graphs = []

    # ------------------------------------------------------------------ core
    def _issue(self, stream, jobs=None):
        jobs = self._jobs if jobs is None else jobs
        if self.channels == 1:
            for s, r, n in jobs:
                self.comm.all_reduce(s, r, op=self.op, algo=self._algo_for(n), stream=stream)
            return
        # fork: lane streams wait for everything already queued on `stream`
        self._fork.record(stream)
        for lane in self._lanes:
            lane.wait_event(self._fork)
        # biggest tensors first, round-robin over the lanes.  With
        # `small_lanes` > 0 the latency-bound small messages get lanes of their
        # own so they run in the shadow of the bandwidth-bound large ones.
        order = sorted(range(len(jobs)), key=lambda i: -jobs[i][2])
        k_small = self._small_lanes if self.channels >= 4 else 0
        big_lanes = self.channels - k_small
        nb = ns = 0
        for i in order:
            nbytes = jobs[i][2] * self._esize
            if k_small > 0 and nbytes <= 32 * 1024:
                ch = big_lanes + (ns % k_small)
                ns += 1
            else:
                ch = nb % big_lanes
                nb += 1
            st = stream if ch == 0 else self._lanes[ch - 1]
            self.comm.all_reduce(
                jobs[i][0], jobs[i][1], op=self.op, algo=self._algo_for(jobs[i][2]), stream=st, channel=ch
            )
        # join
        for lane, ev in zip(self._lanes, self._joins):
            ev.record(lane)
            stream.wait_event(ev)

    def _algo_for(self, numel: int) -> str:
        if self.algo != "tuned":
            return self.algo
        fn = POLICIES[self.policy or "auto"]
        return fn(numel * self._esize, self.comm.has_multicast and self.comm.size >= 4)

    def _tune(self):
        """Time each policy for the whole step (CUDA graph, device events) and
        agree on the fastest across ranks (max over ranks, via an all-reduce)."""
        timings = {}
        graphs = {}
        candidates = []
        for name in POLICIES:
            if name == "nvls-large" and not (self.comm.has_multicast and self.comm.size >= 4):
                continue
            candidates.append((name, 0))
        # (Dedicated lanes for the small messages were measured too and lost:
        # 0.49-0.60 ms vs 0.44 ms at N=4; `_small_lanes` stays as a knob.)
        for name, k_small in candidates:
            self.policy = name
            self._small_lanes = k_small
            self._graph = None
            self._capture()
            g = self._graph
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(self._stream):
                g.replay()
                e0.record(self._stream)
                for _ in range(5):
                    g.replay()
                e1.record(self._stream)
            self._stream.synchronize()
            key = name if k_small == 0 else f"{name}+{k_small}small-lanes"
            timings[key] = e0.elapsed_time(e1) / 5
            graphs[key] = (g, name, k_small)
        names = list(timings)
        if self.comm.size > 1:
            # every rank must pick the same policy: compare the slowest rank
            t = self.comm.empty(len(names), torch.float32)
            t.copy_(torch.tensor([timings[n] for n in names], dtype=torch.float32))
            with torch.cuda.stream(self._stream):
                self.comm.all_reduce(t, t, op="max", stream=self._stream)
            self._stream.synchronize()
            agreed = t.cpu().tolist()
            self.comm.free(t)
            timings = dict(zip(names, agreed))
        best = min(names, key=lambda n: timings[n])
        self._graph, self.policy, self._small_lanes = graphs[best]
        self.policy_name = best
        self.policy_timings = {k: round(v, 4) for k, v in timings.items()}

    def _capture(self):
        # warm the launch path once outside capture (lazy module loading)
        torch.cuda.synchronize(self.comm.device)
        self._issue(self._stream)
        self._stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            self._issue(self._stream)
        self._graph = g

    def step(self, stream: Optional[torch.cuda.Stream] = None):
        """All-reduce every gradient tensor (async on `stream`)."""
        stream = stream or torch.cuda.current_stream(self.comm.device)
        if self._plan is not None:
            self.comm.all_reduce_group(self._plan, op=self.op, stream=stream)
            return
        if self.use_graph:
            if self._graph is None:
                if self.algo == "tuned" and self.policy is None:
                    self._tune()
                else:
                    self._capture()
            with torch.cuda.stream(stream):
                self._graph.replay()
        else:
            self._issue(stream)

    def _pipeline_chunks(self, k: int):
        """Split the jobs (in buffer order) into k contiguous chunks of similar
        byte size: [(elem_begin, elem_end, jobs)]."""
        spans = []
        for (s, r, n) in self._jobs:
            begin = s.data_ptr() - self.send.data_ptr()
            spans.append((begin // self.send.element_size(), s, r, n))
        spans.sort(key=lambda t: t[0])
        total = self.total_padded
        chunks = []
        cur = []
        cur_begin = 0
        for idx, (begin, s, r, n) in enumerate(spans):
            cur.append((s, r, n))
            nxt = spans[idx + 1][0] if idx + 1 < len(spans) else total
            if len(chunks) < k - 1 and nxt - cur_begin >= total / k:
                chunks.append((cur_begin, nxt, cur))
                cur = []
                cur_begin = nxt
        if cur:
            chunks.append((cur_begin, total, cur))
        return chunks

    def step_from_host(
        self,
        host_grads: torch.Tensor,
        stream: Optional[torch.cuda.Stream] = None,
        pipeline: int = 8,
        out_host: Optional[torch.Tensor] = None,
    ):
        """End-to-end step: copy this step's gradients from pinned host memory,
        all-reduce, and copy the reduced gradients back to pinned host memory
        (``out_host``, grouped mode: the FULL result; lanes mode: a digest).  The H2D copy is cut into `pipeline`
        chunks on a copy stream; the all-reduces of a chunk start as soon as it
        has landed, overlapping the rest of the transfer.  Returns the host
        digest tensor after synchronising the stream."""
        stream = stream or torch.cuda.current_stream(self.comm.device)
        if self._plan is not None:
            return self._step_from_host_grouped(host_grads, out_host, stream, pipeline)
        if self.algo == "tuned" and self.policy is None and self.use_graph:
            self._tune()
        if pipeline <= 1 or not self.use_graph:
            with torch.cuda.stream(stream):
                self.send[: host_grads.numel()].copy_(host_grads, non_blocking=True)
                self.step(stream)
        else:
            if self._chunks is None or len(self._chunks) != pipeline:
                self._chunks = self._pipeline_chunks(pipeline)
                self._chunk_graphs = [None] * len(self._chunks)
                self._copy_stream = torch.cuda.Stream(device=self.comm.device)
                self._chunk_events = [torch.cuda.Event() for _ in self._chunks]
            # copies are ordered after whatever the caller queued on `stream`
            self._copy_stream.wait_stream(stream)
            limit = host_grads.numel()
            with torch.cuda.stream(self._copy_stream):
                for (a, b, _), ev in zip(self._chunks, self._chunk_events):
                    hi = min(b, limit)
                    if hi > a:
                        self.send[a:hi].copy_(host_grads[a:hi], non_blocking=True)
                    ev.record(self._copy_stream)
            for k, ((a, b, jobs), ev) in enumerate(zip(self._chunks, self._chunk_events)):
                stream.wait_event(ev)
                if self._chunk_graphs[k] is None:
                    # warm, then capture this chunk's launch sequence
                    stream.synchronize()
                    self._issue(self._stream, jobs)
                    self._stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._stream):
                        self._issue(self._stream, jobs)
                    self._chunk_graphs[k] = g
                with torch.cuda.stream(stream):
                    self._chunk_graphs[k].replay()
        with torch.cuda.stream(stream):
            first = self.recv_views[0]
            self._result[0] = first[0].to(torch.int64)
            self._result[1] = first.to(torch.int64).sum()
            self._result_host.copy_(self._result, non_blocking=True)
        stream.synchronize()
        return self._result_host

    def _step_from_host_grouped(self, host_grads, out_host, stream, pipeline):
        """H2D -> grouped all-reduce -> D2H, cut into `pipeline` chunks of
        whole tensors on three streams so the PCIe copies of both directions
        overlap each other and the all-reduces of the neighbouring chunks."""
        dev = self.comm.device
        if self._e2e is None or self._e2e["k"] != pipeline:
            chunks = self._pipeline_chunks(max(1, pipeline))
            plans = [self.comm.prepare_group([j[0] for j in jobs], [j[1] for j in jobs]) for _, _, jobs in chunks]
            self._e2e = {
                "k": pipeline,
                "chunks": chunks,
                "plans": plans,
                "h2d": torch.cuda.Stream(device=dev),
                "d2h": torch.cuda.Stream(device=dev),
                "in_ev": [torch.cuda.Event() for _ in chunks],
                "red_ev": [torch.cuda.Event() for _ in chunks],
            }
        e = self._e2e
        if out_host is None:
            if e.get("out") is None:
                e["out"] = torch.empty(self.total_padded, dtype=self.dtype).pin_memory()
            out_host = e["out"]
        e["h2d"].wait_stream(stream)
        e["d2h"].wait_stream(stream)
        limit = host_grads.numel()
        with torch.cuda.stream(e["h2d"]):
            for (a, b, _), ev in zip(e["chunks"], e["in_ev"]):
                hi = min(b, limit)
                if hi > a:
                    self.send[a:hi].copy_(host_grads[a:hi], non_blocking=True)
                ev.record(e["h2d"])
        for (a, b, _), plan, iev, rev in zip(e["chunks"], e["plans"], e["in_ev"], e["red_ev"]):
            stream.wait_event(iev)
            self.comm.all_reduce_group(plan, op=self.op, stream=stream)
            rev.record(stream)
            with torch.cuda.stream(e["d2h"]):
                e["d2h"].wait_event(rev)
                out_host[a:b].copy_(self.recv[a:b], non_blocking=True)
        stream.wait_stream(e["d2h"])
        stream.synchronize()
        return out_host

    @property
    def h2d_bytes_per_step(self):
        return self.total_padded * self.send.element_size()

    @property
    def d2h_bytes_per_step(self):
        if self._plan is not None:
            return self.total_padded * self.send.element_size()
        return self._result_host.numel() * 8

    def close(self):
        if self._e2e is not None:
            for p in self._e2e["plans"]:
                p.close()
            self._e2e = None
        if self._plan is not None:
            self._plan.close()
            self._plan = None
        self._graph = None
        self._chunk_graphs = []
        if self.recv is not self.send:
            self.comm.free(self.recv)
        self.comm.free(self.send)
