"""CPU-side benchmarks of the native runtime (BASELINE.md configs 1 and 5, and
the `refcpu` comparator: the reference's host-path design run on this box)."""

from __future__ import annotations

import json
import subprocess

from .cluster import BINDIR, LocalCluster


def planner_fanout_bench(n_functions: int = 1024, n_hosts: int = 8, iters: int = 20, warmup: int = 3, mode: str = "native") -> dict:
    """Batch-schedule `n_functions` no-ops over `n_hosts` (virtual GPU) hosts,
    fan-out + fan-in, planner and worker in one process.  ``mode="refcpu"``
    runs the reference's control-plane design on the same box: every request
    and result encoded, sent over a loopback socket and decoded."""
    exe = BINDIR / "planner_bench"
    if not exe.exists():
        from .. import build as _build

        _build.build(verbose=False)
    r = subprocess.run(
        [str(exe), "--functions", str(n_functions), "--hosts", str(n_hosts), "--iters", str(iters), "--warmup", str(warmup), "--mode", mode],
        capture_output=True,
        text=True,
        timeout=600,
    )
    if r.returncode != 0:
        raise RuntimeError(f"planner_bench failed: {r.stdout[-500:]} {r.stderr[-500:]}")
    res = json.loads(r.stdout.strip().splitlines()[-1])
    res["us_per_batch_median"] = res["e2e_us_median"]
    return res


def threads_forkjoin_bench(memory: str = "device", hosts: int = 0, mem_mb: int = 1024, dirty_pct: float = 1.0,
                           iters: int = 10, warmup: int = 2) -> dict:
    """THREADS fork-join through planner, scheduler, executors and the snapshot
    registry: one thread per (virtual GPU) host, each dirtying its share of a
    `mem_mb` function memory, merged back at the join.  memory="device": HBM
    images, peer-copy restore, one fused diff+push kernel per host;
    memory="host": the reference's mprotect + byte-diff design on this box."""
    exe = BINDIR / "threads_bench"
    if not exe.exists():
        from .. import build as _build

        _build.build(verbose=False)
    r = subprocess.run(
        [str(exe), "--memory", memory, "--hosts", str(hosts), "--mem-mb", str(mem_mb), "--dirty-pct", str(dirty_pct),
         "--iters", str(iters), "--warmup", str(warmup)],
        capture_output=True,
        text=True,
        timeout=900,
    )
    if r.returncode != 0:
        raise RuntimeError(f"threads_bench failed: {r.stdout[-500:]} {r.stderr[-800:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


def _first_output(status: dict) -> dict:
    for m in sorted(status["messageResults"], key=lambda m: m.get("mpiRank", 0)):
        if m.get("output_data"):
            return json.loads(m["output_data"])
    raise RuntimeError(f"no output in {status}")


def cpu_pingpong_bench(sizes=(8, 1024, 65536), n_workers: int = 1) -> list[dict]:
    """MPI ping-pong between two ranks on host memory.  n_workers=1: both ranks
    in one worker (in-memory queues); 2: one rank per worker process (TCP)."""
    out = []
    with LocalCluster(n_workers=n_workers, slots_per_worker=2 // n_workers, log_level="warn") as c:
        for s in sizes:
            st = c.client.invoke("mpi", "bench-pingpong", mpi_world_size=2, input_data=str(s), timeout=300)
            res = _first_output(st)
            res["transport"] = "local-queue" if n_workers == 1 else "tcp"
            out.append(res)
    return out


def mpi_allreduce_bench(
    counts: list[int],
    world_size: int,
    steps: int = 3,
    warmup: int = 1,
    memory: str = "host",
    host_algo: str = "shared",
) -> dict:
    """The headline workload (one MPI_Allreduce per gradient tensor) through
    the MPI C API of a worker process.  memory="host": reduce-to-root +
    broadcast over in-memory queues, malloc+memcpy per hop - the design the
    reference ships.  memory="device": buffers in HBM, each call one fused
    P2P/NVLS kernel (needs GPUs)."""
    payload = f"{steps};{warmup};{memory};" + ",".join(str(int(c)) for c in counts)
    # host_algo="reference": reduce to rank 0 + broadcast of malloc'ed copies
    # (what the reference does); "shared": slice-parallel in shared memory
    env = {"FAABRIC_MPI_HOST_ALLREDUCE": host_algo}
    with LocalCluster(n_workers=1, slots_per_worker=world_size, log_level="warn", extra_env=env) as c:
        st = c.client.invoke("mpi", "bench-allreduce-list", mpi_world_size=world_size, input_data=payload, timeout=1800)
        bad = [m for m in st["messageResults"] if m.get("returnValue", 0) != 0]
        if bad:
            raise RuntimeError(f"refcpu ranks failed: {bad[:2]}")
        return _first_output(st)


def cpu_allreduce_bench(counts: list[int], world_size: int, steps: int = 3, warmup: int = 1) -> dict:
    """The `refcpu` baseline: the reference's algorithm on host memory."""
    return mpi_allreduce_bench(counts, world_size, steps, warmup, memory="host", host_algo="reference")


def host_collectives_bench(ranks=(4, 8), sizes=(65536, 1 << 20, 8 << 20), repeats: int = 3, calls: int = 10) -> dict:
    """Host-buffer collectives with every rank in one worker: the reference's
    message algorithms (FAABRIC_MPI_HOST_ALLREDUCE=reference) against the
    shared-memory path.  Microseconds per call, minimum over `repeats` runs."""
    best: dict = {}
    for _ in range(repeats):
        for mode in ("shared", "reference"):
            with LocalCluster(
                n_workers=1, slots_per_worker=max(ranks), log_level="warn", extra_env={"FAABRIC_MPI_HOST_ALLREDUCE": mode}
            ) as c:
                for n in ranks:
                    for b in sizes:
                        st = c.client.invoke("mpi", "bench-collectives", mpi_world_size=n, input_data=f"{b},{calls}", timeout=600)
                        out = _first_output(st)
                        cell = best.setdefault(f"{mode}-{n}-{b}", out)
                        for k, v in out.items():
                            if k.endswith("_us"):
                                cell[k] = min(cell[k], v)
    return best

