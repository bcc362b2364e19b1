"""Single-box cluster of native processes for tests and CPU baselines."""

from __future__ import annotations

import os
import signal
import socket
import subprocess
import time
from pathlib import Path

from .client import PlannerHttpClient

ROOT = Path(__file__).resolve().parents[2]
BINDIR = ROOT / "build" / "bin"


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


_instances = 0
# port slots of clusters that are up right now (a long-lived fixture must not
# be hit by the rotation coming round again)
_live_slots: set = set()
_N_SLOTS = 5


class LocalCluster:
    """planner_server + ``n_workers`` faabric_worker processes.

    Workers share 127.0.0.1 and are told apart by FAABRIC_PORT_OFFSET, so a
    worker's host name is ``127.0.0.1:<offset>``.
    """

    def __init__(
        self,
        n_workers: int = 2,
        slots_per_worker: int = 2,
        *,
        log_level: str = "info",
        extra_env: dict | None = None,
        log_dir: str | os.PathLike | None = None,
        base_offset: int | None = None,
    ):
        self.n_workers = n_workers
        self.slots = slots_per_worker
        self.log_level = log_level
        self.extra_env = dict(extra_env or {})
        self.log_dir = Path(log_dir) if log_dir else None
        # Keep concurrent clusters apart: other processes (xdist) by pid, other
        # instances in this process by a counter.  Offsets stay far below the
        # ephemeral port range (highest port = 8100 + offset + 100 * workers)
        global _instances
        self._slot = None
        if base_offset is None:
            # (everything stays below 32768, the start of the ephemeral range:
            # a listening port there can collide with an outgoing connection)
            for probe in range(_N_SLOTS):
                slot = (_instances + probe) % _N_SLOTS
                if slot not in _live_slots:
                    break
            else:
                raise RuntimeError(f"more than {_N_SLOTS} LocalClusters alive in one process")
            self._slot = slot
            _live_slots.add(slot)
            base_offset = 1000 + (os.getpid() % 7) * 3000 + slot * 600
        _instances += 1
        self.base_offset = base_offset
        self.procs: list[subprocess.Popen] = []
        self.http_port = _free_port()
        self.client = PlannerHttpClient("127.0.0.1", self.http_port)

    # ------------------------------------------------------------------
    def _env(self, offset: int, slots: int | None) -> dict:
        env = dict(os.environ)
        env.update(
            {
                "LOG_LEVEL": self.log_level,
                "ENDPOINT_HOST": "127.0.0.1",
                "PLANNER_HOST": f"127.0.0.1:{self.base_offset}",
                "PLANNER_PORT": str(self.http_port),
                "FAABRIC_PORT_OFFSET": str(offset),
                # a killed test runner must not leave servers behind
                "FAABRIC_EXIT_WITH_PARENT": "1",
                "FAABRIC_DEVICE_BACKEND": os.environ.get("FAABRIC_DEVICE_BACKEND", "cuda"),
            }
        )
        if slots is not None:
            env["OVERRIDE_CPU_COUNT"] = str(slots)
        env.update(self.extra_env)
        return env

    def _spawn(self, binary: str, offset: int, slots: int | None, tag: str) -> subprocess.Popen:
        exe = BINDIR / binary
        if not exe.exists():
            from .. import build as _build

            _build.build(verbose=False)
        out = subprocess.DEVNULL
        if self.log_dir:
            self.log_dir.mkdir(parents=True, exist_ok=True)
            out = open(self.log_dir / f"{tag}.log", "w")
        p = subprocess.Popen([str(exe)], env=self._env(offset, slots), stdout=out, stderr=subprocess.STDOUT, start_new_session=True)
        self.procs.append(p)
        return p

    def start(self, timeout: float = 30.0) -> "LocalCluster":
        self._spawn("planner_server", self.base_offset, None, "planner")
        deadline = time.time() + timeout
        while True:
            try:
                self.client.available_hosts()
                break
            except Exception:
                if time.time() > deadline:
                    self.stop()
                    raise RuntimeError("planner did not come up")
                time.sleep(0.05)
        for i in range(self.n_workers):
            self._spawn("faabric_worker", self.base_offset + 100 * (i + 1), self.slots, f"worker{i}")
        while len(self.client.available_hosts()) < self.n_workers:
            if time.time() > deadline:
                self.stop()
                raise RuntimeError("workers did not register")
            for p in self.procs:
                if p.poll() is not None:
                    self.stop()
                    raise RuntimeError(f"process {p.args} exited with {p.returncode}")
            time.sleep(0.05)
        return self

    def add_worker(self, slots: int | None = None, timeout: float = 30.0) -> str:
        """Start one more worker while the cluster runs; returns its host name
        once the planner lists it.  (At most 5 workers fit a port slot.)"""
        if self.n_workers >= 5:
            raise RuntimeError("port slot exhausted: at most 5 workers per LocalCluster")
        i = self.n_workers
        self.n_workers += 1
        host = self.worker_hosts()[i]
        proc = self._spawn("faabric_worker", self.base_offset + 100 * (i + 1), slots or self.slots, f"worker{i}")
        deadline = time.time() + timeout
        while host not in {h["ip"] for h in self.client.available_hosts()}:
            if proc.poll() is not None:
                raise RuntimeError(f"new worker exited with {proc.returncode}")
            if time.time() > deadline:
                raise RuntimeError("new worker did not register")
            time.sleep(0.05)
        return host

    def worker_hosts(self) -> list[str]:
        return [f"127.0.0.1:{self.base_offset + 100 * (i + 1)}" for i in range(self.n_workers)]

    def stop(self) -> None:
        # Exact PIDs we started, workers first
        for p in reversed(self.procs):
            if p.poll() is None:
                p.send_signal(signal.SIGTERM)
        for p in reversed(self.procs):
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
                p.wait()
        self.procs.clear()
        if self._slot is not None:
            _live_slots.discard(self._slot)
            self._slot = None

    def __del__(self):
        # a cluster that was never started (or never stopped) gives its slot back
        slot = getattr(self, "_slot", None)
        if slot is not None:
            _live_slots.discard(slot)

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()
