"""Distributed tests of the native runtime: planner + 2 worker PROCESSES on this
box (strategy: the reference's tests/dist suite, which does the same over a
docker-compose cluster).  CPU only."""

import base64
import json
import subprocess
from pathlib import Path

import pytest

from faabric_b200 import build as fb_build
from faabric_b200.runtime import LocalCluster, PlannerError

ROOT = Path(__file__).resolve().parents[1]
BIN = ROOT / "build" / "bin"


@pytest.fixture(scope="module", autouse=True)
def _built():
    fb_build.build(verbose=False)


@pytest.fixture(scope="module")
def cluster(tmp_path_factory):
    logs = tmp_path_factory.mktemp("cluster-logs")
    c = LocalCluster(n_workers=2, slots_per_worker=2, log_dir=logs)
    c.start()
    yield c
    c.stop()
    for f in sorted(logs.glob("*.log")):
        tail = f.read_text()[-1500:]
        print(f"---- {f.name} ----\n{tail}")


def _results(status):
    return sorted(status.get("messageResults", []), key=lambda m: m.get("mpiRank", 0))


def test_cpp_unit_suite():
    """The in-tree C++ test runner (util, schedulers, transport, planner,
    executor, state, redis, snapshots, MPI)."""
    r = subprocess.run([str(BIN / "faabric_tests")], capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0, tail
    assert " 0 failed" in r.stdout.splitlines()[-1], tail


def test_hosts_register_and_functions_spread(cluster):
    hosts = cluster.client.available_hosts()
    assert sorted(h["ip"] for h in hosts) == sorted(cluster.worker_hosts())
    assert all(h["slots"] == 2 for h in hosts)
    st = cluster.client.invoke("demo", "hello", count=4)
    outs = [m["output_data"] for m in st["messageResults"]]
    assert len(outs) == 4
    # 4 functions over 2x2 slots: both workers took part
    assert {o.split()[-1] for o in outs} == set(cluster.worker_hosts())
    assert all(m.get("returnValue", 0) == 0 for m in st["messageResults"])
    assert all(h.get("usedSlots", 0) == 0 for h in cluster.client.available_hosts())


def test_echo_and_errors(cluster):
    st = cluster.client.invoke("demo", "echo", input_data="ping")
    assert st["messageResults"][0]["output_data"] == "ping"
    st = cluster.client.invoke("demo", "error")
    assert st["messageResults"][0]["returnValue"] == 1
    st = cluster.client.invoke("demo", "nope")
    assert "Unknown function" in st["messageResults"][0]["output_data"]
    with pytest.raises(PlannerError) as e:
        cluster.client.invoke("demo", "echo", count=50)
    assert e.value.status == 500 and "No available hosts" in e.value.body


MPI_FUNCTIONS = [
    "helloworld",
    "allreduce",
    "allgather",
    "alltoall",
    "bcast",
    "barrier",
    "gather-scatter",
    "reduce-scan",
    "sendrecv",
    "isendrecv",
    "order",
    "status-probe",
    "cart",
    "rma",
    "subcomm",
    "send-many",
    "reduce-many",
    "alltoall-many",
    "sync-async",
    "typesize",
    "checks",
    "send",
    "alltoall-sleep",
]


@pytest.mark.parametrize("fn", MPI_FUNCTIONS)
def test_mpi_across_two_workers(cluster, fn):
    """World of 4 ranks, 2 per worker process: local queues inside a worker,
    TCP between the workers."""
    st = cluster.client.invoke("mpi", fn, mpi_world_size=4, timeout=60)
    res = _results(st)
    assert len(res) == 4, res
    assert [m.get("mpiRank", 0) for m in res] == [0, 1, 2, 3]
    assert all(m.get("returnValue", 0) == 0 for m in res), res
    assert {m["executedHost"] for m in res} == set(cluster.worker_hosts())
    if fn == "subcomm":
        # MPI_COMM_TYPE_SHARED groups the two ranks of each worker process
        assert all(m["output_data"] == "node of 2" for m in res), res


def test_mpi_across_three_workers(tmp_path):
    """Six ranks over three worker processes: every rank has peers in its own
    process and in two others (pairwise fence exchanges, sub-communicators
    that straddle processes, two-level collectives with three leaders)."""
    with LocalCluster(n_workers=3, slots_per_worker=2, log_dir=tmp_path) as c:
        for fn in ("rma", "subcomm", "reduce-scan", "alltoall"):
            st = c.client.invoke("mpi", fn, mpi_world_size=6, timeout=60)
            res = _results(st)
            assert len(res) == 6, (fn, res)
            assert all(m.get("returnValue", 0) == 0 for m in res), (fn, res)
            assert len({m["executedHost"] for m in res}) == 3, (fn, res)


def test_servers_survive_garbage_on_every_port(tmp_path):
    """Random bytes, truncated frames and absurd length fields on every
    listening port of the planner and the workers (RPC, HTTP): connections are
    dropped, the processes keep serving."""
    import random
    import socket

    import psutil

    with LocalCluster(n_workers=2, slots_per_worker=2, log_dir=tmp_path) as c:
        # only THIS cluster's processes (the module-wide fixture keeps serving
        # other tests and must not be poked)
        children = [psutil.Process(p.pid) for p in c.procs]
        ports = set()
        for ch in children:
            try:
                for con in ch.net_connections(kind="tcp"):
                    if con.status == "LISTEN":
                        ports.add(con.laddr.port)
            except psutil.Error:
                pass
        assert len(ports) >= 6, ports
        rnd = random.Random(7)
        for p in sorted(ports):
            for i in range(12):
                try:
                    with socket.create_connection(("127.0.0.1", p), timeout=1) as s:
                        n = rnd.choice([1, 7, 16, 17, 64, 1000, 70000])
                        s.sendall(bytes(rnd.getrandbits(8) for _ in range(n)))
                        if i % 3 == 0:
                            s.sendall(b"\x05\x00\x00\x00" + (2**40).to_bytes(8, "little") + b"\x00" * 4)
                except OSError:
                    pass
        assert all(ch.is_running() and ch.status() != psutil.STATUS_ZOMBIE for ch in children)
        assert len(c.client.available_hosts()) == 2
        st = c.client.invoke("mpi", "allreduce", mpi_world_size=4, timeout=60)
        assert [m.get("returnValue", 0) for m in _results(st)] == [0, 0, 0, 0]


def test_killed_worker_expires_and_the_rest_keep_serving(tmp_path):
    """Failure detection is membership by keep-alive (reference
    src/planner/Planner.cpp:267-291): a worker that dies silently drops out of
    the host set after the timeout and new batches avoid it."""
    import signal
    import time

    with LocalCluster(
        n_workers=2, slots_per_worker=2, log_dir=tmp_path, extra_env={"PLANNER_HOST_KEEPALIVE_TIMEOUT": "1"}
    ) as c:
        assert len(c.client.available_hosts()) == 2
        survivor = c.worker_hosts()[0]
        victim = c.procs[-1]
        victim.send_signal(signal.SIGKILL)
        victim.wait()
        deadline = time.time() + 10
        while time.time() < deadline and len(c.client.available_hosts()) != 1:
            time.sleep(0.1)
        hosts = c.client.available_hosts()
        assert [h["ip"] for h in hosts] == [survivor], hosts
        st = c.client.invoke("demo", "hello", count=2, timeout=30)
        assert [m["output_data"] for m in st["messageResults"]] == [f"hello from {survivor}"] * 2
        # a batch larger than what is left is refused, not queued
        with pytest.raises(PlannerError):
            c.client.invoke("demo", "hello", count=3, timeout=10)


def test_worker_joining_later_takes_work(tmp_path):
    """Elastic membership: a batch that does not fit is refused; once another
    worker has registered the same batch spans both, MPI world included."""
    with LocalCluster(n_workers=1, slots_per_worker=2, log_dir=tmp_path) as c:
        first = c.worker_hosts()[0]
        with pytest.raises(PlannerError):
            c.client.invoke("demo", "hello", count=4, timeout=10)
        second = c.add_worker()
        assert {h["ip"] for h in c.client.available_hosts()} == {first, second}
        st = c.client.invoke("demo", "hello", count=4, timeout=30)
        outs = sorted(m["output_data"] for m in st["messageResults"])
        assert outs == sorted([f"hello from {first}"] * 2 + [f"hello from {second}"] * 2)
        st = c.client.invoke("mpi", "allreduce", mpi_world_size=4, timeout=60)
        res = _results(st)
        assert all(m.get("returnValue", 0) == 0 for m in res), res
        assert {m["executedHost"] for m in res} == {first, second}


def test_ordered_ptp_streams_between_workers(tmp_path):
    """Every pair of four functions (two per worker process) exchanges 500
    sequence-numbered messages; receivers re-order what the network delivers
    out of order."""
    with LocalCluster(n_workers=2, slots_per_worker=2, log_dir=tmp_path) as c:
        st = c.client.invoke("ptp", "stream", count=4, input_data="500", timeout=90)
        res = st["messageResults"]
        assert len(res) == 4 and all(m.get("returnValue", 0) == 0 for m in res), res
        assert all(m["output_data"].startswith("0 out of order") for m in res), res
        assert {m["output_data"].split()[-1] for m in res} == set(c.worker_hosts())


def test_two_mpi_worlds_run_concurrently(tmp_path):
    """Two applications share the workers at the same time (reference dist
    test "multiple MPI worlds"): each gets its own world, group and ports."""
    with LocalCluster(n_workers=2, slots_per_worker=4, log_dir=tmp_path) as c:
        first = c.client.make_batch("mpi", "alltoall-many", mpi_world_size=4)
        second = c.client.make_batch("mpi", "reduce-many", mpi_world_size=4)
        c.client.execute_batch(first)
        c.client.execute_batch(second)
        for batch in (first, second):
            st = c.client.wait_for_batch(batch["appId"], timeout=90)
            res = _results(st)
            assert len(res) == 4, res
            assert all(m.get("returnValue", 0) == 0 for m in res), res
            assert len({m["mpiWorldId"] for m in res}) == 1
        assert all(h.get("usedSlots", 0) == 0 for h in c.client.available_hosts())


def test_cluster_processes_do_not_outlive_a_killed_parent(tmp_path):
    """A test runner that dies without clean-up (SIGKILL, os._exit) must not
    leave planners and workers behind holding their ports."""
    import subprocess
    import sys
    import time

    import psutil

    script = tmp_path / "abrupt.py"
    script.write_text(
        "import os, sys, pathlib\n"
        f"sys.path.insert(0, {str(ROOT)!r})\n"
        "from faabric_b200.runtime import LocalCluster\n"
        # (its own port block: the child cannot see which slots this process uses)
        f"c = LocalCluster(n_workers=2, slots_per_worker=2, base_offset=23000, log_dir=pathlib.Path({str(tmp_path)!r}))\n"
        "c.start()\n"
        "print(' '.join(str(p.pid) for p in c.procs), flush=True)\n"
        "os._exit(1)\n"
    )
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    pids = [int(x) for x in out.stdout.split()[-3:]]
    assert len(pids) == 3, (out.stdout, out.stderr)
    deadline = time.time() + 10
    alive = pids
    while alive and time.time() < deadline:
        alive = [p for p in alive if psutil.pid_exists(p) and psutil.Process(p).status() != psutil.STATUS_ZOMBIE]
        time.sleep(0.1)
    assert not alive, alive


def test_mpi_benchmarks_report(cluster):
    st = cluster.client.invoke("mpi", "bench-pingpong", mpi_world_size=2, input_data="64", timeout=120)
    out = json.loads(_results(st)[0]["output_data"])
    assert out["bytes"] == 64 and out["rtt_us"] > 0
    st = cluster.client.invoke("mpi", "bench-allreduce", mpi_world_size=4, input_data="65536,5", timeout=120)
    out = json.loads(_results(st)[0]["output_data"])
    assert out["count"] == 65536 and out["algbw_GBps"] > 0


def test_mpi_migration_between_worker_processes(tmp_path):
    """Ranks start 2 + 2 on two workers; at the migration point bin-pack
    consolidates them on one worker: two ranks snapshot their memory, push it
    to the other PROCESS and resume there."""
    with LocalCluster(n_workers=2, slots_per_worker=4, log_dir=tmp_path) as c:
        a, b = c.worker_hosts()
        batch = c.client.make_batch("mpi", "migrate", mpi_world_size=4)
        c.client.preload_decision(batch, [a, a, b, b])
        c.client.execute_batch(batch)
        st = c.client.wait_for_batch(batch["appId"], timeout=60)
        res = _results(st)
        assert len(res) == 4, res
        assert all(m.get("returnValue", 0) == 0 for m in res), res
        assert len({m["executedHost"] for m in res}) == 1, res
        assert sorted(m["output_data"] for m in res) == ["resumed at 3", "resumed at 3", "stayed", "stayed"]
        assert c.client.in_flight_apps().get("numMigrations", 0) == 1
        assert all(h.get("usedSlots", 0) == 0 for h in c.client.available_hosts())


def test_spot_eviction_moves_ranks_off_the_tainted_worker(tmp_path):
    """Policy `spot`: while the app runs, one worker is announced as the next
    to be evicted; at the migration point its ranks move to the other worker
    process (reference dist test "SPOT eviction migration")."""
    with LocalCluster(n_workers=2, slots_per_worker=4, log_dir=tmp_path) as c:
        a, b = c.worker_hosts()
        c.client.set_policy("spot")
        gate = tmp_path / "gate"
        batch = c.client.make_batch("mpi", "migrate", mpi_world_size=4)
        batch["messages"][0]["cmdline"] = str(gate)  # ranks wait for this file before the migration point
        c.client.preload_decision(batch, [a, a, b, b])
        c.client.execute_batch(batch)
        c.client.set_next_evicted_vm([b])
        assert c.client.in_flight_apps().get("nextEvictedVmIps") == [b]
        gate.write_text("go")
        st = c.client.wait_for_batch(batch["appId"], timeout=60)
        res = _results(st)
        assert len(res) == 4, res
        assert all(m.get("returnValue", 0) == 0 for m in res), res
        assert {m["executedHost"] for m in res} == {a}, res
        assert sorted(m["output_data"] for m in res) == ["resumed at 3", "resumed at 3", "stayed", "stayed"]
        assert c.client.in_flight_apps().get("numMigrations", 0) == 1
        assert all(h.get("usedSlots", 0) == 0 for h in c.client.available_hosts())


def test_group_locks_barriers_and_shared_state_across_workers(cluster):
    """Four functions of one batch, two per worker process: a counter in
    distributed state (main elected through the planner, replicas pull/push
    over the state RPCs) incremented under the point-to-point group lock."""
    st = cluster.client.invoke("ptp", "counter", count=4, input_data="5", timeout=60)
    res = st["messageResults"]
    assert len(res) == 4 and all(m.get("returnValue", 0) == 0 for m in res), res
    outs = [m["output_data"] for m in res]
    assert all(o.startswith("20 on ") for o in outs), outs
    assert {o.split()[-1] for o in outs} == set(cluster.worker_hosts())


def test_threads_fork_join_across_workers(tmp_path):
    """THREADS batch spanning two worker processes: remote threads restore the
    main thread's snapshot, their dirty pages come back as diffs and are merged
    (bytewise slots + an int Sum region)."""
    with LocalCluster(n_workers=2, slots_per_worker=3, log_dir=tmp_path) as c:
        st = c.client.invoke("demo", "threads", input_data="4", timeout=60)
        res = st["messageResults"]
        main = [m for m in res if m.get("output_data", "").startswith("merged")]
        assert len(main) == 1, res
        assert main[0]["output_data"] == "merged sum 114" and main[0].get("returnValue", 0) == 0
        thread_hosts = {m["output_data"].split()[-1] for m in res if m.get("output_data", "").startswith("thread")}
        assert thread_hosts == set(c.worker_hosts()), res


def test_repeated_reduction_across_workers(tmp_path):
    """The reference's "repeated reduction" dist test: 20 rounds of a 4-thread
    fork-join spanning two worker processes, two Sum-merged counters (one on the
    page the threads also write an array to) checked after every round."""
    with LocalCluster(n_workers=2, slots_per_worker=3, log_dir=tmp_path) as c:
        st = c.client.invoke("demo", "reduction", input_data="20", timeout=120)
        res = st["messageResults"]
        main = [m for m in res if m.get("output_data", "").startswith(("reduced", "round"))]
        assert len(main) == 1, res
        assert main[0]["output_data"] == "reduced 20 rounds to 800 / 1600", main[0]
        assert main[0].get("returnValue", 0) == 0
        thread_hosts = {m["output_data"].split()[-1] for m in res if m.get("output_data", "").startswith("thread")}
        assert thread_hosts == set(c.worker_hosts()), res


def test_exec_graph_and_policy(cluster):
    batch = cluster.client.make_batch("demo", "echo", input_data="g", record_exec_graph=True)
    cluster.client.execute_batch(batch)
    cluster.client.wait_for_batch(batch["appId"])
    graph = cluster.client.exec_graph(batch["appId"], batch["messages"][0]["id"])
    assert graph["root"]["msg"]["id"] == batch["messages"][0]["id"]
    assert cluster.client.get_policy() == "bin-pack"
    cluster.client.set_policy("compact")
    assert cluster.client.get_policy() == "compact"
    # spot policy: the next evicted VMs are visible in the in-flight report
    with pytest.raises(PlannerError):
        cluster.client.set_next_evicted_vm(cluster.worker_hosts()[:1])  # only valid under "spot"
    cluster.client.set_policy("spot")
    cluster.client.set_next_evicted_vm(cluster.worker_hosts()[:1])
    assert cluster.client.in_flight_apps().get("nextEvictedVmIps") == cluster.worker_hosts()[:1]
    cluster.client.set_next_evicted_vm([])
    cluster.client.set_policy("bin-pack")
    assert cluster.client.in_flight_apps().get("apps", []) == []


def test_is_app_migratable_tool(tmp_path):
    csv = tmp_path / "occ.csv"
    csv.write_text("WorkerIp,Slots\nfoo,7,7,-1,-1\nbar,7,7,-1,-1\n")
    assert subprocess.run([str(BIN / "is_app_migratable"), "bin-pack", "7", str(csv)], capture_output=True).returncode == 0
    csv.write_text("WorkerIp,Slots\nfoo,7,7,7,7\nbar,-1,-1,-1,-1\n")
    assert subprocess.run([str(BIN / "is_app_migratable"), "compact", "7", str(csv)], capture_output=True).returncode == 1
    assert subprocess.run([str(BIN / "example_check")], capture_output=True).returncode == 0
