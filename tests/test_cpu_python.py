"""CPU-only checks of the Python layer: the benchmark workload definition, the
all-reduce policies, NUMA helpers and the HTTP client's request building."""

import json

import pytest

from faabric_b200.models import resnet50_grads
from faabric_b200.models.grad_sync import POLICIES
from faabric_b200.runtime.client import PlannerHttpClient
from faabric_b200.utils import numa


def test_resnet50_gradient_list_matches_the_reference_benchmark():
    # reference tests/dist/mpi/benchmarks/mpi_bench.cpp:25-56: 214 tensors,
    # 25,583,592 ints, the largest 2,359,296 (3x3x512x512), fc layer first
    sizes = resnet50_grads.resnet50_grad_sizes()
    assert len(sizes) == 214
    assert sum(sizes) == 25_583_592
    assert max(sizes) == 2_359_296
    assert sizes[0] == 1000 and sizes[1] == 2048 * 1000
    assert sizes[-1] == 7 * 7 * 3 * 64 and sizes[-4:-1] == [64, 64, 64]
    assert all(s > 0 for s in sizes)
    small = resnet50_grads.small_sizes()
    assert len(small) == 1000 and set(small) == {8}


def test_policies_pick_one_algorithm_per_size():
    valid = {"auto", "ll", "oneshot", "twoshot", "nvls"}
    for name, policy in POLICIES.items():
        for nvls in (False, True):
            picks = {n: policy(n, nvls) for n in (64, 4096, 1 << 17, 1 << 20, 1 << 24, 1 << 27)}
            assert set(picks.values()) <= valid, (name, picks)
            if not nvls:
                assert "nvls" not in picks.values(), (name, picks)
    assert POLICIES["auto"](1 << 20, True) == "auto"
    assert POLICIES["twoshot"](1 << 22, True) == "twoshot"


def test_cpulist_parsing_and_fallbacks():
    assert numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("5") == [5]
    assert numa._parse_cpulist("") == []
    # no such GPU (or no sysfs entry for it): empty list, and binding is a no-op
    assert numa.gpu_local_cpus(4096) == []
    assert numa.bind_process_near_gpu(4096) == []


class _Recorder(PlannerHttpClient):
    """Captures requests instead of sending them."""

    def __init__(self):
        super().__init__("127.0.0.1", 1)
        self.sent = []
        self.reply = (200, "{}")

    def post(self, msg_type, payload=None):
        self.sent.append((int(msg_type), payload))
        return self.reply


def test_http_client_builds_the_reference_requests():
    c = _Recorder()
    batch = c.make_batch("mpi", "allreduce", mpi_world_size=4, input_data="abc")
    assert batch["user"] == "mpi" and batch["function"] == "allreduce"
    assert len(batch["messages"]) == 1
    m = batch["messages"][0]
    assert m["appId"] == batch["appId"] and m["mpi"] is True and m["mpi_world_size"] == 4
    assert m["input_data"] == "YWJj"  # bytes fields are base64 in protobuf JSON
    several = c.make_batch("demo", "echo", count=3)
    ids = [x["id"] for x in several["messages"]]
    assert len(set(ids)) == 3 and [x["appIdx"] for x in several["messages"]] == [0, 1, 2]

    c.reply = (200, json.dumps({"hosts": [{"ip": "a", "slots": 2}]}))
    assert c.available_hosts() == [{"ip": "a", "slots": 2}]
    c.reply = (200, "ok")
    c.set_policy("compact")
    c.set_next_evicted_vm(["10.0.0.1"])
    c.preload_decision(several, ["h1", "h2", "h1"])
    types = [t for t, _ in c.sent]
    assert types == [5, 13, 15, 12]
    assert c.sent[1][1] == "compact"
    assert json.loads(c.sent[2][1]) == {"vmIps": ["10.0.0.1"]}
    preload = json.loads(c.sent[3][1])
    assert [x["executedHost"] for x in preload["messages"]] == ["h1", "h2", "h1"]

    from faabric_b200.runtime.client import PlannerError

    c.reply = (500, "No available hosts")
    with pytest.raises(PlannerError) as e:
        c.execute_batch(several)
    assert e.value.status == 500 and "No available hosts" in e.value.body


def test_local_cluster_example_runs():
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    # (a port block of its own: the child cannot see which ones this process uses)
    r = subprocess.run(
        [sys.executable, str(root / "examples" / "local_cluster.py"), "22000"], capture_output=True, text=True, timeout=180
    )
    assert r.returncode == 0, r.stdout + r.stderr
    assert "echo -> hello" in r.stdout
    assert r.stdout.count("ran on") == 4
