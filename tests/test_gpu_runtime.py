"""GPU tests of the native host runtime's device paths (C++ suite, tag [gpu]):
MPI collectives on device buffers through the planner/executor stack, device
snapshots vs the host implementation, device-resident state."""

import subprocess
from pathlib import Path

import pytest

from faabric_b200 import build as fb_build

ROOT = Path(__file__).resolve().parents[1]
BIN = ROOT / "build" / "bin"

pytestmark = pytest.mark.gpu


def test_cpp_gpu_suite():
    fb_build.build(verbose=False)
    try:
        # (shorter than the per-test limit of the session, so that a wedged run
        # is killed here - with its output - and does not outlive pytest)
        r = subprocess.run([str(BIN / "faabric_tests"), "--tag", "gpu"], capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        raise AssertionError("C++ GPU suite timed out; last output:\n" + "\n".join(out.splitlines()[-40:]))
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-60:])
    assert r.returncode == 0, tail
    last = [l for l in r.stdout.splitlines() if l.startswith("====")][-1]
    assert " 0 failed" in last and " 0 skipped" in last, tail


def test_planner_fanout_and_cpu_baselines_run():
    """The CPU-side benches must also work on the GPU box (more cores there)."""
    from faabric_b200.runtime import planner_fanout_bench

    res = planner_fanout_bench(n_functions=1024, n_hosts=8, iters=5, warmup=2)
    assert res["functions"] == 1024 and res["e2e_us_median"] > 0
