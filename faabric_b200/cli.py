"""Developer CLI (the reference ships `invoke` tasks: tasks/dev.py, tasks/tests.py).

  python -m faabric_b200.cli build [--force]
  python -m faabric_b200.cli test [--gpu] [--cpp FILTER]
  python -m faabric_b200.cli bench [bench.py args…]
  python -m faabric_b200.cli cluster [--workers N] [--slots S]   # planner + workers until Ctrl-C
  python -m faabric_b200.cli invoke USER FUNCTION [--count N] [--mpi SIZE] [--input DATA] --port P
  python -m faabric_b200.cli sanitise {address,thread,undefined}   # rebuild host code with a sanitizer, run the C++ suite
  python -m faabric_b200.cli coverage                              # gcov line coverage of the host code
  python -m faabric_b200.cli sass KERNEL_REGEX                     # dump SASS of matching kernels
  python -m faabric_b200.cli format [--check]                      # clang-format over csrc/ (when installed)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    return subprocess.call(cmd, **kw)


def _coverage() -> int:
    """Line coverage of the host code by the C++ suite and the multi-process
    tests (gcov; the reference's `inv dev.coverage-report`)."""
    import glob
    import gzip  # noqa: F401  (gcov -t prints plain JSON)
    import subprocess

    env = dict(os.environ)
    env["FAABRIC_B200_COVERAGE"] = "1"
    if os.path.exists("/usr/bin/g++"):
        env["CXX"] = "/usr/bin/g++"
    obj = ROOT / "build" / "obj"
    for f in glob.glob(str(obj / "*.gcda")):
        os.unlink(f)
    rc = _run([sys.executable, "-m", "faabric_b200.build", "--force"], cwd=ROOT, env=env)
    if rc != 0:
        return rc
    _run([str(ROOT / "build" / "bin" / "faabric_tests")], env=env)
    _run([sys.executable, "-m", "pytest", "tests", "-q", "-m", "not gpu"], cwd=ROOT, env=env)
    per_file: dict[str, dict[int, int]] = {}
    for gcda in sorted(glob.glob(str(obj / "*.gcda"))):
        r = subprocess.run(["gcov", "-j", "-t", gcda], capture_output=True, text=True, cwd=obj)
        if r.returncode != 0 or not r.stdout.strip():
            continue
        for doc in r.stdout.splitlines():
            try:
                data = json.loads(doc)
            except ValueError:
                continue
            for f in data.get("files", []):
                name = os.path.normpath(os.path.join(str(obj), f["file"]))
                if "/csrc/src/" not in name and "/csrc/capi/" not in name:
                    continue
                lines = per_file.setdefault(name, {})
                for ln in f["lines"]:
                    lines[ln["line_number"]] = lines.get(ln["line_number"], 0) + ln["count"]
    by_dir: dict[str, list[int]] = {}
    for name, lines in per_file.items():
        d = os.path.relpath(os.path.dirname(name), ROOT / "csrc")
        tot = by_dir.setdefault(d, [0, 0])
        tot[0] += sum(1 for c in lines.values() if c > 0)
        tot[1] += len(lines)
    print(f"{'directory':<28}{'lines':>8}{'covered':>9}{'%':>7}")
    all_cov = all_tot = 0
    for d in sorted(by_dir):
        cov, tot = by_dir[d]
        all_cov += cov
        all_tot += tot
        print(f"{d:<28}{tot:>8}{cov:>9}{100.0 * cov / max(tot, 1):>7.1f}")
    print(f"{'TOTAL':<28}{all_tot:>8}{all_cov:>9}{100.0 * all_cov / max(all_tot, 1):>7.1f}")
    files = {
        os.path.relpath(name, ROOT / "csrc"): [sum(1 for c in lines.values() if c > 0), len(lines)]
        for name, lines in per_file.items()
    }
    (ROOT / "build" / "coverage.json").write_text(
        json.dumps({"dirs": by_dir, "files": files, "covered": all_cov, "lines": all_tot}, indent=1)
    )
    print("(rebuild without instrumentation: python -m faabric_b200.build --force)")
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="faabric_b200.cli")
    sub = ap.add_subparsers(dest="cmd", required=True)
    b = sub.add_parser("build")
    b.add_argument("--force", action="store_true")
    t = sub.add_parser("test")
    t.add_argument("--gpu", action="store_true")
    t.add_argument("--cpp", default=None, help="run only the C++ suite with this name filter")
    be = sub.add_parser("bench")
    be.add_argument("rest", nargs=argparse.REMAINDER)
    c = sub.add_parser("cluster")
    c.add_argument("--workers", type=int, default=2)
    c.add_argument("--slots", type=int, default=4)
    i = sub.add_parser("invoke")
    i.add_argument("user")
    i.add_argument("function")
    i.add_argument("--count", type=int, default=1)
    i.add_argument("--mpi", type=int, default=0)
    i.add_argument("--input", default=None)
    i.add_argument("--host", default="127.0.0.1")
    i.add_argument("--port", type=int, default=8080)
    s = sub.add_parser("sanitise")
    s.add_argument("kind", choices=["address", "thread", "undefined"])
    sub.add_parser("coverage")
    fm = sub.add_parser("format")
    fm.add_argument("--check", action="store_true")
    sa = sub.add_parser("sass")
    sa.add_argument("regex")
    a = ap.parse_args(argv)

    if a.cmd == "build":
        from . import build as _b

        _b.build(force=a.force)
        return 0
    if a.cmd == "test":
        if a.cpp is not None:
            return _run([str(ROOT / "build" / "bin" / "faabric_tests"), a.cpp])
        marker = "gpu" if a.gpu else "not gpu"
        return _run([sys.executable, "-m", "pytest", "tests", "-x", "-q", "-m", marker], cwd=ROOT)
    if a.cmd == "bench":
        return _run([sys.executable, str(ROOT / "bench.py")] + a.rest, cwd=ROOT)
    if a.cmd == "cluster":
        from .runtime import LocalCluster

        with LocalCluster(n_workers=a.workers, slots_per_worker=a.slots, log_dir=ROOT / "build" / "cluster-logs") as cl:
            print(f"planner HTTP on 127.0.0.1:{cl.http_port}; workers: {cl.worker_hosts()} (Ctrl-C to stop)", flush=True)
            try:
                while True:
                    time.sleep(1)
            except KeyboardInterrupt:
                pass
        return 0
    if a.cmd == "invoke":
        from .runtime import PlannerHttpClient

        cli = PlannerHttpClient(a.host, a.port)
        st = cli.invoke(a.user, a.function, a.count, input_data=a.input, mpi_world_size=a.mpi)
        print(json.dumps(st, indent=1))
        return 0
    if a.cmd == "sanitise":
        # Host code only (device code has compute-sanitizer): separate object dir
        env = dict(os.environ)
        env["FAABRIC_B200_SANITISE"] = a.kind
        # the sanitizer runtimes ship with the system compiler
        if os.path.exists("/usr/bin/g++"):
            env["CXX"] = "/usr/bin/g++"
        rc = _run([sys.executable, "-m", "faabric_b200.build", "--force"], cwd=ROOT, env=env)
        if rc != 0:
            return rc
        return _run([str(ROOT / "build" / "bin" / "faabric_tests")], env=env)
    if a.cmd == "coverage":
        return _coverage()
    if a.cmd == "format":
        # the reference's `inv format-code`; style file at the repo root
        import shutil

        exe = shutil.which("clang-format")
        if exe is None:
            print("clang-format is not installed", file=sys.stderr)
            return 2
        files = [str(p) for pat in ("**/*.cpp", "**/*.h", "**/*.cu", "**/*.cuh") for p in (ROOT / "csrc").glob(pat)]
        files = [f for f in files if not f.endswith(".pb.h")]
        return _run([exe, "--dry-run", "--Werror"] + files if a.check else [exe, "-i"] + files)
    if a.cmd == "sass":
        import re
        import subprocess

        lib = ROOT / "faabric_b200" / "lib" / "libfaabric_b200.so"
        listing = subprocess.run(["cuobjdump", "-res-usage", str(lib)], capture_output=True, text=True).stdout
        names = [n for n in re.findall(r"Function (\S+):", listing) if re.search(a.regex, n)]
        if not names:
            print(f"no kernel matches {a.regex!r}", file=sys.stderr)
            return 1
        if len(names) > 3:
            print(f"# {len(names)} kernels match, showing the first 3", file=sys.stderr)
        rc = 0
        for n in names[:3]:
            rc |= _run(["cuobjdump", "-sass", "-fun", n, str(lib)])
        return rc
    return 1


if __name__ == "__main__":
    raise SystemExit(main())
