"""A planner and two worker processes on this machine (no GPU needed): run a
function, an MPI world that spans both workers, and look at what the planner
knows.  `python examples/local_cluster.py`"""

import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from faabric_b200.runtime import LocalCluster  # noqa: E402


def main() -> int:
    # optional: first port block to use (python examples/local_cluster.py 22000)
    base_offset = int(sys.argv[1]) if len(sys.argv) > 1 else None
    with LocalCluster(n_workers=2, slots_per_worker=2, base_offset=base_offset) as cluster:
        client = cluster.client
        print("hosts:", [(h["ip"], h["slots"]) for h in client.available_hosts()])

        status = client.invoke("demo", "echo", input_data="hello")
        print("echo ->", status["messageResults"][0]["output_data"])

        # four ranks, two per worker: queues inside a worker, TCP between them
        status = client.invoke("mpi", "allreduce", mpi_world_size=4, record_exec_graph=True)
        ranks = sorted(status["messageResults"], key=lambda m: m.get("mpiRank", 0))
        for m in ranks:
            print(f"rank {m.get('mpiRank', 0)} ran on {m['executedHost']} -> {m.get('returnValue', 0)}")
        graph = client.exec_graph(ranks[0]["appId"], ranks[0]["id"])
        print("exec graph nodes:", 1 + len(graph["root"].get("chained", [])))

        print("policy:", client.get_policy(), "| in flight:", json.dumps(client.in_flight_apps()))
        ok = all(m.get("returnValue", 0) == 0 for m in ranks)
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
