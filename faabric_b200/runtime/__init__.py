"""Python front-end to the native host runtime.

* :class:`PlannerHttpClient` – the planner's JSON-over-HTTP control API
  (reference: src/planner/PlannerEndpointHandler.cpp).
* :class:`LocalCluster` – spawns ``planner_server`` plus N ``faabric_worker``
  processes on this box (distinct port offsets), the single-node analogue of
  the reference's docker-compose dist-test cluster.
"""

from .client import PlannerHttpClient, HttpMessageType, PlannerError
from .cluster import LocalCluster
from .benchmarks import (
    planner_fanout_bench,
    threads_forkjoin_bench,
    cpu_pingpong_bench,
    cpu_allreduce_bench,
    mpi_allreduce_bench,
    host_collectives_bench,
)

__all__ = [
    "PlannerHttpClient",
    "HttpMessageType",
    "PlannerError",
    "LocalCluster",
    "planner_fanout_bench",
    "threads_forkjoin_bench",
    "cpu_pingpong_bench",
    "host_collectives_bench",
    "cpu_allreduce_bench",
    "mpi_allreduce_bench",
]
