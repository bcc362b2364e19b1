"""HTTP client for the planner control API."""

from __future__ import annotations

import enum
import itertools
import json
import os
import random
import time
import urllib.error
import urllib.request


class HttpMessageType(enum.IntEnum):
    NO_TYPE = 0
    RESET = 1
    FLUSH_AVAILABLE_HOSTS = 2
    FLUSH_EXECUTORS = 3
    FLUSH_SCHEDULING_STATE = 4
    GET_AVAILABLE_HOSTS = 5
    GET_CONFIG = 6
    GET_EXEC_GRAPH = 7
    GET_IN_FLIGHT_APPS = 8
    EXECUTE_BATCH = 10
    EXECUTE_BATCH_STATUS = 11
    PRELOAD_SCHEDULING_DECISION = 12
    SET_POLICY = 13
    GET_POLICY = 14
    SET_NEXT_EVICTED_VM = 15


class PlannerError(RuntimeError):
    def __init__(self, status: int, body: str):
        super().__init__(f"planner returned {status}: {body}")
        self.status = status
        self.body = body


_gid = itertools.count(random.randint(1, 1 << 20) * 1000 + (os.getpid() % 1000) * 1_000_000)


def _next_id() -> int:
    return next(_gid) & 0x7FFFFFFF


class PlannerHttpClient:
    def __init__(self, host: str = "127.0.0.1", port: int = 8080, timeout: float = 30.0):
        self.url = f"http://{host}:{port}/"
        self.timeout = timeout

    # ---- raw ----
    def post(self, msg_type: HttpMessageType, payload: str | None = None) -> tuple[int, str]:
        body = {"http_type": int(msg_type)}
        if payload is not None:
            body["payload"] = payload
        req = urllib.request.Request(self.url, data=json.dumps(body).encode(), method="POST")
        try:
            with urllib.request.urlopen(req, timeout=self.timeout) as r:
                return r.status, r.read().decode()
        except urllib.error.HTTPError as e:
            return e.code, e.read().decode()

    def _ok(self, msg_type, payload=None) -> str:
        status, body = self.post(msg_type, payload)
        if status != 200:
            raise PlannerError(status, body)
        return body

    # ---- control ----
    def reset(self) -> str:
        return self._ok(HttpMessageType.RESET)

    def flush_hosts(self) -> str:
        return self._ok(HttpMessageType.FLUSH_AVAILABLE_HOSTS)

    def flush_executors(self) -> str:
        return self._ok(HttpMessageType.FLUSH_EXECUTORS)

    def flush_scheduling_state(self) -> str:
        return self._ok(HttpMessageType.FLUSH_SCHEDULING_STATE)

    def available_hosts(self) -> list[dict]:
        return json.loads(self._ok(HttpMessageType.GET_AVAILABLE_HOSTS)).get("hosts", [])

    def config(self) -> dict:
        return json.loads(self._ok(HttpMessageType.GET_CONFIG))

    def in_flight_apps(self) -> dict:
        return json.loads(self._ok(HttpMessageType.GET_IN_FLIGHT_APPS))

    def get_policy(self) -> str:
        return self._ok(HttpMessageType.GET_POLICY)

    def set_policy(self, policy: str) -> str:
        return self._ok(HttpMessageType.SET_POLICY, policy)

    def set_next_evicted_vm(self, ips: list[str]) -> str:
        return self._ok(HttpMessageType.SET_NEXT_EVICTED_VM, json.dumps({"vmIps": ips}))

    def exec_graph(self, app_id: int, msg_id: int) -> dict:
        return json.loads(self._ok(HttpMessageType.GET_EXEC_GRAPH, json.dumps({"id": msg_id, "appId": app_id})))

    # ---- execution ----
    @staticmethod
    def make_batch(
        user: str,
        function: str,
        count: int = 1,
        *,
        input_data: str | None = None,
        mpi_world_size: int = 0,
        record_exec_graph: bool = False,
    ) -> dict:
        """A BatchExecuteRequest in its JSON form."""
        app_id = _next_id()
        msgs = []
        for i in range(count):
            m = {"id": _next_id(), "appId": app_id, "appIdx": i, "groupIdx": i, "user": user, "function": function}
            if input_data is not None:
                # bytes fields travel base64 encoded
                import base64

                m["input_data"] = base64.b64encode(input_data.encode()).decode()
            if mpi_world_size > 0:
                m["mpi"] = True
                m["mpi_world_size"] = mpi_world_size
            if record_exec_graph:
                m["record_exec_graph"] = True
            msgs.append(m)
        return {"appId": app_id, "user": user, "function": function, "messages": msgs}

    def preload_decision(self, batch: dict, hosts: list[str]) -> str:
        """Pin the placement of an app before it is scheduled: hosts[i] is the
        host of group idx i (MPI rank i)."""
        msgs = [
            {"id": 0, "appId": batch["appId"], "appIdx": i, "groupIdx": i, "executedHost": h}
            for i, h in enumerate(hosts)
        ]
        payload = {"appId": batch["appId"], "messages": msgs}
        return self._ok(HttpMessageType.PRELOAD_SCHEDULING_DECISION, json.dumps(payload))

    def execute_batch(self, batch: dict) -> dict:
        return json.loads(self._ok(HttpMessageType.EXECUTE_BATCH, json.dumps(batch)))

    def batch_status(self, app_id: int) -> dict | None:
        status, body = self.post(HttpMessageType.EXECUTE_BATCH_STATUS, json.dumps({"appId": app_id}))
        if status == 500 and "not registered" in body:
            return None
        if status != 200:
            raise PlannerError(status, body)
        return json.loads(body)

    def wait_for_batch(self, app_id: int, timeout: float = 60.0, poll: float = 0.01) -> dict:
        deadline = time.time() + timeout
        while time.time() < deadline:
            st = self.batch_status(app_id)
            if st is not None and st.get("finished"):
                return st
            time.sleep(poll)
        raise TimeoutError(f"app {app_id} did not finish within {timeout}s")

    def invoke(self, user: str, function: str, count: int = 1, timeout: float = 60.0, **kw) -> dict:
        batch = self.make_batch(user, function, count, **kw)
        self.execute_batch(batch)
        return self.wait_for_batch(batch["appId"], timeout)
