#!/usr/bin/env python
"""Generates include/faabric/proto/faabric.pb.h: plain C++ message classes with
the protobuf-style accessor surface the reference's generated code exposes
(foo(), set_foo(), mutable_foo(), add_foo(), foo_size(), SerializeAsString,
ParseFromString, ...) so reference-style user code compiles unchanged.

No protoc / libprotobuf in this environment, so the schema lives here as Python
data and the wire format is implemented by include/faabric/proto/wire.h (it is
protobuf wire-format compatible: varints + length-delimited fields).  JSON uses
the same json_name mapping as the reference schema.

Schema source of truth for parity: the reference's src/proto/faabric.proto:21-242
and src/planner/planner.proto:9-153.
"""
from __future__ import annotations

import sys
from pathlib import Path

# field: (name, type, number, json_name or None, label)
# type: int32 int64 uint32 uint64 bool string bytes enum:<Enum> msg:<Message>
#       map<string,int32> map<string,string>
# label: "" | "repeated"

def F(name, typ, num, json=None, rep=False):
    return dict(name=name, type=typ, num=num, json=json or name, rep=rep)

FAABRIC = dict(
    namespace="faabric",
    messages=[
        dict(name="EmptyResponse", fields=[F("empty", "int32", 1)]),
        dict(name="EmptyRequest", fields=[F("empty", "int32", 1)]),
        dict(name="HostResources", fields=[F("slots", "int32", 1), F("usedSlots", "int32", 2)]),
        dict(
            name="FunctionStatusResponse",
            enums=[("FunctionStatus", [("OK", 0), ("ERROR", 1)])],
            fields=[F("status", "enum:FunctionStatus", 1)],
        ),
        dict(
            name="Message",
            enums=[("MessageType", [("CALL", 0), ("KILL", 1), ("EMPTY", 2), ("FLUSH", 3)])],
            fields=[
                F("id", "int32", 1), F("appId", "int32", 2), F("appIdx", "int32", 3),
                F("mainHost", "string", 4), F("type", "enum:MessageType", 5),
                F("user", "string", 6), F("function", "string", 7),
                F("inputData", "bytes", 8, "input_data"), F("outputData", "string", 9, "output_data"),
                F("funcPtr", "int32", 10), F("returnValue", "int32", 11), F("snapshotKey", "string", 12),
                F("startTimestamp", "int64", 14, "start_ts"), F("resultKey", "string", 15),
                F("executesLocally", "bool", 16), F("statusKey", "string", 17),
                F("executedHost", "string", 18), F("finishTimestamp", "int64", 19, "finish_ts"),
                F("isPython", "bool", 21, "python"), F("pythonUser", "string", 24, "py_user"),
                F("pythonFunction", "string", 25, "py_func"), F("pythonEntry", "string", 26),
                F("groupId", "int32", 27), F("groupIdx", "int32", 28), F("groupSize", "int32", 29),
                F("isMpi", "bool", 30, "mpi"), F("mpiWorldId", "int32", 31), F("mpiRank", "int32", 32),
                F("mpiWorldSize", "int32", 33, "mpi_world_size"), F("cmdline", "string", 34),
                F("recordExecGraph", "bool", 35, "record_exec_graph"),
                F("chainedMsgIds", "int32", 36, rep=True),
                F("intExecGraphDetails", "map<string,int32>", 37),
                F("execGraphDetails", "map<string,string>", 38),
                F("isOmp", "bool", 39), F("ompNumThreads", "int32", 40),
            ],
        ),
        dict(
            name="BatchExecuteRequest",
            enums=[("BatchExecuteType", [("FUNCTIONS", 0), ("THREADS", 1), ("PROCESSES", 2), ("MIGRATION", 3)])],
            fields=[
                F("appId", "int32", 1), F("groupId", "int32", 2), F("user", "string", 3),
                F("function", "string", 4), F("type", "enum:BatchExecuteType", 5),
                F("snapshotKey", "string", 6), F("messages", "msg:Message", 7, rep=True),
                F("subType", "int32", 8), F("contextData", "bytes", 9), F("singleHost", "bool", 10),
                F("singleHostHint", "bool", 11), F("elasticScaleHint", "bool", 12),
            ],
        ),
        dict(
            name="BatchExecuteRequestStatus",
            fields=[
                F("appId", "int32", 1), F("finished", "bool", 2),
                F("messageResults", "msg:Message", 3, rep=True), F("expectedNumMessages", "int32", 4),
            ],
        ),
        dict(name="StateRequest", fields=[F("user", "string", 1), F("key", "string", 2), F("data", "bytes", 3)]),
        dict(name="StateChunkRequest", fields=[F("user", "string", 1), F("key", "string", 2), F("offset", "uint64", 3), F("chunkSize", "uint64", 4)]),
        dict(name="StateResponse", fields=[F("user", "string", 1), F("key", "string", 2), F("data", "bytes", 3)]),
        dict(name="StatePart", fields=[F("user", "string", 1), F("key", "string", 2), F("offset", "uint64", 3), F("data", "bytes", 4)]),
        dict(name="StateSizeResponse", fields=[F("user", "string", 1), F("key", "string", 2), F("stateSize", "uint64", 3)]),
        dict(name="StateAppendedRequest", fields=[F("user", "string", 1), F("key", "string", 2), F("nValues", "uint32", 3)]),
        dict(
            name="StateAppendedResponse",
            nested=[dict(name="AppendedValue", fields=[F("data", "bytes", 2)])],
            fields=[F("user", "string", 1), F("key", "string", 2), F("values", "msg:AppendedValue", 3, rep=True)],
        ),
        dict(
            name="PointToPointMessage",
            fields=[F("appId", "int32", 1), F("groupId", "int32", 2), F("sendIdx", "int32", 3), F("recvIdx", "int32", 4), F("data", "bytes", 5)],
        ),
        dict(
            name="PointToPointMappings",
            nested=[dict(name="PointToPointMapping", fields=[F("host", "string", 1), F("messageId", "int32", 2), F("appIdx", "int32", 3), F("groupIdx", "int32", 4), F("mpiPort", "int32", 5)])],
            fields=[F("appId", "int32", 1), F("groupId", "int32", 2), F("mappings", "msg:PointToPointMapping", 3, rep=True)],
        ),
        dict(
            name="PendingMigration",
            fields=[F("appId", "int32", 1), F("groupId", "int32", 2), F("groupIdx", "int32", 3), F("srcHost", "string", 4), F("dstHost", "string", 5)],
        ),
        # --- snapshot RPC payloads (flatbuffers tables in the reference:
        # src/flat/faabric.fbs:1-38; offsets widened to 64 bit here) ---
        dict(name="SnapshotMergeRegionRequest", fields=[F("offset", "uint64", 1), F("length", "uint64", 2), F("dataType", "int32", 3, "data_type"), F("mergeOp", "int32", 4, "merge_op")]),
        dict(name="SnapshotDiffRequest", fields=[F("offset", "uint64", 1), F("dataType", "int32", 2, "data_type"), F("mergeOp", "int32", 3, "merge_op"), F("data", "bytes", 4)]),
        dict(name="SnapshotPushRequest", fields=[F("key", "string", 1), F("maxSize", "uint64", 2, "max_size"), F("contents", "bytes", 3), F("mergeRegions", "msg:SnapshotMergeRegionRequest", 4, "merge_regions", rep=True), F("deviceResident", "bool", 5), F("devicePtr", "uint64", 6), F("deviceId", "int32", 7), F("ownerPid", "int32", 8, "owner_pid"), F("ipcHandle", "bytes", 9, "ipc_handle"), F("deviceSize", "uint64", 10, "device_size")]),
        dict(name="SnapshotDeleteRequest", fields=[F("key", "string", 1)]),
        dict(name="SnapshotUpdateRequest", fields=[F("key", "string", 1), F("mergeRegions", "msg:SnapshotMergeRegionRequest", 2, "merge_regions", rep=True), F("diffs", "msg:SnapshotDiffRequest", 3, rep=True)]),
        dict(name="ThreadResultRequest", fields=[F("appId", "int32", 1, "app_id"), F("messageId", "int32", 2, "message_id"), F("returnValue", "int32", 3, "return_value"), F("key", "string", 4), F("diffs", "msg:SnapshotDiffRequest", 5, rep=True), F("deviceMerged", "bool", 6, "device_merged"), F("deviceDiffBytes", "uint64", 7, "device_diff_bytes")]),
    ],
)

PLANNER = dict(
    namespace="faabric::planner",
    messages=[
        dict(name="EmptyResponse", fields=[F("empty", "int32", 1)]),
        dict(name="EmptyRequest", fields=[F("empty", "int32", 1)]),
        dict(name="ResponseStatus", enums=[("Status", [("OK", 0), ("ERROR", 1)])], fields=[F("status", "enum:Status", 1)]),
        dict(name="Timestamp", fields=[F("epochMs", "int64", 1)]),
        dict(
            name="HttpMessage",
            enums=[("Type", [("NO_TYPE", 0), ("RESET", 1), ("FLUSH_AVAILABLE_HOSTS", 2), ("FLUSH_EXECUTORS", 3),
                             ("FLUSH_SCHEDULING_STATE", 4), ("GET_AVAILABLE_HOSTS", 5), ("GET_CONFIG", 6),
                             ("GET_EXEC_GRAPH", 7), ("GET_IN_FLIGHT_APPS", 8), ("EXECUTE_BATCH", 10),
                             ("EXECUTE_BATCH_STATUS", 11), ("PRELOAD_SCHEDULING_DECISION", 12), ("SET_POLICY", 13),
                             ("GET_POLICY", 14), ("SET_NEXT_EVICTED_VM", 15)])],
            fields=[F("type", "enum:Type", 1, "http_type"), F("payloadJson", "string", 2, "payload")],
        ),
        dict(
            name="GetInFlightAppsResponse",
            nested=[
                dict(name="InFlightApp", fields=[F("appId", "int32", 1), F("subType", "int32", 2), F("size", "int32", 3), F("hostIps", "string", 4, rep=True)]),
                dict(name="FrozenApp", fields=[F("appId", "int32", 1), F("subType", "int32", 2), F("size", "int32", 3)]),
            ],
            fields=[F("apps", "msg:InFlightApp", 1, rep=True), F("numMigrations", "int32", 2), F("nextEvictedVmIps", "string", 3, rep=True), F("frozenApps", "msg:FrozenApp", 4, rep=True)],
        ),
        dict(name="NumMigrationsResponse", fields=[F("numMigrations", "int32", 1)]),
        dict(name="PlannerConfig", fields=[F("ip", "string", 1), F("hostTimeout", "int32", 2), F("numThreadsHttpServer", "int32", 3)]),
        dict(
            name="Host",
            nested=[dict(name="MpiPort", fields=[F("port", "int32", 1), F("used", "bool", 2)])],
            fields=[F("ip", "string", 1), F("slots", "int32", 2), F("usedSlots", "int32", 3), F("registerTs", "msg:Timestamp", 4), F("mpiPorts", "msg:MpiPort", 5, rep=True)],
        ),
        dict(name="PingResponse", fields=[F("config", "msg:PlannerConfig", 1)]),
        dict(name="RegisterHostRequest", fields=[F("host", "msg:Host", 1), F("overwrite", "bool", 2)]),
        dict(name="RegisterHostResponse", fields=[F("status", "msg:ResponseStatus", 1), F("config", "msg:PlannerConfig", 2), F("hostId", "int32", 3)]),
        dict(name="RemoveHostRequest", fields=[F("host", "msg:Host", 1)]),
        dict(name="RemoveHostResponse", fields=[F("status", "msg:ResponseStatus", 1)]),
        dict(name="AvailableHostsResponse", fields=[F("hosts", "msg:Host", 1, rep=True)]),
        dict(name="SetEvictedVmIpsRequest", fields=[F("vmIps", "string", 1, rep=True)]),
        # Election of the "main" host of an in-memory state value.  (The reference
        # does this through a Redis key + Redis lock; here the planner arbitrates.)
        dict(name="StateMainRequest", fields=[F("user", "string", 1), F("key", "string", 2), F("host", "string", 3), F("claim", "bool", 4), F("drop", "bool", 5)]),
        dict(name="StateMainResponse", fields=[F("host", "string", 1)]),
    ],
)

CPP_TYPE = {"int32": "int32_t", "int64": "int64_t", "uint32": "uint32_t", "uint64": "uint64_t", "bool": "bool"}


def lower(s):
    return s.lower()


def gen_message(m, out, indent=""):
    name = m["name"]
    w = lambda s="": out.append(indent + s)
    w(f"class {name}")
    w("{")
    w("  public:")
    for en, vals in m.get("enums", []):
        w(f"    enum {en} : int")
        w("    {")
        for vn, vv in vals:
            w(f"        {vn} = {vv},")
        w("    };")
        w(f"    static const char* {en}_Name(int v)")
        w("    {")
        w("        switch (v) {")
        for vn, vv in vals:
            w(f"            case {vv}: return \"{vn}\";")
        w("            default: return \"\";")
        w("        }")
        w("    }")
        w(f"    static bool {en}_Parse(const std::string& s, {en}* out)")
        w("    {")
        for vn, vv in vals:
            w(f"        if (s == \"{vn}\") {{ *out = {vn}; return true; }}")
        w("        return false;")
        w("    }")
    for nm in m.get("nested", []):
        gen_message(nm, out, indent + "    ")
    w()
    w(f"    {name}() = default;")
    w()
    # accessors
    for f in m["fields"]:
        n, t, rep = f["name"], f["type"], f["rep"]
        ln = lower(n)
        member = f"{n}_"
        if t.startswith("map<"):
            vt = "int32_t" if "int32" in t else "std::string"
            w(f"    const std::map<std::string, {vt}>& {ln}() const {{ return {member}; }}")
            w(f"    std::map<std::string, {vt}>* mutable_{ln}() {{ return &{member}; }}")
            w(f"    int {ln}_size() const {{ return (int){member}.size(); }}")
            w(f"    void clear_{ln}() {{ {member}.clear(); }}")
        elif rep:
            if t.startswith("msg:"):
                ct = t[4:]
                w(f"    int {ln}_size() const {{ return (int){member}.size(); }}")
                w(f"    const {ct}& {ln}(int i) const {{ return {member}.at(i); }}")
                w(f"    {ct}* mutable_{ln}(int i) {{ return &{member}.at(i); }}")
                w(f"    {ct}* add_{ln}() {{ {member}.emplace_back(); return &{member}.back(); }}")
                w(f"    const faabric::proto::RepeatedField<{ct}>& {ln}() const {{ return {member}; }}")
                w(f"    faabric::proto::RepeatedField<{ct}>* mutable_{ln}() {{ return &{member}; }}")
                w(f"    void clear_{ln}() {{ {member}.clear(); }}")
            else:
                ct = "std::string" if t in ("string", "bytes") else CPP_TYPE[t]
                w(f"    int {ln}_size() const {{ return (int){member}.size(); }}")
                ret = f"const {ct}&" if ct == "std::string" else ct
                w(f"    {ret} {ln}(int i) const {{ return {member}.at(i); }}")
                w(f"    void set_{ln}(int i, {ret} v) {{ {member}.at(i) = v; }}")
                w(f"    void add_{ln}({ret} v) {{ {member}.push_back(v); }}")
                w(f"    const faabric::proto::RepeatedField<{ct}>& {ln}() const {{ return {member}; }}")
                w(f"    faabric::proto::RepeatedField<{ct}>* mutable_{ln}() {{ return &{member}; }}")
                w(f"    void clear_{ln}() {{ {member}.clear(); }}")
        elif t in ("string", "bytes"):
            w(f"    const std::string& {ln}() const {{ return {member}; }}")
            w(f"    void set_{ln}(const std::string& v) {{ {member} = v; }}")
            w(f"    void set_{ln}(std::string&& v) {{ {member} = std::move(v); }}")
            w(f"    void set_{ln}(const char* v) {{ {member} = v; }}")
            w(f"    void set_{ln}(const void* v, size_t n) {{ {member}.assign((const char*)v, n); }}")
            w(f"    std::string* mutable_{ln}() {{ return &{member}; }}")
            w(f"    void clear_{ln}() {{ {member}.clear(); }}")
        elif t.startswith("enum:"):
            et = t[5:]
            w(f"    {et} {ln}() const {{ return {member}; }}")
            w(f"    void set_{ln}({et} v) {{ {member} = v; }}")
            w(f"    void clear_{ln}() {{ {member} = ({et})0; }}")
        elif t.startswith("msg:"):
            ct = t[4:]
            w(f"    const {ct}& {ln}() const {{ return {member}; }}")
            w(f"    {ct}* mutable_{ln}() {{ has_{ln}_ = true; return &{member}; }}")
            w(f"    bool has_{ln}() const {{ return has_{ln}_; }}")
            w(f"    void clear_{ln}() {{ {member} = {ct}(); has_{ln}_ = false; }}")
        else:
            ct = CPP_TYPE[t]
            w(f"    {ct} {ln}() const {{ return {member}; }}")
            w(f"    void set_{ln}({ct} v) {{ {member} = v; }}")
            w(f"    void clear_{ln}() {{ {member} = 0; }}")
    w()
    # Clear / CopyFrom
    w("    void Clear() { *this = " + name + "(); }")
    w(f"    void CopyFrom(const {name}& other) {{ *this = other; }}")
    w(f"    void Swap({name}* other) {{ std::swap(*this, *other); }}")
    # serialise
    w("    void encode(faabric::proto::Writer& w) const")
    w("    {")
    for f in m["fields"]:
        n, t, rep, num = f["name"], f["type"], f["rep"], f["num"]
        member = f"{n}_"
        if t.startswith("map<"):
            isint = "int32" in t
            w(f"        for (const auto& kv : {member}) {{")
            w("            faabric::proto::Writer e;")
            w("            e.str(1, kv.first, true);")
            w("            " + ("e.varint(2, (uint64_t)(int64_t)kv.second, true);" if isint else "e.str(2, kv.second, true);"))
            w(f"            w.str({num}, e.data(), true);")
            w("        }")
        elif rep:
            if t.startswith("msg:"):
                w(f"        for (const auto& v : {member}) {{ faabric::proto::Writer e; v.encode(e); w.str({num}, e.data(), true); }}")
            elif t in ("string", "bytes"):
                w(f"        for (const auto& v : {member}) {{ w.str({num}, v, true); }}")
            else:
                w(f"        if (!{member}.empty()) {{ faabric::proto::Writer e; for (auto v : {member}) {{ e.rawVarint((uint64_t)(int64_t)v); }} w.str({num}, e.data(), true); }}")
        elif t in ("string", "bytes"):
            w(f"        w.str({num}, {member}, false);")
        elif t.startswith("msg:"):
            w(f"        if (has_{lower(n)}_) {{ faabric::proto::Writer e; {member}.encode(e); w.str({num}, e.data(), true); }}")
        else:
            w(f"        w.varint({num}, (uint64_t)(int64_t){member}, false);")
    w("    }")
    w("    bool decode(faabric::proto::Reader& r)")
    w("    {")
    w("        uint32_t field; int wt;")
    w("        while (r.next(field, wt)) {")
    w("            switch (field) {")
    for f in m["fields"]:
        n, t, rep, num = f["name"], f["type"], f["rep"], f["num"]
        member = f"{n}_"
        w(f"                case {num}: {{")
        if t.startswith("map<"):
            isint = "int32" in t
            w("                    std::string_view sub; if (!r.bytes(wt, sub)) return false;")
            w("                    faabric::proto::Reader er(sub); std::string k; " + ("int32_t v = 0;" if isint else "std::string v;"))
            w("                    uint32_t ef; int ewt;")
            w("                    while (er.next(ef, ewt)) {")
            w("                        if (ef == 1) { std::string_view s; if (!er.bytes(ewt, s)) return false; k.assign(s); }")
            if isint:
                w("                        else if (ef == 2) { uint64_t x; if (!er.varint(ewt, x)) return false; v = (int32_t)x; }")
            else:
                w("                        else if (ef == 2) { std::string_view s; if (!er.bytes(ewt, s)) return false; v.assign(s); }")
            w("                        else if (!er.skip(ewt)) return false;")
            w("                    }")
            w(f"                    {member}[k] = v;")
        elif rep:
            if t.startswith("msg:"):
                w("                    std::string_view sub; if (!r.bytes(wt, sub)) return false;")
                w(f"                    faabric::proto::Reader er(sub); {member}.emplace_back(); if (!{member}.back().decode(er)) return false;")
            elif t in ("string", "bytes"):
                w("                    std::string_view sub; if (!r.bytes(wt, sub)) return false;")
                w(f"                    {member}.emplace_back(sub);")
            else:
                ct = CPP_TYPE[t]
                w("                    if (wt == 2) {")
                w("                        std::string_view sub; if (!r.bytes(wt, sub)) return false;")
                w(f"                        faabric::proto::Reader er(sub); uint64_t x; while (er.rawVarint(x)) {{ {member}.push_back(({ct})x); }}")
                w("                    } else {")
                w(f"                        uint64_t x; if (!r.varint(wt, x)) return false; {member}.push_back(({ct})x);")
                w("                    }")
        elif t in ("string", "bytes"):
            w(f"                    std::string_view s; if (!r.bytes(wt, s)) return false; {member}.assign(s);")
        elif t.startswith("msg:"):
            w("                    std::string_view sub; if (!r.bytes(wt, sub)) return false;")
            w(f"                    faabric::proto::Reader er(sub); has_{lower(n)}_ = true; if (!{member}.decode(er)) return false;")
        elif t.startswith("enum:"):
            w(f"                    uint64_t x; if (!r.varint(wt, x)) return false; {member} = ({t[5:]})(int)x;")
        else:
            w(f"                    uint64_t x; if (!r.varint(wt, x)) return false; {member} = ({CPP_TYPE[t]})x;")
        w("                    break;")
        w("                }")
    w("                default:")
    w("                    if (!r.skip(wt)) return false;")
    w("            }")
    w("        }")
    w("        return r.ok();")
    w("    }")
    w("    std::string SerializeAsString() const { faabric::proto::Writer w; encode(w); return w.take(); }")
    w("    bool SerializeToString(std::string* out) const { *out = SerializeAsString(); return true; }")
    w("    size_t ByteSizeLong() const { return SerializeAsString().size(); }")
    w("    bool ParseFromArray(const void* data, int size) { Clear(); faabric::proto::Reader r(std::string_view((const char*)data, (size_t)size)); return decode(r); }")
    w("    bool ParseFromString(const std::string& s) { return ParseFromArray(s.data(), (int)s.size()); }")
    # JSON
    w("    void toJson(faabric::proto::JsonWriter& j) const")
    w("    {")
    w("        j.beginObject();")
    for f in m["fields"]:
        n, t, rep, js = f["name"], f["type"], f["rep"], f["json"]
        member = f"{n}_"
        if t.startswith("map<"):
            w(f"        if (!{member}.empty()) {{ j.key(\"{js}\"); j.beginObject(); for (const auto& kv : {member}) {{ j.key(kv.first); j.value(kv.second); }} j.endObject(); }}")
        elif rep:
            if t.startswith("msg:"):
                w(f"        if (!{member}.empty()) {{ j.key(\"{js}\"); j.beginArray(); for (const auto& v : {member}) {{ v.toJson(j); }} j.endArray(); }}")
            elif t == "bytes":
                w(f"        if (!{member}.empty()) {{ j.key(\"{js}\"); j.beginArray(); for (const auto& v : {member}) {{ j.bytesValue(v); }} j.endArray(); }}")
            else:
                w(f"        if (!{member}.empty()) {{ j.key(\"{js}\"); j.beginArray(); for (const auto& v : {member}) {{ j.value(v); }} j.endArray(); }}")
        elif t == "bytes":
            w(f"        if (!{member}.empty()) {{ j.key(\"{js}\"); j.bytesValue({member}); }}")
        elif t == "string":
            w(f"        if (!{member}.empty()) {{ j.key(\"{js}\"); j.value({member}); }}")
        elif t.startswith("msg:"):
            w(f"        if (has_{lower(n)}_) {{ j.key(\"{js}\"); {member}.toJson(j); }}")
        elif t.startswith("enum:"):
            # enums as ints (reference: always_print_enums_as_ints)
            w(f"        if ((int){member} != 0) {{ j.key(\"{js}\"); j.value((int64_t){member}); }}")
        elif t == "bool":
            w(f"        if ({member}) {{ j.key(\"{js}\"); j.value(true); }}")
        else:
            w(f"        if ({member} != 0) {{ j.key(\"{js}\"); j.value(({'uint64_t' if t.startswith('u') else 'int64_t'}){member}); }}")
    w("        j.endObject();")
    w("    }")
    w("    bool fromJson(const faabric::proto::JsonValue& v)")
    w("    {")
    w("        if (!v.isObject()) return false;")
    w("        for (const auto& [k, x] : v.members()) {")
    first = True
    for f in m["fields"]:
        n, t, rep, js = f["name"], f["type"], f["rep"], f["json"]
        member = f"{n}_"
        cond = f"k == \"{js}\"" + (f" || k == \"{n}\"" if js != n else "")
        w(f"            {'if' if first else 'else if'} ({cond}) {{")
        first = False
        if t.startswith("map<"):
            isint = "int32" in t
            w("                if (!x.isObject()) return false;")
            w(f"                for (const auto& [mk, mv] : x.members()) {{ {member}[mk] = " + ("(int32_t)mv.asInt();" if isint else "mv.asString();") + " }")
        elif rep:
            w("                if (!x.isArray()) return false;")
            if t.startswith("msg:"):
                w(f"                for (const auto& e : x.elements()) {{ {member}.emplace_back(); if (!{member}.back().fromJson(e)) return false; }}")
            elif t == "bytes":
                w(f"                for (const auto& e : x.elements()) {{ {member}.push_back(e.asBytes()); }}")
            elif t == "string":
                w(f"                for (const auto& e : x.elements()) {{ {member}.push_back(e.asString()); }}")
            else:
                w(f"                for (const auto& e : x.elements()) {{ {member}.push_back(({CPP_TYPE[t]})e.asInt()); }}")
        elif t == "bytes":
            w(f"                {member} = x.asBytes();")
        elif t == "string":
            w(f"                {member} = x.asString();")
        elif t.startswith("msg:"):
            w(f"                has_{lower(n)}_ = true; if (!{member}.fromJson(x)) return false;")
        elif t.startswith("enum:"):
            et = t[5:]
            w(f"                if (x.isString()) {{ {et} ev; if (!{et}_Parse(x.asString(), &ev)) return false; {member} = ev; }} else {{ {member} = ({et})(int)x.asInt(); }}")
        elif t == "bool":
            w(f"                {member} = x.asBool();")
        else:
            w(f"                {member} = ({CPP_TYPE[t]})x.asInt();")
        w("            }")
    w("        }")
    w("        return true;")
    w("    }")
    w(f"    bool operator==(const {name}& o) const {{ return SerializeAsString() == o.SerializeAsString(); }}")
    w(f"    bool operator!=(const {name}& o) const {{ return !(*this == o); }}")
    w()
    w("  private:")
    for f in m["fields"]:
        n, t, rep = f["name"], f["type"], f["rep"]
        member = f"{n}_"
        if t.startswith("map<"):
            vt = "int32_t" if "int32" in t else "std::string"
            w(f"    std::map<std::string, {vt}> {member};")
        elif rep:
            ct = t[4:] if t.startswith("msg:") else ("std::string" if t in ("string", "bytes") else CPP_TYPE[t])
            w(f"    faabric::proto::RepeatedField<{ct}> {member};")
        elif t in ("string", "bytes"):
            w(f"    std::string {member};")
        elif t.startswith("enum:"):
            w(f"    {t[5:]} {member} = ({t[5:]})0;")
        elif t.startswith("msg:"):
            w(f"    {t[4:]} {member};")
            w(f"    bool has_{lower(n)}_ = false;")
        elif t == "bool":
            w(f"    bool {member} = false;")
        else:
            w(f"    {CPP_TYPE[t]} {member} = 0;")
    w("};")
    # protobuf's namespace-level spellings of nested enums and their values
    # (Message_MessageType, Message_MessageType_EMPTY, ...): code written
    # against the reference uses both forms
    for en, vals in m.get("enums", []):
        w(f"using {name}_{en} = {name}::{en};")
        for vn, _ in vals:
            w(f"inline constexpr {name}_{en} {name}_{en}_{vn} = {name}::{vn};")
    w()


def gen_file(schema, guard):
    out = []
    out.append("// GENERATED by csrc/tools/gen_messages.py - do not edit by hand.")
    out.append("// Message classes with a protobuf-style accessor surface and a")
    out.append("// protobuf-wire-compatible binary codec + JSON codec.")
    out.append("#pragma once")
    out.append("")
    out.append("#include <faabric/proto/wire.h>")
    out.append("")
    out.append("#include <cstdint>")
    out.append("#include <map>")
    out.append("#include <string>")
    out.append("#include <utility>")
    out.append("#include <vector>")
    out.append("")
    for sch in schema:
        out.append(f"namespace {sch['namespace']} {{")
        out.append("")
        for m in sch["messages"]:
            gen_message(m, out)
        out.append(f"}} // namespace {sch['namespace']}")
        out.append("")
    return "\n".join(out)


if __name__ == "__main__":
    root = Path(__file__).resolve().parent.parent
    dst = root / "include" / "faabric" / "proto" / "faabric.pb.h"
    dst.write_text(gen_file([FAABRIC, PLANNER], "FAABRIC_PB_H"))
    # the reference includes <faabric/proto/faabric.pb.h> and
    # <faabric/planner/planner.pb.h>: provide the second as a forwarding header
    (root / "include" / "faabric" / "planner" / "planner.pb.h").write_text(
        "#pragma once\n#include <faabric/proto/faabric.pb.h>\n"
    )
    print("wrote", dst)
