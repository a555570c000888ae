"""Runtime protobuf classes for ``command_session.proto`` built from a compact table — no protoc, no
generated ``_pb2`` blob.  ``MESSAGES`` mirrors proto/command_session.proto one-to-one."""

from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
PKG = "command_session"

# field spec: (number, name, type, extra) ; extra: "rep" | "opt" | ("oneof", name) | None ; type "msg:<Name>" / "enum:<Name>"
_T = {"string": F.TYPE_STRING, "uint32": F.TYPE_UINT32, "sint32": F.TYPE_SINT32, "bool": F.TYPE_BOOL, "bytes": F.TYPE_BYTES}

MESSAGES: dict = {
    "PTY": {"fields": [(1, "size", "msg:PTY.Size", None)], "nested": {"Size": {"fields": [(1, "cols", "uint32", None), (2, "rows", "uint32", None)]}}},
    "CommandSpec": {"fields": [(1, "cmd", "string", None), (2, "args", "string", "rep"), (3, "envs", "map", None), (4, "cwd", "string", "opt")]},
    "CommandSessionSelector": {"fields": [(1, "pid", "uint32", ("oneof", "selector")), (2, "tag", "string", ("oneof", "selector"))]},
    "CommandSessionInfo": {"fields": [(1, "command", "msg:CommandSpec", None), (2, "pid", "uint32", None), (3, "tag", "string", "opt")]},
    "CommandInput": {"fields": [(1, "stdin", "bytes", ("oneof", "input")), (2, "pty", "bytes", ("oneof", "input"))]},
    "CommandSessionEvent": {
        "fields": [(1, "start", "msg:CommandSessionEvent.StartEvent", ("oneof", "event")),
                   (2, "data", "msg:CommandSessionEvent.DataEvent", ("oneof", "event")),
                   (3, "end", "msg:CommandSessionEvent.EndEvent", ("oneof", "event")),
                   (4, "keepalive", "msg:CommandSessionEvent.KeepAlive", ("oneof", "event"))],
        "nested": {
            "StartEvent": {"fields": [(1, "pid", "uint32", None)]},
            "DataEvent": {"fields": [(1, "stdout", "bytes", ("oneof", "output")), (2, "stderr", "bytes", ("oneof", "output")),
                                     (3, "pty", "bytes", ("oneof", "output"))]},
            "EndEvent": {"fields": [(1, "exit_code", "sint32", None), (2, "exited", "bool", None), (3, "status", "string", None),
                                    (4, "error", "string", "opt")]},
            "KeepAlive": {"fields": []},
        },
    },
    "ListRequest": {"fields": []},
    "ListResponse": {"fields": [(1, "sessions", "msg:CommandSessionInfo", "rep")]},
    "StartRequest": {"fields": [(1, "command", "msg:CommandSpec", None), (2, "pty", "msg:PTY", "opt"), (3, "tag", "string", "opt"),
                                (4, "stdin", "bool", "opt")]},
    "StartResponse": {"fields": [(1, "event", "msg:CommandSessionEvent", None)]},
    "ConnectRequest": {"fields": [(1, "session", "msg:CommandSessionSelector", None)]},
    "ConnectResponse": {"fields": [(1, "event", "msg:CommandSessionEvent", None)]},
    "UpdateRequest": {"fields": [(1, "session", "msg:CommandSessionSelector", None), (2, "pty", "msg:PTY", "opt")]},
    "UpdateResponse": {"fields": []},
    "SendInputRequest": {"fields": [(1, "session", "msg:CommandSessionSelector", None), (2, "input", "msg:CommandInput", None)]},
    "SendInputResponse": {"fields": []},
    "StreamInputRequest": {
        "fields": [(1, "start", "msg:StreamInputRequest.StartEvent", ("oneof", "event")),
                   (2, "data", "msg:StreamInputRequest.DataEvent", ("oneof", "event")),
                   (3, "keepalive", "msg:StreamInputRequest.KeepAlive", ("oneof", "event"))],
        "nested": {"StartEvent": {"fields": [(1, "session", "msg:CommandSessionSelector", None)]},
                   "DataEvent": {"fields": [(2, "input", "msg:CommandInput", None)]}, "KeepAlive": {"fields": []}},
    },
    "StreamInputResponse": {"fields": []},
    "SendSignalRequest": {"fields": [(1, "session", "msg:CommandSessionSelector", None), (2, "signal", "enum:Signal", None)]},
    "SendSignalResponse": {"fields": []},
}  # fmt: skip
ENUMS = {"Signal": [("SIGNAL_UNSPECIFIED", 0), ("SIGNAL_SIGKILL", 9), ("SIGNAL_SIGTERM", 15)]}
SERVICE = {  # name: (input, output, client_streaming, server_streaming)
    "List": ("ListRequest", "ListResponse", False, False),
    "Connect": ("ConnectRequest", "ConnectResponse", False, True),
    "Start": ("StartRequest", "StartResponse", False, True),
    "Update": ("UpdateRequest", "UpdateResponse", False, False),
    "StreamInput": ("StreamInputRequest", "StreamInputResponse", True, False),
    "SendInput": ("SendInputRequest", "SendInputResponse", False, False),
    "SendSignal": ("SendSignalRequest", "SendSignalResponse", False, False),
}


def _fill(msg: descriptor_pb2.DescriptorProto, name: str, spec: dict) -> None:
    msg.name = name
    oneofs: dict[str, int] = {}
    for number, fname, ftype, extra in spec["fields"]:
        f = msg.field.add()
        f.name, f.number, f.label = fname, number, F.LABEL_OPTIONAL
        f.json_name = "".join(p if i == 0 else p.capitalize() for i, p in enumerate(fname.split("_")))
        if ftype == "map":
            entry = msg.nested_type.add()
            entry.name = "".join(p.capitalize() for p in fname.split("_")) + "Entry"
            entry.options.map_entry = True
            for n, (k, t) in enumerate((("key", F.TYPE_STRING), ("value", F.TYPE_STRING)), 1):
                ef = entry.field.add()
                ef.name, ef.number, ef.label, ef.type, ef.json_name = k, n, F.LABEL_OPTIONAL, t, k
            f.type, f.label, f.type_name = F.TYPE_MESSAGE, F.LABEL_REPEATED, f".{PKG}.{name_path(msg, name)}.{entry.name}"
        elif ftype.startswith("msg:"):
            f.type, f.type_name = F.TYPE_MESSAGE, f".{PKG}.{ftype[4:]}"
        elif ftype.startswith("enum:"):
            f.type, f.type_name = F.TYPE_ENUM, f".{PKG}.{ftype[5:]}"
        else:
            f.type = _T[ftype]
        if extra == "rep":
            f.label = F.LABEL_REPEATED
        elif extra == "opt":
            oneofs[f"_{fname}"] = len(msg.oneof_decl)
            msg.oneof_decl.add().name = f"_{fname}"
            f.oneof_index, f.proto3_optional = oneofs[f"_{fname}"], True
        elif isinstance(extra, tuple):
            if extra[1] not in oneofs:
                oneofs[extra[1]] = len(msg.oneof_decl)
                msg.oneof_decl.add().name = extra[1]
            f.oneof_index = oneofs[extra[1]]
    for nname, nspec in spec.get("nested", {}).items():
        _PATHS[id(nested := msg.nested_type.add())] = f"{name_path(msg, name)}.{nname}"
        _fill(nested, nname, nspec)


_PATHS: dict[int, str] = {}


def name_path(msg, name: str) -> str:
    return _PATHS.get(id(msg), name)


def build_file_descriptor() -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "command_session/command_session.proto", PKG, "proto3"
    # real oneofs must precede synthetic (proto3 optional) ones: emit messages in two passes per message
    for name, spec in MESSAGES.items():
        ordered = dict(spec)
        ordered["fields"] = sorted(spec["fields"], key=lambda t: (t[3] == "opt", 0))
        m = fd.message_type.add()
        _fill(m, name, ordered)
        m.field.sort(key=lambda f: f.number)
    for ename, values in ENUMS.items():
        e = fd.enum_type.add()
        e.name = ename
        for vn, vv in values:
            v = e.value.add()
            v.name, v.number = vn, vv
    svc = fd.service.add()
    svc.name = "CommandSession"
    for mname, (i, o, cs, ss) in SERVICE.items():
        me = svc.method.add()
        me.name, me.input_type, me.output_type = mname, f".{PKG}.{i}", f".{PKG}.{o}"
        me.client_streaming, me.server_streaming = cs, ss
    return fd


_pool = descriptor_pool.DescriptorPool()
_file = _pool.Add(build_file_descriptor()) if hasattr(_pool, "Add") else None
if _file is None:  # pragma: no cover - older protobuf API
    _pool.AddSerializedFile(build_file_descriptor().SerializeToString())


def message_class(name: str):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName(f"{PKG}.{name}"))


StartRequest = message_class("StartRequest")
StartResponse = message_class("StartResponse")
CommandSpec = message_class("CommandSpec")
CommandSessionEvent = message_class("CommandSessionEvent")
