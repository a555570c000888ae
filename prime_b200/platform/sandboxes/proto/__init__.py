"""Wire schema of the VM command-session service: ``command_session.proto`` is the source of truth; the message classes are built
from it at import time by ``..rpc_schema`` (no protoc step, no generated blob).  ``command_session_pb2`` is the name code written
against a generated module expects (reference: packages/prime-sandboxes/src/prime_sandboxes/_proto/command_session/)."""

from .. import rpc_schema as command_session_pb2  # noqa: F401
