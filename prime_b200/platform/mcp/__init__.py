"""MCP (Model Context Protocol) tool server exposing compute operations to AI agents over stdio.

Same export surface as the reference package (packages/prime-mcp-server/src/prime_mcp/__init__.py:1-13) — ``mcp``,
``make_prime_request`` and the three tool modules — resolved lazily so that importing the package does not pull in FastMCP.
"""

__version__ = "0.1.0"
__all__ = ["mcp", "make_prime_request", "availability", "pods", "ssh"]


def __getattr__(name: str):
    if name == "mcp":
        from .server import mcp

        return mcp
    if name == "make_prime_request":
        from .client import make_prime_request

        return make_prime_request
    if name in ("availability", "pods", "ssh"):
        import importlib

        return importlib.import_module(f"{__name__}.tools.{name}")
    raise AttributeError(name)
