"""Platform layer: the CLI and SDKs of the Prime Intellect compute platform (pods, disks, sandboxes,
environments hub, evals, hosted RL, inference, tunnels, MCP server), re-implemented on ONE shared
transport/config core instead of the reference's four copied ``core/`` packages
(reference: packages/*/src/*/core/{client,config}.py)."""

__version__ = "0.1.0"

# The reference's ``prime_cli`` package re-exports the transport and the sandbox SDK (packages/prime/src/prime_cli/__init__.py:1-43).
# Same names here, resolved on first use so that ``import prime_b200.platform`` stays cheap.
_CORE = ("APIClient", "APIError", "APITimeoutError", "AsyncAPIClient", "Config")
_SANDBOX = ("AsyncSandboxClient", "CommandRequest", "CommandResponse", "CommandTimeoutError", "CreateSandboxRequest", "Sandbox",
            "SandboxClient", "SandboxNotRunningError", "SandboxStatus", "UpdateSandboxRequest")  # fmt: skip
__all__ = [*_CORE, *_SANDBOX]


def __getattr__(name: str):
    if name in _CORE:
        from . import core

        return getattr(core, name)
    if name in _SANDBOX:
        from . import sandboxes

        return getattr(sandboxes, name)
    raise AttributeError(name)

