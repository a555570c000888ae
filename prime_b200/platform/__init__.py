"""Platform layer: the CLI and SDKs of the Prime Intellect compute platform (pods, disks, sandboxes,
environments hub, evals, hosted RL, inference, tunnels, MCP server), re-implemented on ONE shared
transport/config core instead of the reference's four copied ``core/`` packages
(reference: packages/*/src/*/core/{client,config}.py)."""

__version__ = "0.1.0"
