"""MCP (Model Context Protocol) tool server exposing compute operations to AI agents over stdio."""

__version__ = "0.1.0"
