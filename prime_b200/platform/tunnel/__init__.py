"""Expose a local port through a managed ``frpc`` child process."""

from ..core import Config  # noqa: F401
from .binary import FRPC_VERSION, get_frpc_path  # noqa: F401
from .client import TunnelClient  # noqa: F401
from .exceptions import (  # noqa: F401
    BinaryDownloadError,
    TunnelAuthError,
    TunnelConnectionError,
    TunnelError,
    TunnelLimitReachedError,
    TunnelTimeoutError,
)
from .models import TunnelInfo  # noqa: F401
from .tunnel import Tunnel  # noqa: F401

__version__ = "0.1.0"
