"""HTTP clients (sync + async, retrying) and the on-disk configuration every other platform module builds on."""

from .client import APIClient, AsyncAPIClient, RetryPolicy
from .client import APIError, APITimeoutError, PaymentRequiredError, UnauthorizedError, ValidationError  # isort: skip
from .config import Config, ConfigModel

__all__ = ["APIClient", "AsyncAPIClient", "RetryPolicy", "Config", "ConfigModel",
           "APIError", "APITimeoutError", "PaymentRequiredError", "UnauthorizedError", "ValidationError"]  # fmt: skip
