from .client import (  # noqa: F401
    APIClient,
    APIError,
    APITimeoutError,
    AsyncAPIClient,
    PaymentRequiredError,
    RetryPolicy,
    UnauthorizedError,
    ValidationError,
)
from .config import Config, ConfigModel  # noqa: F401
