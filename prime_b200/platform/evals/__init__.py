"""Evaluation results SDK: create → push samples (size-adaptive batches, concurrent, retried) → finalize."""

from ..core import APIError, APITimeoutError, Config, PaymentRequiredError, UnauthorizedError  # noqa: F401
from ..core.client import APIClient, AsyncAPIClient  # noqa: F401
from .evals import AsyncEvalsClient, EvalsClient, build_batches, encode_batches  # noqa: F401
from .exceptions import EnvironmentNotFoundError, EvalsAPIError, EvaluationNotFoundError, InvalidEvaluationError, InvalidSampleError  # noqa: F401
from .models import (  # noqa: F401
    CreateEvaluationRequest,
    Environment,
    EnvironmentReference,
    Evaluation,
    EvaluationListResponse,
    EvaluationStatus,
    FinalizeEvaluationRequest,
    PushSamplesRequest,
    Sample,
    SamplesResponse,
)

__version__ = "0.1.0"
