"""(reference: packages/prime-evals/src/prime_evals/exceptions.py:4-31)"""

from ..core.client import APIError


class EvalsAPIError(APIError):
    pass


class EnvironmentNotFoundError(EvalsAPIError):
    """The hub has no environment under that slug / name / id (reference: packages/prime-evals/src/prime_evals/exceptions.py:10-13)."""


class EvaluationNotFoundError(EvalsAPIError):
    pass


class InvalidEvaluationError(EvalsAPIError):
    pass


class InvalidSampleError(EvalsAPIError):
    pass
