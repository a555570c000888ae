"""(reference: packages/prime-evals/src/prime_evals/exceptions.py:4-31)"""

from ..core.client import APIError


class EvalsAPIError(APIError):
    pass


class EvaluationNotFoundError(EvalsAPIError):
    pass


class InvalidEvaluationError(EvalsAPIError):
    pass


class InvalidSampleError(EvalsAPIError):
    pass
