"""What the evaluations SDK raises. All of them are ``APIError``s, so ``except APIError`` around SDK calls keeps working
(class names as in the reference: packages/prime-evals/src/prime_evals/exceptions.py:4-31)."""

from __future__ import annotations

from ..core.client import APIError


class EvalsAPIError(APIError):
    """Any failure of an evaluations call that is not a plain transport error."""


def _sub(name: str, doc: str) -> type[EvalsAPIError]:
    return type(name, (EvalsAPIError,), {"__doc__": doc, "__module__": __name__})


InvalidSampleError = _sub("InvalidSampleError", "A sample is not a JSON object or cannot be serialised.")
InvalidEvaluationError = _sub("InvalidEvaluationError", "create/update arguments that can never be valid (e.g. neither a run id nor an environment).")
EvaluationNotFoundError = _sub("EvaluationNotFoundError", "No evaluation with that id is visible to the caller.")
EnvironmentNotFoundError = _sub("EnvironmentNotFoundError", "The hub has no environment under that slug / name / id.")
