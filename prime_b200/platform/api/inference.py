"""OpenAI-compatible inference client: ``/models``, ``/models/{id}``, ``/chat/completions`` (+ SSE streaming)
(reference: packages/prime/src/prime_cli/api/inference.py:31-159).  Own timeouts: connect 10 s, read 600 s."""

from __future__ import annotations

import json
from typing import Any, Iterator

import httpx

from ..core.config import Config


class InferenceAPIError(Exception):
    pass


class InferencePaymentRequiredError(InferenceAPIError):
    pass


def _message(resp: httpx.Response) -> str:
    return resp.text.strip() or resp.reason_phrase or "Unknown error"


def _payment_message(resp: httpx.Response) -> str:
    try:
        return resp.json()["error"]["message"].strip()
    except Exception:
        return _message(resp)


def parse_sse_lines(lines) -> Iterator[dict[str, Any]]:
    """Yield JSON chunks from an SSE byte/str line stream; stops at ``[DONE]``; tolerates unprefixed lines."""
    for line in lines:
        if not line:
            continue
        data = (line[6:] if line.startswith("data: ") else line).strip()
        if not data:
            continue
        if data == "[DONE]":
            return
        try:
            yield json.loads(data)
        except json.JSONDecodeError:
            continue


class InferenceClient:
    def __init__(self, api_key: str | None = None, team_id: str | None = None, inference_url: str | None = None,
                 timeout: float | httpx.Timeout | None = None, transport: httpx.BaseTransport | None = None,
                 config: Config | None = None) -> None:  # fmt: skip
        self.config = config or Config()
        self.api_key = api_key or self.config.api_key
        if not self.api_key:
            raise InferenceAPIError("No API key. Run `prime config set-api-key` or set PRIME_API_KEY.")
        self.team_id = team_id if team_id is not None else self.config.team_id
        self.inference_url = (inference_url or self.config.inference_url).rstrip("/")
        headers = {"Authorization": f"Bearer {self.api_key}", "Content-Type": "application/json", "Accept": "application/json"}
        if self.team_id:
            headers["X-Prime-Team-ID"] = self.team_id
        self._client = httpx.Client(headers=headers, transport=transport,
                                    timeout=timeout or httpx.Timeout(connect=10.0, read=600.0, write=60.0, pool=60.0))  # fmt: skip

    def _fail(self, verb: str, url: str, resp: httpx.Response, model_id: str | None = None) -> None:
        code = resp.status_code
        if code == 402:
            raise InferencePaymentRequiredError(f"Payment required. {_payment_message(resp)}")
        if model_id is not None and code in (400, 404, 422):
            raise InferenceAPIError(f"Model '{model_id}' not found or unavailable ({verb} {url} → {code}).")
        raise InferenceAPIError(f"{verb} {url} failed: {code} {_message(resp)}")

    def list_models(self) -> dict[str, Any]:
        url = f"{self.inference_url}/models"
        r = self._client.get(url)
        if r.is_error:
            self._fail("GET", url, r)
        return r.json()

    def retrieve_model(self, model_id: str) -> dict[str, Any]:
        url = f"{self.inference_url}/models/{model_id}"
        r = self._client.get(url)
        if r.is_error:
            self._fail("GET", url, r, model_id)
        return r.json()

    def chat_completion(self, payload: dict[str, Any], stream: bool = False, headers: dict[str, str] | None = None):
        url = f"{self.inference_url}/chat/completions"
        extra = {"headers": headers} if headers else {}
        if not stream:
            r = self._client.post(url, json=payload, **extra)
            if r.is_error:
                self._fail("POST", url, r)
            return r.json()

        def gen() -> Iterator[dict[str, Any]]:
            with self._client.stream("POST", url, json=payload, **extra) as r:
                try:
                    r.raise_for_status()
                except httpx.HTTPStatusError as e:
                    e.response.read()  # a streamed body is not loaded yet: read it BEFORE the message is formatted from it
                    self._fail("POST", url, e.response)
                yield from parse_sse_lines(r.iter_lines())

        return gen()

    def close(self) -> None:
        self._client.close()
