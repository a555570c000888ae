"""Evals client (sync + async).

Parity: reference packages/prime-evals/src/prime_evals/evals.py:28-737 —
  * environment references resolve three ways: ``owner/name`` slug → lookup (never creates), bare name →
    resolve (get-or-create), id → lookup to verify; unknown environments are dropped, and the request is
    rejected only if nothing resolvable remains and no ``run_id`` was given;
  * ``push_samples``: batches sized by serialized bytes (2 MiB cap incl. envelope), oversize samples are
    skipped with a warning, 4 uploads in flight, each retried 5× (1–16 s back-off) on 429/transport errors.
The planning code (reference normalisation, payload, batching, retry predicate) is shared by both clients.
"""

from __future__ import annotations

import asyncio
import json
import threading
import time
import warnings
from concurrent.futures import ThreadPoolExecutor, as_completed
from typing import Any

import httpx

from ..core.client import APIError, user_agent
from .exceptions import EvalsAPIError, InvalidEvaluationError

ENVELOPE_BYTES = 20  # `{"samples": [` … `]}`
MAX_PAYLOAD_BYTES = 2 * 1024 * 1024
UPLOAD_ATTEMPTS = 5


def is_retryable_upload(exc: BaseException) -> bool:
    if isinstance(exc, httpx.HTTPStatusError):
        return exc.response.status_code == 429
    return isinstance(exc, httpx.RequestError)


def upload_backoff(attempt: int) -> float:
    return float(min(16, max(1, 2**attempt)))


def build_batches(samples: list[dict[str, Any]], max_payload_bytes: int = MAX_PAYLOAD_BYTES) -> tuple[list[list[dict]], int]:
    batches: list[list[dict]] = []
    cur: list[dict] = []
    used, skipped = ENVELOPE_BYTES, 0
    for i, s in enumerate(samples):
        size = len(json.dumps(s)) + 1
        if size + ENVELOPE_BYTES > max_payload_bytes:
            warnings.warn(f"Sample {i} exceeds maximum payload size ({size} bytes > {max_payload_bytes - ENVELOPE_BYTES} "
                          "bytes limit), skipping", stacklevel=3)  # fmt: skip
            skipped += 1
            continue
        if cur and used + size > max_payload_bytes:
            batches.append(cur)
            cur, used = [], ENVELOPE_BYTES
        cur.append(s)
        used += size
    if cur:
        batches.append(cur)
    return batches, skipped


def encode_batches(samples: list[dict[str, Any]], max_payload_bytes: int = MAX_PAYLOAD_BYTES) -> tuple[list[tuple[bytes, int]], int]:
    """Same cut as :func:`build_batches` (same size accounting, so the same batch boundaries), but every sample is serialised
    exactly ONCE: the request bodies are assembled from the per-sample encodings that the size check produced anyway. The
    reference measures a sample with ``json.dumps`` and then lets httpx serialise the whole batch again.
    → ([(request body, number of samples)], skipped)"""
    out: list[tuple[bytes, int]] = []
    cur: list[str] = []
    used, skipped = ENVELOPE_BYTES, 0

    def flush() -> None:
        out.append((('{"samples": [' + ", ".join(cur) + "]}").encode("ascii"), len(cur)))

    for i, s in enumerate(samples):
        enc = json.dumps(s)  # ensure_ascii: characters == bytes
        size = len(enc) + 1
        if size + ENVELOPE_BYTES > max_payload_bytes:
            warnings.warn(f"Sample {i} exceeds maximum payload size ({size} bytes > {max_payload_bytes - ENVELOPE_BYTES} "
                          "bytes limit), skipping", stacklevel=3)  # fmt: skip
            skipped += 1
            continue
        if cur and used + size > max_payload_bytes:
            flush()
            cur, used = [], ENVELOPE_BYTES
        cur.append(enc)
        used += size
    if cur:
        flush()
    return out, skipped


def normalise_env_ref(env: str | dict[str, str]) -> tuple[str, str, dict[str, str]] | None:
    """→ (kind, value, remaining-fields) with kind ∈ {slug, name, id}; None if unusable."""
    if isinstance(env, str):
        env = {"slug": env} if "/" in env else {"name": env}
    rest = dict(env)
    for kind in ("slug", "name", "id"):
        if kind in rest:
            value = rest.pop(kind) if kind != "id" else rest["id"]
            if kind == "slug" and "/" not in value:
                return None
            return kind, value, rest
    return None


def lookup_request(kind: str, value: str, team_id: str | None) -> tuple[str, dict[str, Any], str]:
    """→ (endpoint, body, human error) for one environment reference."""
    if kind == "slug":
        owner, name = value.split("/", 1)
        return ("/environmentshub/lookup", {"name": name, "team_slug": owner},
                f"Environment '{value}' does not exist in the hub. Please ensure the environment exists and you have access to it.")  # fmt: skip
    if kind == "name":
        body: dict[str, Any] = {"name": value}
        if team_id:
            body["team_id"] = team_id
        return ("/environmentshub/resolve", body,
                f"Environment '{value}' does not exist in the hub. Please push the environment first with: prime env push")  # fmt: skip
    return ("/environmentshub/lookup", {"id": value},
            f"Environment with ID '{value}' does not exist in the hub. Please verify the environment ID is correct.")  # fmt: skip


def evaluation_payload(name: str, resolved_envs: list[dict] | None, team_id: str | None, is_public: bool | None,
                       **fields: Any) -> dict[str, Any]:  # fmt: skip
    body = {"name": name, "environments": resolved_envs, **fields, "tags": fields.get("tags") or []}
    if team_id:
        body["team_id"] = team_id
    if is_public is not None:
        body["is_public"] = is_public
    return {k: v for k, v in body.items() if v is not None or k == "tags"}


def check_create_args(run_id: str | None, environments: list | None) -> None:
    if not run_id and not environments:
        raise InvalidEvaluationError(
            "Either 'run_id' or 'environments' must be provided. For environment evals, provide "
            "environments=[{'id': 'env-id', 'version_id': 'v1'}]")  # fmt: skip


def check_resolved(resolved: list, run_id: str | None) -> None:
    if not resolved and not run_id:
        raise InvalidEvaluationError("All provided environments lack valid identifiers (slug, name, or id). "
                                     "Either provide valid environment identifiers or provide a 'run_id'. ")  # fmt: skip


def list_params(env_name, suite_id, skip, limit, team_id) -> dict[str, Any]:
    p: dict[str, Any] = {"skip": skip, "limit": limit}
    if env_name:
        p["environment_name"] = env_name
    if suite_id:
        p["suite_id"] = suite_id
    if team_id:
        p["team_id"] = team_id
    return p


def update_payload(**fields: Any) -> dict[str, Any]:
    fields["tags"] = fields.get("tags") if fields.get("tags") is not None else []
    return {k: v for k, v in fields.items() if v is not None or k == "tags"}


_CREATE_FIELDS = ("suite_id", "run_id", "model_name", "dataset", "framework", "task_type", "description", "tags", "metadata", "metrics")


class _Common:
    def __init__(self, api_client: Any) -> None:
        self.client = api_client

    def _upload_target(self, evaluation_id: str) -> tuple[str, dict[str, str]]:
        headers = {"Content-Type": "application/json", "User-Agent": user_agent("prime-b200-evals")}
        if getattr(self.client, "api_key", None):
            headers["Authorization"] = f"Bearer {self.client.api_key}"
        return f"{self.client.base_url}/api/v1/evaluations/{evaluation_id}/samples", headers


class EvalsClient(_Common):
    _sleep = staticmethod(time.sleep)
    _http: httpx.Client | None = None
    _http_lock = threading.Lock()

    def _post(self, url: str, **kw: Any) -> httpx.Response:
        """Sample uploads share one pooled client (``httpx.post`` — what the reference calls — builds a client, i.e. a TLS context
        and a connection, per batch). Injectable for tests. Built once even when the four upload threads arrive together."""
        http = self._http
        if http is None or http.is_closed:
            with self._http_lock:
                http = self._http
                if http is None or http.is_closed:
                    http = self._http = httpx.Client(timeout=30.0)
        return http.post(url, **kw)

    def close(self) -> None:
        if self._http is not None and not self._http.is_closed:
            self._http.close()
        self._http = None

    def _resolve_one(self, kind: str, value: str) -> str:
        endpoint, body, err = lookup_request(kind, value, self.client.config.team_id)
        try:
            return self.client.post(endpoint, json=body)["data"]["id"]
        except APIError as e:
            raise EvalsAPIError(err) from e

    def _resolve_environments(self, environments: list[str | dict[str, str]]) -> list[dict[str, str]]:
        out = []
        for env in environments:
            ref = normalise_env_ref(env)
            if ref is None:
                continue
            kind, value, rest = ref
            try:
                rest["id"] = self._resolve_one(kind, value)
            except EvalsAPIError:
                continue  # unknown environments are dropped, the rest still get attached
            out.append(rest)
        return out

    def create_evaluation(self, name: str, environments: list | None = None, is_public: bool | None = None, **fields: Any) -> dict[str, Any]:
        unknown = set(fields) - set(_CREATE_FIELDS)
        if unknown:
            raise TypeError(f"unexpected arguments: {sorted(unknown)}")
        check_create_args(fields.get("run_id"), environments)
        resolved = None
        if environments:
            resolved = self._resolve_environments(environments)
            check_resolved(resolved, fields.get("run_id"))
        return self.client.request("POST", "/evaluations/", json=evaluation_payload(name, resolved, self.client.config.team_id, is_public, **fields))

    def _upload_batch(self, evaluation_id: str, batch: list[dict] | tuple[bytes, int]) -> int:
        """``batch``: an encoded (body, count) pair from :func:`encode_batches`, or a plain list of samples."""
        if isinstance(batch, tuple):
            body, count = batch
        else:
            body, count = json.dumps({"samples": batch}).encode("ascii"), len(batch)
        url, headers = self._upload_target(evaluation_id)
        headers = {**headers, "Content-Type": "application/json"}
        for attempt in range(UPLOAD_ATTEMPTS):
            try:
                r = self._post(url, content=body, headers=headers, timeout=30.0)
                r.raise_for_status()
                return count
            except (httpx.HTTPStatusError, httpx.RequestError) as e:
                if attempt + 1 < UPLOAD_ATTEMPTS and is_retryable_upload(e):
                    self._sleep(upload_backoff(attempt))
                    continue
                if isinstance(e, httpx.HTTPStatusError):
                    raise EvalsAPIError(f"HTTP {e.response.status_code}: {e.response.text}") from e
                raise EvalsAPIError(f"Request failed: {e}") from e
        raise AssertionError("unreachable")

    def push_samples(self, evaluation_id: str, samples: list[dict[str, Any]], max_payload_bytes: int = MAX_PAYLOAD_BYTES,
                     max_workers: int = 4) -> dict[str, Any]:  # fmt: skip
        if not samples:
            return {"samples_pushed": 0, "samples_skipped": 0}
        if max_workers < 1:
            raise ValueError("max_workers must be at least 1")
        batches, skipped = encode_batches(samples, max_payload_bytes)
        pushed, errors = 0, []
        with ThreadPoolExecutor(max_workers=max_workers) as pool:
            futs = {pool.submit(self._upload_batch, evaluation_id, b): i for i, b in enumerate(batches)}
            for f in as_completed(futs):
                try:
                    pushed += f.result()
                except Exception as e:
                    errors.append(f"Batch {futs[f] + 1}: {e}")
        if errors:
            raise EvalsAPIError(f"Failed to push samples: {'; '.join(errors)}")
        return {"samples_pushed": pushed, "samples_skipped": skipped}

    def finalize_evaluation(self, evaluation_id: str, metrics: dict[str, Any] | None = None) -> dict[str, Any]:
        return self.client.request("POST", f"/evaluations/{evaluation_id}/finalize", json={"metrics": metrics} if metrics else {})

    def list_evaluations(self, env_name: str | None = None, suite_id: str | None = None, skip: int = 0, limit: int = 50, *,
                         team_id: str | None = None) -> dict[str, Any]:  # fmt: skip
        return self.client.request("GET", "/evaluations/", params=list_params(env_name, suite_id, skip, limit, team_id))

    def get_evaluation(self, evaluation_id: str) -> dict[str, Any]:
        return self.client.request("GET", f"/evaluations/{evaluation_id}")

    def update_evaluation(self, evaluation_id: str, **fields: Any) -> dict[str, Any]:
        return self.client.request("PUT", f"/evaluations/{evaluation_id}", json=update_payload(**fields))

    def get_samples(self, evaluation_id: str, page: int = 1, limit: int = 100) -> dict[str, Any]:
        return self.client.request("GET", f"/evaluations/{evaluation_id}/samples", params={"page": page, "limit": limit})


class AsyncEvalsClient(_Common):
    _sleep = staticmethod(asyncio.sleep)

    def __init__(self, api_client: Any = None, http: httpx.AsyncClient | None = None, *, api_key: str | None = None) -> None:
        """``AsyncEvalsClient(api_key="…")`` / ``AsyncEvalsClient("…")`` as in the reference (evals.py:383-384), or an explicit client."""
        if api_client is None or isinstance(api_client, str):
            from ..core.client import AsyncAPIClient

            api_client = AsyncAPIClient(api_key=api_key or api_client, user_agent=user_agent("prime-b200-evals"))
        super().__init__(api_client)
        self._http = http

    async def _resolve_one(self, kind: str, value: str) -> str:
        endpoint, body, err = lookup_request(kind, value, self.client.config.team_id)
        try:
            return (await self.client.post(endpoint, json=body))["data"]["id"]
        except APIError as e:
            raise EvalsAPIError(err) from e

    async def _resolve_environments(self, environments: list[str | dict[str, str]]) -> list[dict[str, str]]:
        out = []
        for env in environments:
            ref = normalise_env_ref(env)
            if ref is None:
                continue
            kind, value, rest = ref
            try:
                rest["id"] = await self._resolve_one(kind, value)
            except EvalsAPIError:
                continue
            out.append(rest)
        return out

    async def create_evaluation(self, name: str, environments: list | None = None, is_public: bool | None = None, **fields: Any) -> dict[str, Any]:
        unknown = set(fields) - set(_CREATE_FIELDS)
        if unknown:
            raise TypeError(f"unexpected arguments: {sorted(unknown)}")
        check_create_args(fields.get("run_id"), environments)
        resolved = None
        if environments:
            resolved = await self._resolve_environments(environments)
            check_resolved(resolved, fields.get("run_id"))
        body = evaluation_payload(name, resolved, self.client.config.team_id, is_public, **fields)
        return await self.client.request("POST", "/evaluations/", json=body)

    async def _upload_batch(self, http: httpx.AsyncClient, evaluation_id: str, batch: list[dict] | tuple[bytes, int]) -> int:
        if isinstance(batch, tuple):
            body, count = batch
        else:
            body, count = json.dumps({"samples": batch}).encode("ascii"), len(batch)
        url, headers = self._upload_target(evaluation_id)
        headers = {**headers, "Content-Type": "application/json"}
        for attempt in range(UPLOAD_ATTEMPTS):
            try:
                r = await http.post(url, content=body, headers=headers, timeout=30.0)
                r.raise_for_status()
                return count
            except (httpx.HTTPStatusError, httpx.RequestError) as e:
                if attempt + 1 < UPLOAD_ATTEMPTS and is_retryable_upload(e):
                    await self._sleep(upload_backoff(attempt))
                    continue
                if isinstance(e, httpx.HTTPStatusError):
                    raise EvalsAPIError(f"HTTP {e.response.status_code}: {e.response.text}") from e
                raise EvalsAPIError(f"Request failed: {e}") from e
        raise AssertionError("unreachable")

    async def push_samples(self, evaluation_id: str, samples: list[dict[str, Any]], max_payload_bytes: int = MAX_PAYLOAD_BYTES,
                           max_concurrent: int = 4) -> dict[str, Any]:  # fmt: skip
        if not samples:
            return {"samples_pushed": 0, "samples_skipped": 0}
        if max_concurrent < 1:
            raise ValueError("max_concurrent must be at least 1")
        batches, skipped = encode_batches(samples, max_payload_bytes)
        sem = asyncio.Semaphore(max_concurrent)
        own = self._http is None
        http = self._http or httpx.AsyncClient()

        async def one(i: int, b: tuple[bytes, int]) -> int | str:
            async with sem:
                try:
                    return await self._upload_batch(http, evaluation_id, b)
                except Exception as e:
                    return f"Batch {i + 1}: {e}"

        try:
            results = await asyncio.gather(*(one(i, b) for i, b in enumerate(batches)))
        finally:
            if own:
                await http.aclose()
        errors = [r for r in results if isinstance(r, str)]
        if errors:
            raise EvalsAPIError(f"Failed to push samples: {'; '.join(errors)}")
        return {"samples_pushed": sum(r for r in results if isinstance(r, int)), "samples_skipped": skipped}

    async def finalize_evaluation(self, evaluation_id: str, metrics: dict[str, Any] | None = None) -> dict[str, Any]:
        return await self.client.request("POST", f"/evaluations/{evaluation_id}/finalize", json={"metrics": metrics} if metrics else {})

    async def list_evaluations(self, env_name: str | None = None, suite_id: str | None = None, skip: int = 0, limit: int = 50, *,
                               team_id: str | None = None) -> dict[str, Any]:  # fmt: skip
        return await self.client.request("GET", "/evaluations/", params=list_params(env_name, suite_id, skip, limit, team_id))

    async def get_evaluation(self, evaluation_id: str) -> dict[str, Any]:
        return await self.client.request("GET", f"/evaluations/{evaluation_id}")

    async def update_evaluation(self, evaluation_id: str, **fields: Any) -> dict[str, Any]:
        return await self.client.request("PUT", f"/evaluations/{evaluation_id}", json=update_payload(**fields))

    async def get_samples(self, evaluation_id: str, page: int = 1, limit: int = 100) -> dict[str, Any]:
        return await self.client.request("GET", f"/evaluations/{evaluation_id}/samples", params={"page": page, "limit": limit})

    async def aclose(self) -> None:
        """Close the transport (reference: packages/prime-evals/src/prime_evals/evals.py:727-737)."""
        closer = getattr(self.client, "aclose", None)
        if closer is not None:
            await closer()
        if self._http is not None:
            await self._http.aclose()

    async def __aenter__(self) -> "AsyncEvalsClient":
        return self

    async def __aexit__(self, *exc: Any) -> None:
        await self.aclose()

