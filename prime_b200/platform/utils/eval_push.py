"""Upload the newest local ``outputs/evals/<env>--<model>/<run>/`` (metadata.json + results.jsonl) to the evals hub
(reference: packages/prime/src/prime_cli/utils/eval_push.py:17-220)."""

from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Any

from ..core import APIClient, Config
from ..evals import EvalsAPIError, EvalsClient
from .env_metadata import find_environment_metadata
from .plain import get_console

console = get_console()


def get_eval_viewer_url(evaluation_id: str) -> str:
    return f"{Config(writable=False).frontend_url}/dashboard/evaluations/{evaluation_id}"


def load_results_jsonl(path: Path) -> list[dict]:
    good, bad = [], []
    for n, line in enumerate(Path(path).read_text(encoding="utf-8").splitlines(), 1):
        if not line.strip():
            continue
        try:
            obj = json.loads(line)
        except json.JSONDecodeError:
            bad.append(f"line {n}: invalid JSON")
            continue
        if isinstance(obj, dict):
            good.append(obj)
        else:
            bad.append(f"line {n}: expected dict, got {type(obj).__name__}")
    if bad:
        console.print(f"[yellow]Warning: Skipped {len(bad)} invalid lines in results.jsonl ({', '.join(bad[:5])}{', ...' if len(bad) > 5 else ''})[/yellow]")
    return good


def find_latest_run_dir(env_name: str, model: str, root: Path = Path(".")) -> Path:
    """``environments/<module>/outputs/evals/...`` when a local checkout exists, else ``./outputs/evals/...``."""
    key = f"{env_name}--{model.replace('/', '--')}"
    local = root / "environments" / env_name.replace("-", "_")
    base = (local if local.exists() else root) / "outputs" / "evals" / key
    if not base.exists():
        raise FileNotFoundError(f"Evaluation output directory not found: {base}")
    runs = [d for d in base.iterdir() if d.is_dir()]
    if not runs:
        raise FileNotFoundError(f"No evaluation results found in {base}")
    return max(runs, key=lambda d: d.stat().st_mtime)


def resolve_upstream(env_name: str, env_path: Path | None, upstream_slug: str | None) -> tuple[str | None, str | None]:
    """→ (slug, environment_id)."""
    if upstream_slug:
        return upstream_slug, None
    md = find_environment_metadata(env_name=env_name, env_path=env_path, module_name=env_name.replace("-", "_")) or {}
    slug = f"{md['owner']}/{md['name']}" if md.get("owner") and md.get("name") else None
    return slug, md.get("environment_id")


def to_hub_samples(samples: list[dict]) -> list[dict[str, Any]]:
    return [{"example_id": s.get("id", 0), "reward": s.get("reward", 0.0), **{k: v for k, v in s.items() if k not in ("id", "reward")}} for s in samples]


def push_eval_results_to_hub(env_name: str, model: str, job_id: str, env_path: Path | None = None,
                             upstream_slug: str | None = None, client: APIClient | None = None) -> str | None:  # fmt: skip
    run_dir = find_latest_run_dir(env_name, model)
    for needed in ("metadata.json", "results.jsonl"):
        if not (run_dir / needed).exists():
            raise FileNotFoundError(f"{needed} not found in {run_dir}")
    metadata = json.loads((run_dir / "metadata.json").read_text(encoding="utf-8"))
    samples = load_results_jsonl(run_dir / "results.jsonl")
    slug, env_id = resolve_upstream(env_name, env_path, upstream_slug)
    if not slug and not env_id:
        console.print("[yellow]No upstream environment found. Evaluation results will not be uploaded or viewable on the platform. "
                      "Use `prime env push` to set an upstream, or `--env-path` to point at the environment.[/yellow]")  # fmt: skip
        return None
    console.print(f"\n[blue]Uploading evaluation results, using upstream: {slug or env_id}[/blue]")
    client = client or APIClient()
    if env_id:
        envs = [{"id": env_id}]
    else:
        envs = [{"slug": slug}]
        try:
            owner, name = slug.split("/", 1)
            resp = client.get(f"/environmentshub/{owner}/{name}/@latest")
            found = (resp.get("data", resp) or {}).get("id")
            if found:
                envs = [{"id": found}]
        except Exception:
            pass
    metrics = {k: v for k, v in metadata.items() if k.startswith("avg_")}
    evals = EvalsClient(client)
    created = evals.create_evaluation(
        name=f"{env_name}--{model}--{datetime.now():%Y%m%d_%H%M%S}", environments=envs, model_name=model, dataset=env_name,
        framework="verifiers", task_type=metadata.get("task_type"), metadata={"framework": "verifiers", "job_id": job_id, **metadata},
        metrics=metrics, is_public=False,
    )  # fmt: skip
    eval_id = created.get("evaluation_id")
    if not eval_id:
        raise EvalsAPIError("Failed to get evaluation ID from create_evaluation response")
    if samples:
        evals.push_samples(eval_id, to_hub_samples(samples))
    evals.finalize_evaluation(eval_id, metrics=metrics)
    url = get_eval_viewer_url(eval_id)
    console.print(f"[green]✓ Successfully uploaded evaluation results[/green]\n\n[green]View results at:[/green]\n  [link={url}]{url}[/link]")
    return eval_id
