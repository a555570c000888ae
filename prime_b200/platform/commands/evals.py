"""``prime eval [run] ENV …`` plus list / get / samples / push / tui / logs / stop
(reference: packages/prime/src/prime_cli/commands/evals.py:85-1503).

Local runs are a pass-through to the verifiers toolkit (see ``verifiers_bridge.run_eval_passthrough``); ``--hosted``
submits ``POST /hosted-evaluations`` instead — from the CLI flags or from a TOML file with ``[[eval]]`` tables merged
over its top-level defaults — grouping environments that share identical settings into one request.
"""

from __future__ import annotations

import json
import re
import time
from pathlib import Path
from typing import Any, Optional

import typer
from click.core import ParameterSource

from ..core import APIClient, APIError, Config
from ..evals import EvalsClient
from ..utils.config import load_toml
from ..utils.display import output_data_as_json
from ..utils.env_metadata import find_environment_metadata
from ..utils.eval_push import get_eval_viewer_url, load_results_jsonl
from ..utils.hosted_eval import EvalStatus, HostedEvalConfig, clean_logs, get_new_log_lines
from ..utils.json_help import json_output_help, list_json_help
from ..utils.time_utils import format_time_ago
from ..verifiers_bridge import (
    DEFAULT_ENV_DIR_PATH,
    DEFAULT_MODEL,
    is_config_target,
    is_help_request,
    parse_value_option,
    print_eval_run_help,
    resolve_environment_reference,
    run_eval_passthrough,
    run_eval_tui,
    split_owner_and_name,
)
from ._common import api, console, emit, fail, handle_errors, make_app

app = make_app("Run and manage evaluations", default_cmd="run")

LOGS_TAIL, LOGS_POLL_S, RUN_POLL_S = 1000, 5.0, 10.0
DEFAULT_NUM_EXAMPLES, DEFAULT_ROLLOUTS = 5, 3
RATE_LIMIT_AFTER, RATE_LIMIT_WAIT_S, RETRY_WAIT_S, HOSTED_LOGS_STATUS_UPDATE_EVERY_POLLS = 3, 30, 10, 6
EXAMPLE = "prime eval run gsm8k -n 10"
HOSTED_ALLOWED = {"env_id", "env_args", "env_dir_path", "endpoints_path", "endpoint_id", "model", "num_examples", "rollouts_per_example",
                  "timeout_minutes", "allow_sandbox_access", "allow_instances_access", "sampling_args", "api_base_url", "api_key_var", "eval_name"}  # fmt: skip
HOSTED_TYPES: dict[str, tuple[type, str]] = {
    "env_dir_path": (str, "a non-empty string"), "num_examples": (int, "an integer"), "rollouts_per_example": (int, "an integer"),
    "timeout_minutes": (int, "an integer"), "allow_sandbox_access": (bool, "a boolean"), "allow_instances_access": (bool, "a boolean"),
    "api_base_url": (str, "a non-empty string"), "api_key_var": (str, "a non-empty string"), "eval_name": (str, "a non-empty string"),
}  # fmt: skip


# ------------------------------------------------------------------------------------------------- option parsing
def parse_json_object_option(raw: str | None, option: str) -> dict[str, Any] | None:
    if raw is None:
        return None
    try:
        v = json.loads(raw)
    except json.JSONDecodeError as e:
        raise fail(f"{option} must be valid JSON: {e}")
    if not isinstance(v, dict):
        raise fail(f"{option} must be a JSON object")
    return v


def parse_string_map_option(raw: str | None, option: str) -> dict[str, str] | None:
    v = parse_json_object_option(raw, option)
    if v is not None and not all(isinstance(k, str) and isinstance(x, str) for k, x in v.items()):
        raise fail(f"{option} must map strings to strings")
    return v


def freeze(value: Any) -> Any:
    """Hashable view of a JSON value (for grouping identical settings)."""
    if isinstance(value, dict):
        return tuple(sorted((k, freeze(v)) for k, v in value.items()))
    if isinstance(value, list):
        return tuple(freeze(v) for v in value)
    return value


def resolve_config_model(cfg: dict[str, Any], config_path: Path) -> str:
    ep, model = cfg.get("endpoint_id"), cfg.get("model")
    if ep is not None and model is not None:
        raise fail("hosted eval config cannot set both `endpoint_id` and `model`")
    if ep is None:
        if model is None:
            return DEFAULT_MODEL
        if type(model) is not str or not model:
            raise fail("`model` must be a non-empty string")
        return model
    if type(ep) is not str or not ep:
        raise fail("`endpoint_id` must be a non-empty string")
    path = cfg.get("endpoints_path", "./configs/endpoints.toml")
    if type(path) is not str or not path:
        raise fail("`endpoints_path` must be a non-empty string")
    if "endpoints_path" in cfg and not Path(path).is_absolute():
        path = str((config_path.parent / path).resolve())
    try:
        from verifiers.utils.eval_utils import load_endpoints, resolve_endpoints_file
    except ImportError:
        raise fail("verifiers is required to resolve `endpoint_id`. Install it or use `model` instead.")
    f = resolve_endpoints_file(path)
    if f is None or f.suffix != ".toml":
        raise fail("`endpoint_id` requires an endpoints.toml registry via `endpoints_path`")
    registry = load_endpoints(path)
    if ep not in registry:
        raise fail(f"endpoint_id '{ep}' not found in {path}")
    models = {e["model"] for e in registry[ep]}
    if len(models) != 1:
        raise fail(f"endpoint_id '{ep}' resolves to multiple models: {sorted(models)}")
    return registry[ep][0]["model"]


def validate_hosted_entry(merged: dict[str, Any], config_path: Path) -> dict[str, Any]:
    extra = sorted(set(merged) - HOSTED_ALLOWED)
    if extra:
        raise fail("hosted eval config does not support: " + ", ".join(f"`{k}`" for k in extra))
    if type(merged.get("env_id")) is not str or not merged["env_id"]:
        raise fail("hosted eval config requires a non-empty `env_id`")
    ea = merged.get("env_args")
    if ea is not None and not (isinstance(ea, dict) and all(isinstance(k, str) and isinstance(v, str) for k, v in ea.items())):
        raise fail("`env_args` must be a table of strings")
    if merged.get("sampling_args") is not None:
        if not isinstance(merged["sampling_args"], dict):
            raise fail("`sampling_args` must be a table")
        try:  # it travels as JSON: TOML dates/times (and anything else json cannot encode) are caught here, not by the server
            json.dumps(merged["sampling_args"])
        except (TypeError, ValueError) as e:
            raise fail(f"`sampling_args` must be JSON-serializable ({e})")
    for field, (typ, desc) in HOSTED_TYPES.items():
        v = merged.get(field)
        if v is not None and (type(v) is not typ or (typ is str and not v)):
            raise fail(f"`{field}` must be {desc}")
    merged["model"] = resolve_config_model(merged, config_path)
    return merged


def load_hosted_eval_configs(config_path: str) -> list[dict[str, Any]]:
    raw = load_toml(config_path, console)
    entries = raw.get("eval")
    if not isinstance(entries, list) or not entries:
        raise fail("hosted eval config must use [[eval]] and contain at least one entry")
    out = []
    for e in entries:
        if not isinstance(e, dict):
            raise fail("[[eval]] must be a TOML table")
        merged = {k: v for k, v in raw.items() if k != "eval"}
        merged.update(e)
        # an entry that picks one of {endpoint_id, model} overrides the other inherited from the top level
        if "endpoint_id" in e and "model" not in e:
            merged.pop("model", None)
        if "model" in e and "endpoint_id" not in e:
            merged.pop("endpoint_id", None)
        out.append(validate_hosted_entry(merged, Path(config_path)))
    return out


# ------------------------------------------------------------------------------------------------- hosted plumbing
def hosted_payload(c: HostedEvalConfig) -> dict[str, Any]:
    return c.payload()


def create_hosted(client: APIClient, c: HostedEvalConfig, environment_ids: list[str] | None = None) -> dict[str, Any]:
    body = hosted_payload(c)
    if environment_ids is not None:
        body["environment_ids"] = environment_ids
    if client.config.team_id:
        body["team_id"] = client.config.team_id
    created = client.post("/hosted-evaluations", json=body)
    if not created.get("evaluation_id") and not created.get("evaluation_ids"):
        raise APIError(f"Failed to get evaluation ID from response: {created}")
    return created


def resolve_hosted_environment(client: APIClient, environment: str, env_dir_path: str | None, env_path: str | None) -> tuple[str, str]:
    """→ (platform slug, environment id).  Hosted runs need a published environment."""
    r = resolve_environment_reference(environment, env_dir_path or DEFAULT_ENV_DIR_PATH)
    if r.recommend_push:
        console.print("[red]Error:[/red] hosted evaluations require an environment that is published to the platform")
        console.print(f"[yellow]Publish the latest local changes for {r.platform_slug} first.[/yellow]" if r.platform_slug
                      else "[yellow]Publish the environment with `prime env push` first.[/yellow]")  # fmt: skip
        raise typer.Exit(1)
    slug = r.platform_slug or r.upstream_slug
    if slug is None and env_path:
        md = find_environment_metadata(env_name=r.env_name, env_path=Path(env_path), module_name=r.env_name.replace("-", "_")) or {}
        if md.get("owner") and md.get("name"):
            slug = f"{md['owner']}/{md['name']}"
    if slug is None:
        console.print("[red]Error:[/red] hosted evaluations require an upstream environment on the platform\n"
                      "[yellow]Use an environment slug or publish the local environment with `prime env push`.[/yellow]")  # fmt: skip
        raise typer.Exit(1)
    parts = split_owner_and_name(slug)
    if parts is None:
        raise fail(f"invalid environment slug: {slug}")
    resp = client.get(f"/environmentshub/{parts[0]}/{parts[1]}/@latest")
    env_id = (resp.get("data", resp) or {}).get("id")
    if not env_id:
        raise fail(f"could not resolve environment id for {slug}")
    console.print(f"[dim]Using hosted environment {slug}[/dim]")
    return slug, str(env_id)


def parse_status(data: dict[str, Any]) -> tuple[str, EvalStatus | None]:
    raw = str(data.get("status") or "UNKNOWN").upper()
    try:
        return raw, EvalStatus(raw)
    except ValueError:
        return raw, None


def print_eval_status(data: dict[str, Any]) -> None:
    raw, st = parse_status(data)
    console.print(f"[bold]Status:[/bold] [{st.color if st else 'white'}]{raw}[/]")
    if data.get("total_samples") is not None:
        console.print(f"[bold]Samples:[/bold] {data['total_samples']}")
    for k, v in (data.get("metrics") or {}).items():
        console.print(f"  {k}: {v}")
    if data.get("error_message"):
        console.print(f"[red]Error:[/red] {data['error_message']}")
    # where to look at it: the URL the server names (self-hosted frontends differ), else the one derived from the id
    where = data.get("viewer_url") or (get_eval_viewer_url(data["evaluation_id"]) if data.get("evaluation_id") else None)
    if where:
        console.print(f"[dim]View: {where}[/dim]")


def _fetch_eval_status(client: APIClient, eval_id: str) -> dict[str, Any]:
    return client.get(f"/evaluations/{eval_id}")


def _fetch_logs(client: APIClient, eval_id: str) -> str:
    return client.get(f"/hosted-evaluations/{eval_id}/logs").get("logs") or ""


def follow_logs(client: APIClient, eval_id: str, poll_interval: float, sleep=None) -> None:
    """Poll status + logs (diffing the sliding tail window) until the evaluation reaches a terminal status; every
    ``HOSTED_LOGS_STATUS_UPDATE_EVERY_POLLS`` polls without a new line, say what the status is so a quiet run does not look hung
    (reference: commands/evals.py:483-527).  ``sleep`` defaults to ``time.sleep`` looked up at call time."""
    console.print(f"[dim]Watching logs for evaluation {eval_id}... (Ctrl+C to stop)[/dim]\n")
    shown, errors, quiet = "", 0, 0
    while True:
        nap = sleep or time.sleep
        try:
            data = _fetch_eval_status(client, eval_id)
            logs = clean_logs(_fetch_logs(client, eval_id) or "")
            errors = 0
        except APIError as e:
            errors += 1
            if "429" in str(e):
                wait = RATE_LIMIT_WAIT_S if errors >= RATE_LIMIT_AFTER else RETRY_WAIT_S
                if errors >= RATE_LIMIT_AFTER:
                    console.print(f"[yellow]Rate limited. Waiting {wait}s...[/yellow]")
                nap(wait)
                continue
            if errors >= RATE_LIMIT_AFTER:
                raise
            nap(RETRY_WAIT_S)
            continue
        if logs and logs != shown:
            for ln in get_new_log_lines(shown, logs):
                console.print(ln, markup=False, highlight=False)
            shown, quiet = logs, 0
        else:
            quiet += 1
        raw, st = parse_status(data)
        if st is not None and st.is_terminal:
            console.print()
            print_eval_status({"evaluation_id": eval_id, **data})
            if st is not EvalStatus.COMPLETED:
                raise typer.Exit(1)
            return
        if quiet and quiet % HOSTED_LOGS_STATUS_UPDATE_EVERY_POLLS == 0:
            console.print(f"[dim]Evaluation status: {raw} (waiting for logs...)[/dim]")
        nap(poll_interval)


# ------------------------------------------------------------------------------------------------- list / get / samples
def _label(e: dict[str, Any]) -> str:
    return "HOSTED" if (e.get("is_hosted") or e.get("eval_type") == "hosted") else "LOCAL"


def _clip(s: Any, n: int = 30) -> str:
    s = "" if s is None else str(s)
    return s if len(s) <= n else s[: n - 1] + "…"


@app.command("list", epilog=list_json_help("evaluations", {"evaluation_id": "str", "name": "str", "model_name": "str", "status": "str", "total_samples": "int"}))
@handle_errors
def list_evals(output: str = typer.Option("table", "--output", "-o", help="table|json"), num: int = typer.Option(20, "--num", "-n", help="Items per page"),
               page: int = typer.Option(1, "--page", "-p"), env: Optional[str] = typer.Option(None, "--env", "--env-name", "-e", help="Filter by environment (e.g. 'gsm8k' or 'owner/gsm8k')")) -> None:  # fmt: skip
    """List evaluations of the active account."""
    if page < 1 or num < 1:
        raise fail("--page and --num must be >= 1")
    data = EvalsClient(_client()).list_evaluations(env_name=env, skip=(page - 1) * num, limit=num, team_id=Config().team_id)
    evs = data.get("evaluations", [])
    emit(output, data, f"Evaluations (Total: {data.get('total', len(evs))})",
         [("ID", "cyan"), ("Name", "blue"), ("Model", "green"), "Type", "Status", "Samples", ("Created", "magenta")],
         [[e.get("evaluation_id") or e.get("id"), _clip(e.get("name")), _clip(e.get("model_name")), _label(e), e.get("status") or "", e.get("total_samples") or "",
           format_time_ago(e.get("created_at"))] for e in evs])  # fmt: skip


@app.command("get", no_args_is_help=True)
@handle_errors
def get_eval(eval_id: str = typer.Argument(...), output: str = typer.Option("json", "--output", "-o", help="json|pretty")) -> None:
    """Fetch one evaluation."""
    data = EvalsClient(api()).get_evaluation(eval_id)
    output_data_as_json(data, console) if output == "json" else print_eval_status(data)


@app.command("samples", epilog=json_output_help({"samples": [{"example_id": "int?", "reward": "float?"}], "total": "int?"}))
@handle_errors
def get_samples(eval_id: str = typer.Argument(...), page: int = typer.Option(1, "--page", "-p"), num: int = typer.Option(100, "--num", "-n"),
                output: str = typer.Option("json", "--output", "-o", help="json|pretty")) -> None:  # fmt: skip
    """Fetch samples of an evaluation."""
    data = EvalsClient(api()).get_samples(eval_id, page=page, limit=num)
    if output == "json":
        return output_data_as_json(data, console)
    for s in data.get("samples", []):
        console.print(f"[cyan]#{s.get('example_id')}[/cyan] reward={s.get('reward')} {_clip(s.get('answer'), 60)}")


# ------------------------------------------------------------------------------------------------- push
def has_eval_files(d: Path) -> bool:
    return (d / "metadata.json").exists() and (d / "results.jsonl").exists()


def validate_eval_path(path_str: str) -> Path:
    """→ the run directory to push. Being pointed at ``metadata.json`` / ``results.jsonl`` INSIDE a complete run directory is
    forgiven (it means the directory); everything else gets an error that names what is missing."""
    p = Path(path_str)
    pair = "metadata.json and results.jsonl"
    if p.is_file():
        if p.name not in ("metadata.json", "results.jsonl"):
            raise ValueError(f"Expected a directory path containing {pair}, but got file: {p}")
        if has_eval_files(p.parent):
            return p.parent
        raise ValueError(f"Directory '{p.parent}' must contain both {pair}")
    if p.is_dir():
        missing = [n for n in ("metadata.json", "results.jsonl") if not (p / n).exists()]
        if len(missing) == 2:
            raise ValueError(f"Directory '{p}' is missing both {pair}")
        if missing:
            raise ValueError(f"Directory '{p}' is missing {missing[0]}")
        return p
    raise FileNotFoundError(f"Path not found: {p}")


def load_eval_directory(d: Path) -> dict[str, Any]:
    md = json.loads((d / "metadata.json").read_text())
    env = md.get("env_id") or md.get("env")
    if not env or "model" not in md:
        raise ValueError(f"Missing required 'env_id' or 'model' field in {d / 'metadata.json'}")
    results = load_results_jsonl(d / "results.jsonl")
    for s in results:
        if "id" in s and "example_id" not in s:
            s["example_id"] = s["id"]
    metrics = {m.group(1): v for k, v in md.items() if (m := re.match(r"avg_(.+)$", k))}
    return {"eval_name": f"{env}-{md['model']}", "model_name": md["model"], "env": env, "metrics": metrics,
            "metadata": {k: v for k, v in md.items() if not k.startswith("avg_")}, "results": results}  # fmt: skip


def discover_eval_outputs(root: Path = Path("outputs/evals")) -> list[Path]:
    if not root.exists():
        return []
    return sorted(r for e in root.iterdir() if e.is_dir() for r in e.iterdir() if r.is_dir() and has_eval_files(r))


def push_single_eval(path_str: str, env_slug: str | None, run_id: str | None, eval_id: str | None, is_public: bool = False,
                     evals: EvalsClient | None = None) -> str:  # fmt: skip
    d = validate_eval_path(path_str)
    data = load_eval_directory(d)
    console.print(f"[blue]✓ Loaded eval data:[/blue] {d}")
    if not env_slug and not run_id and not eval_id:
        env_slug = data["env"]
    environments = [{"slug": env_slug} if "/" in env_slug else {"name": env_slug}] if (env_slug and not run_id and not eval_id) else None
    evals = evals or EvalsClient(api())
    md = data["metadata"]
    common = dict(model_name=data["model_name"], framework=md.get("framework", "verifiers"), task_type=md.get("task_type"), metadata=md,
                  metrics=data["metrics"], tags=[])  # fmt: skip
    if eval_id:
        evals.get_evaluation(eval_id)
        evals.update_evaluation(eval_id, name=data["eval_name"], **common)
        console.print(f"[green]✓ Updated evaluation:[/green] {eval_id}")
    else:
        eval_id = evals.create_evaluation(name=data["eval_name"], environments=environments, run_id=run_id, is_public=is_public, **common).get("evaluation_id")
        if not eval_id:
            raise ValueError("Failed to get evaluation ID from response")
        console.print(f"[green]✓ Created evaluation:[/green] {eval_id}")
    if data["results"]:
        console.print(f"[blue]Pushing {len(data['results'])} samples...[/blue]")
        evals.push_samples(eval_id, data["results"])
    evals.finalize_evaluation(eval_id, metrics=data["metrics"])
    url = get_eval_viewer_url(eval_id)
    console.print(f"[green]✓ Success[/green]  [blue]Evaluation ID:[/blue] {eval_id}\n[dim]View:[/dim] [link={url}]{url}[/link]")
    return eval_id


@app.command("push", epilog=json_output_help({"evaluation_id": "str"}, "Auto-discovery batch push: {results: [{path, status, eval_id?, error?}]}"))
def push_eval(
    config_path: Optional[str] = typer.Argument(None, help="Run directory with metadata.json + results.jsonl (auto-discovers outputs/evals/ when omitted)"),
    env_id: Optional[str] = typer.Option(None, "--env", "--env-id", "-e", help="Environment slug (owner/name) or name"),
    run_id: Optional[str] = typer.Option(None, "--run-id", "-r", help="Attach to a training run instead of an environment"),
    eval_id: Optional[str] = typer.Option(None, "--eval", "--eval-id", help="Push to this existing evaluation id"),
    output: str = typer.Option("pretty", "--output", "-o", help="json|pretty"),
    is_public: bool = typer.Option(False, "--public", help="Make the evaluation public"),
) -> None:
    """Upload local evaluation results."""
    if eval_id and is_public:  # visibility is a property given at creation; an existing evaluation keeps the one it has
        raise fail("The --public flag cannot be used with --eval-id. Visibility can only be set when creating a new evaluation.")
    try:
        if config_path:
            eid = push_single_eval(config_path, env_id, run_id, eval_id, is_public)
            if output == "json":
                output_data_as_json({"evaluation_id": eid}, console)
            return
        found = discover_eval_outputs()
        if not found:
            raise fail("No evaluation outputs found under outputs/evals/. Pass a directory explicitly.")
        results = []
        for d in found:
            try:
                results.append({"path": str(d), "status": "success", "eval_id": push_single_eval(str(d), env_id, run_id, None, is_public)})
            except Exception as e:
                console.print(f"[red]Failed {d}:[/red] {e}")
                results.append({"path": str(d), "status": "failed", "error": str(e)})
        if output == "json":
            output_data_as_json({"results": results}, console)
        if any(r["status"] == "failed" for r in results):
            raise typer.Exit(1)
    except (ValueError, FileNotFoundError, APIError) as e:
        raise fail(str(e))


@app.command("tui")
def tui_cmd(env_dir: Optional[str] = typer.Option(None, "--env-dir", "-e", help="Environments directory"),
            outputs_dir: Optional[str] = typer.Option(None, "--outputs-dir", "-o", help="Outputs directory")) -> None:  # fmt: skip
    """Browse local evaluation outputs in the verifiers TUI."""
    run_eval_tui(env_dir, outputs_dir)


@app.command("logs", no_args_is_help=True)
@handle_errors
def logs_cmd(eval_id: str = typer.Argument(...), tail: int = typer.Option(LOGS_TAIL, "--tail", "-n"), follow: bool = typer.Option(False, "--follow", "-f"),
             poll_interval: float = typer.Option(LOGS_POLL_S, "--poll-interval", help="Seconds between polls with --follow")) -> None:  # fmt: skip
    """Logs of a hosted evaluation."""
    _display_logs(eval_id, tail, follow, poll_interval)


def _display_logs(eval_id: str, tail: int, follow: bool, poll_interval: float = LOGS_POLL_S) -> None:
    """What ``eval logs`` and ``eval run --hosted --follow`` both end in (same name and signature as the reference's seam)."""
    c = _client()
    if follow:
        try:
            return follow_logs(c, eval_id, poll_interval)
        except KeyboardInterrupt:  # not a failure: the run goes on server-side
            console.print("\n[dim]Stopped watching logs. Evaluation continues running.[/dim]")
            return None
    # status first, then the whole log (the endpoint takes no tail parameter: the last `tail` lines are cut here), then the status line
    status = _fetch_eval_status(c, eval_id)
    text = clean_logs(_fetch_logs(c, eval_id) or "")
    if text:
        console.print("\n".join(text.splitlines()[-tail:]), markup=False, highlight=False)
    else:
        console.print("[yellow]No logs available.[/yellow]")
    console.print()
    print_eval_status(status)


@app.command("stop", no_args_is_help=True)
@handle_errors
def stop_cmd(eval_id: str = typer.Argument(...)) -> None:
    """Cancel a running hosted evaluation."""
    answer = api().patch(f"/hosted-evaluations/{eval_id}/cancel")
    said = answer.get("message") if isinstance(answer, dict) else None
    console.print(f"[green]✓ {said or f'Evaluation {eval_id} cancelled.'}[/green]")  # the server's own words when it has any
    console.print(f"[dim]View results:[/dim] {get_eval_viewer_url(eval_id)}")


# ------------------------------------------------------------------------------------------------- run
def group_targets(targets: list[dict[str, Any]]) -> list[list[dict[str, Any]]]:
    """Environments with identical settings go into one hosted request (order preserving)."""
    keys = ("model", "num_examples", "rollouts_per_example", "env_args", "timeout_minutes", "allow_sandbox_access", "allow_instances_access",
            "sampling_args", "api_base_url", "api_key_var", "eval_name")  # fmt: skip
    groups: dict[tuple, list[dict[str, Any]]] = {}
    for t in targets:
        groups.setdefault(tuple(freeze(t.get(k)) for k in keys), []).append(t)
    return list(groups.values())


@app.command("run", no_args_is_help=True, context_settings={"allow_extra_args": True, "ignore_unknown_options": True, "help_option_names": []})
def run_eval_cmd(
    ctx: typer.Context,
    environment: Optional[str] = typer.Argument(None, help="Environment name/slug or TOML config"),
    skip_upload: bool = typer.Option(False, "--skip-upload", help="Do not upload results"),
    env_path: Optional[str] = typer.Option(None, "--env-path", help="Where to look for upstream environment metadata"),
    hosted: bool = typer.Option(False, "--hosted", help="Run on the platform instead of locally"),
    poll_interval: float = typer.Option(RUN_POLL_S, "--poll-interval"),
    follow: bool = typer.Option(False, "--follow", help="Follow hosted logs until completion"),
    timeout_minutes: Optional[int] = typer.Option(None, "--timeout-minutes"),
    allow_sandbox_access: bool = typer.Option(False, "--allow-sandbox-access"),
    allow_instances_access: bool = typer.Option(False, "--allow-instances-access"),
    custom_secrets: Optional[str] = typer.Option(None, "--custom-secrets", help="JSON object of sandbox secrets"),
    sampling_args: Optional[str] = typer.Option(None, "--sampling-args", help="JSON object of sampling args"),
    eval_name: Optional[str] = typer.Option(None, "--eval-name"),
) -> None:
    """Run an evaluation (locally through verifiers, or --hosted on the platform)."""
    extra = list(ctx.args)
    if is_help_request(environment or "", extra):
        print_eval_run_help()
        raise typer.Exit(0)
    if environment is None:
        console.print(f"[red]Error:[/red] Missing argument 'ENVIRONMENT'.\n[dim]Example: {EXAMPLE}[/dim]")
        raise typer.Exit(2)
    if environment.startswith("-"):
        console.print(f"[red]Error:[/red] Environment/config must be the first argument.\n[dim]Example: {EXAMPLE}[/dim]")
        raise typer.Exit(2)
    if not hosted:
        used = [flag for flag, on in {"--follow": follow, "--poll-interval": ctx.get_parameter_source("poll_interval") == ParameterSource.COMMANDLINE,
                                      "--timeout-minutes": timeout_minutes is not None, "--allow-sandbox-access": allow_sandbox_access,
                                      "--allow-instances-access": allow_instances_access, "--custom-secrets": custom_secrets is not None,
                                      "--eval-name": eval_name is not None}.items() if on]  # fmt: skip
        if used:
            raise fail("hosted-only options require `--hosted`: " + ", ".join(used))
        local_args = extra + (["--sampling-args", sampling_args] if sampling_args is not None else [])
        return run_eval_passthrough(environment, local_args, skip_upload=skip_upload, env_path=env_path)

    # ---- hosted
    env_dir = parse_value_option(extra, "--env-dir-path", "-p")
    if is_config_target(environment):
        base_targets = load_hosted_eval_configs(environment)
    else:
        base_targets = [{"env_id": environment, "env_dir_path": env_dir, "model": DEFAULT_MODEL, "num_examples": DEFAULT_NUM_EXAMPLES,
                         "rollouts_per_example": DEFAULT_ROLLOUTS}]  # fmt: skip
    cli = {"model": parse_value_option(extra, "--model", "-m"), "num_examples": parse_value_option(extra, "--num-examples", "-n"),
           "rollouts_per_example": parse_value_option(extra, "--rollouts-per-example", "-r"), "api_base_url": parse_value_option(extra, "--api-base-url", "-b"),
           "api_key_var": parse_value_option(extra, "--api-key-var", "-k")}  # fmt: skip
    raw_env_args = parse_value_option(extra, "--env-args", "")
    cli_env_args = parse_string_map_option(raw_env_args, "--env-args") if raw_env_args is not None else None
    secrets = parse_string_map_option(custom_secrets, "--custom-secrets")
    sampling = parse_json_object_option(sampling_args, "--sampling-args")
    targets = []
    for t in base_targets:
        try:
            n = int(cli["num_examples"]) if cli["num_examples"] is not None else int(t.get("num_examples", DEFAULT_NUM_EXAMPLES))
            r = int(cli["rollouts_per_example"]) if cli["rollouts_per_example"] is not None else int(t.get("rollouts_per_example", DEFAULT_ROLLOUTS))
        except ValueError:
            raise fail("--num-examples and --rollouts-per-example must be integers")
        if n < -1 or r < 1:
            raise fail("--num-examples must be >= -1 and --rollouts-per-example must be >= 1")
        targets.append({"env_id": t["env_id"], "env_dir_path": t.get("env_dir_path") or env_dir, "model": cli["model"] or t["model"], "num_examples": n,
                        "rollouts_per_example": r, "env_args": cli_env_args if raw_env_args is not None else t.get("env_args"),
                        "timeout_minutes": timeout_minutes if timeout_minutes is not None else t.get("timeout_minutes"),
                        "allow_sandbox_access": allow_sandbox_access or t.get("allow_sandbox_access", False),
                        "allow_instances_access": allow_instances_access or t.get("allow_instances_access", False), "custom_secrets": secrets,
                        "sampling_args": sampling if sampling is not None else t.get("sampling_args"),
                        "api_base_url": cli["api_base_url"] or t.get("api_base_url"), "api_key_var": cli["api_key_var"] or t.get("api_key_var"),
                        "eval_name": eval_name or t.get("eval_name")})  # fmt: skip
    if follow and len(targets) > 1:
        raise fail("`--follow` is only supported for a single hosted evaluation")
    _shared_client.append(api())
    try:
        slugs: list[str] = []
        eval_ids: list[str] = []
        try:
            # resolve EVERY environment before anything is submitted: an unknown slug in the last [[eval]] table must not leave the earlier
            # groups running (the reference resolves in file order first, then submits group by group)
            resolved_by_target = {id(t): _resolve_hosted_environment(t["env_id"], env_dir_path=t["env_dir_path"], env_path=env_path) for t in targets}
            for group in group_targets(targets):
                resolved = [resolved_by_target[id(t)] for t in group]
                t = group[0]
                cfg = HostedEvalConfig(environment_id=resolved[0][1], inference_model=t["model"], num_examples=t["num_examples"],
                                       rollouts_per_example=t["rollouts_per_example"], env_args=t.get("env_args"), name=t.get("eval_name"),
                                       timeout_minutes=t.get("timeout_minutes"), allow_sandbox_access=t["allow_sandbox_access"],
                                       allow_instances_access=t["allow_instances_access"], custom_secrets=t.get("custom_secrets"),
                                       sampling_args=t.get("sampling_args"), api_base_url=t.get("api_base_url"), api_key_var=t.get("api_key_var"))  # fmt: skip
                created = _create_hosted_evaluations(cfg, environment_ids=[e for _, e in resolved])
                slugs += [s for s, _ in resolved]
                eval_ids += created.get("evaluation_ids") or [created["evaluation_id"]]
        except APIError as e:
            console.print(f"[red]Hosted evaluation failed:[/red] {e}")
            raise typer.Exit(1)
        console.print("[green]✓ Hosted evaluation started[/green]")
        for s, e in zip(slugs, eval_ids):
            url = get_eval_viewer_url(e)
            console.print(f"[cyan]Environment:[/cyan] {s}  [cyan]Evaluation ID:[/cyan] {e}\n  [link={url}]{url}[/link]")
        if len(eval_ids) > 1:  # one greppable line with all of them, in submission order
            console.print(f"[cyan]Evaluation IDs:[/cyan] {', '.join(eval_ids)}")
        if follow:
            console.print()
            return _display_logs(eval_ids[0], LOGS_TAIL, True, poll_interval)
        console.print(f"\n[dim]Follow progress with: prime eval logs {eval_ids[0]} -f[/dim]")
    finally:
        _shared_client.pop()


# ---- seams: the steps of a hosted run as module-level callables with the reference's names and signatures, looked up at call time, so that
# a harness (the reference's own tests; a dry-run wrapper) can replace one step.  They share the command's client while it runs.
_shared_client: list[APIClient] = []


def _client() -> APIClient:
    return _shared_client[-1] if _shared_client else api()


def _resolve_hosted_environment(environment: str, env_dir_path: str | None = None, env_path: str | None = None) -> tuple[str, str]:
    return resolve_hosted_environment(_client(), environment, env_dir_path, env_path)


def _create_hosted_evaluations(config: HostedEvalConfig, environment_ids: list[str] | None = None) -> dict[str, Any]:
    return create_hosted(_client(), config, environment_ids)



_has_eval_files = has_eval_files
_validate_eval_path = validate_eval_path
_load_eval_directory = load_eval_directory
_push_single_eval = push_single_eval
_load_hosted_eval_configs = load_hosted_eval_configs
_print_eval_status = print_eval_status
_build_hosted_evaluation_payload = hosted_payload
