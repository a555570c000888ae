"""``prime rl [run] cfg.toml`` and friends: hosted RL training
(reference: packages/prime/src/prime_cli/commands/rl.py:93-1510).

Subcommands: run (default), models, list|ls, get, stop, delete, restart, logs [-f], init, metrics, rollouts, progress,
distributions, checkpoints.  The TOML schema rejects unknown keys (``extra="forbid"``), parses ``owner/name@version``
environment ids, and serialises to the API with ``None`` omitted.
"""

from __future__ import annotations

import json
import time
import tomllib
from pathlib import Path
from typing import Any, Dict, List, Optional

import typer
from pydantic import BaseModel, ConfigDict, Field, ValidationError as PydanticValidationError, model_validator
from rich.markup import escape

from ..api.rl import RLClient, RLRun
from ..core import APIError, Config
from ..utils.display import RUN_STATUS_COLORS, colorize, output_data_as_json, validate_output_format
from ..utils.env_vars import EnvParseError, collect_env_vars
from ..utils.formatters import strip_ansi
from ..utils.hosted_eval import get_new_log_lines, tqdm_line
from ..utils.json_help import json_output_help, list_json_help
from ..utils.prompt import confirm_or_skip
from ..utils.time_utils import format_time_ago
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app, paginate_hint

app = make_app("Hosted RL training", default_cmd="run")
LEVEL_STYLES = {"DEBUG": "dim", "INFO": "cyan", "WARNING": "yellow", "WARN": "yellow", "ERROR": "red", "CRITICAL": "bold red", "SUCCESS": "green"}
_SKIP = object()
DEPRECATED_KEYS = ("trajectory_strategy", "trajectoryStrategy")
FOLLOW_POLL_S, QUEUED_WAIT_S, RATE_LIMIT_WAIT_S, RATE_LIMIT_LONG_WAIT_S = 5, 10, 10, 30


# ------------------------------------------------------------------------------------------------- config schema
class _Section(BaseModel):
    model_config = ConfigDict(extra="forbid")

    def to_api_dict(self) -> Dict[str, Any] | None:
        """Every set field, ``None`` omitted; an empty section vanishes from the payload."""
        d = {k: v for k, v in self.model_dump().items() if v is not None}
        return d or None


class _EnvRef(_Section):
    id: str
    name: str | None = None
    args: Dict[str, Any] = Field(default_factory=dict)
    version: str | None = None

    @model_validator(mode="after")
    def _split_version(self):
        if "@" in self.id:
            base, ver = self.id.rsplit("@", 1)
            self.id = base
            if self.version is None and ver:
                self.version = ver
        return self

    def to_api_dict(self) -> Dict[str, Any]:
        d = super().to_api_dict() or {}
        if not self.args:
            d.pop("args", None)
        return d


class EnvConfig(_EnvRef):
    pass


class EvalEnvConfig(_EnvRef):
    num_examples: int | None = None
    rollouts_per_example: int | None = None


class TemperatureSchedulerConfig(_Section):
    type: str = "linear"  # linear | cosine
    start_temperature: float
    end_temperature: float
    total_steps: int | None = None


class SamplingConfig(_Section):
    max_tokens: int | None = None
    temperature: float | None = None
    repetition_penalty: float | None = None
    min_tokens: int | None = None
    seed: int | None = None
    temp_scheduler: TemperatureSchedulerConfig | None = None
    extra_body: Dict[str, Any] | None = None


class EvalConfig(_Section):
    interval: int | None = None
    num_examples: int | None = None
    rollouts_per_example: int | None = None
    eval_base_model: bool | None = None
    env: List[EvalEnvConfig] = Field(default_factory=list)

    def to_api_dict(self) -> Dict[str, Any] | None:
        if not self.env:
            return None
        d: Dict[str, Any] = {"environments": [e.to_api_dict() for e in self.env]}
        d.update({k: v for k, v in self.model_dump(exclude={"env"}).items() if v is not None})
        return d


class ValConfig(_Section):
    num_examples: int | None = None
    rollouts_per_example: int | None = None
    interval: int | None = None


class BufferConfig(_Section):
    easy_threshold: float | None = None
    hard_threshold: float | None = None
    easy_fraction: float | None = None
    hard_fraction: float | None = None
    online_difficulty_filtering: bool | None = None
    env_ratios: List[float] | None = None
    skip_verification: bool | None = None
    seed: int | None = None


class WandbConfig(_Section):
    entity: str | None = None
    project: str | None = None
    name: str | None = None


class CheckpointsConfig(_Section):
    interval: int | None = None  # save every N steps
    keep_cloud: int | None = None  # -1 keeps all


class AdaptersConfig(_Section):
    interval: int | None = None  # 0 = only at run end
    keep_last: int | None = None  # -1 keeps all


class InfrastructureConfig(_Section):
    compute_size: str | None = None  # S | M | L


class RLConfig(_Section):
    name: str | None = None
    model: str
    max_steps: int = 100
    batch_size: int = 128
    rollouts_per_example: int = 8
    learning_rate: float | None = None
    lora_alpha: int | None = None
    oversampling_factor: float | None = None
    max_async_level: int | None = None
    checkpoint_id: str | None = None  # warm start
    cluster_name: str | None = None  # admin only
    env: List[EnvConfig] = Field(default_factory=list)
    sampling: SamplingConfig = Field(default_factory=SamplingConfig)
    eval: EvalConfig = Field(default_factory=EvalConfig)
    val: ValConfig = Field(default_factory=ValConfig)
    buffer: BufferConfig = Field(default_factory=BufferConfig)
    wandb: WandbConfig = Field(default_factory=WandbConfig)
    checkpoints: CheckpointsConfig = Field(default_factory=CheckpointsConfig)
    adapters: AdaptersConfig = Field(default_factory=AdaptersConfig)
    infrastructure: InfrastructureConfig = Field(default_factory=InfrastructureConfig)
    env_file: List[str] = Field(default_factory=list)  # deprecated spelling
    env_files: List[str] = Field(default_factory=list)


def format_validation_errors(errors: list[Any]) -> list[str]:
    return [f"{'.'.join(str(x) for x in e['loc'])}: {e['msg'].removeprefix('Value error, ')}" for e in errors]


def load_config(path: str) -> RLConfig:
    p = Path(path)
    if not p.exists():
        raise fail(f"Config file not found: {path}")
    try:
        data = tomllib.loads(p.read_text())
    except tomllib.TOMLDecodeError as e:
        raise fail(f"Invalid TOML in {path}: {e}")
    if any(data.pop(k, _SKIP) is not _SKIP for k in DEPRECATED_KEYS):
        console.print("[yellow]Warning:[/yellow] `trajectory_strategy` is deprecated and ignored.")
    try:
        return RLConfig.model_validate(data)
    except PydanticValidationError as e:
        console.print(f"[red]Error:[/red] Invalid config in {path}:\n")
        for msg in format_validation_errors(e.errors()):
            console.print(f"  [red]•[/red] {msg}")
        console.print()
        raise typer.Exit(1)


CONFIG_TEMPLATE = '''# Hosted RL run — edit and launch with:  prime rl run {filename}
model = "{model}"
# name = "my-run"
max_steps = 100
batch_size = 128
rollouts_per_example = 8
# learning_rate = 1e-5
# lora_alpha = 32
# oversampling_factor = 1.0
# max_async_level = 2
# checkpoint_id = "..."        # warm-start from an earlier run's checkpoint
# env_files = ["secrets.env"]   # KEY=VALUE lines, ${{VAR}} expands from your shell

[[env]]
id = "{environment}"           # owner/name or owner/name@version
# args = {{ difficulty = "hard" }}

[sampling]
max_tokens = 2048
# temperature = 1.0
# [sampling.temp_scheduler]
# type = "linear"               # or "cosine"
# start_temperature = 1.0
# end_temperature = 0.6

# [eval]
# interval = 20
# [[eval.env]]
# id = "{environment}"
# num_examples = 64

# [val]
# num_examples = 32
# interval = 10

# [buffer]
# online_difficulty_filtering = true
# easy_threshold = 1.0
# hard_threshold = 0.0

# [wandb]                       # needs WANDB_API_KEY via -e/--env-file
# project = "my-project"
# entity = "my-team"

# [checkpoints]
# interval = 50
# keep_cloud = 3                # -1 keeps everything

# [adapters]
# interval = 0                  # 0 = upload only at run end
# keep_last = 1

# [infrastructure]
# compute_size = "M"          # S, M (default), or L
'''


def generate_rl_config_template(environment: str | None = None, model: str = "PrimeIntellect/Qwen3-0.6B-Reverse-Text-SFT",
                                filename: str = "rl.toml") -> str:  # fmt: skip
    """Defaults = the reference template's quick-start pair: a 0.6 B model and the reverse-text environment (a run that starts in minutes)."""
    return CONFIG_TEMPLATE.format(environment=environment or "primeintellect/reverse-text", model=model, filename=filename)


# ------------------------------------------------------------------------------------------------- log rendering
def format_json_log_line(line: str):
    """Structured ``{timestamp, level, message, type}`` lines → ``HH:MM:SS [LEVEL] message``; ``type=progress`` is dropped
    (returns the ``_SKIP`` sentinel); non-JSON lines return None so the caller falls back to plain handling."""
    t = line.strip()
    if not (t.startswith("{") and t.endswith("}")):
        return None
    try:
        e = json.loads(t)
    except json.JSONDecodeError:
        return None
    if not isinstance(e, dict) or "timestamp" not in e or "level" not in e:
        return None
    if e.get("type") == "progress":
        return _SKIP
    ts = str(e.get("timestamp", ""))
    clock = ts.split("T")[1][:8] if "T" in ts else ts[:8]
    level = str(e.get("level", "INFO")).upper()
    style = LEVEL_STYLES.get(level, "")
    tag = f"[{style}]\\[{level}][/{style}]" if style else f"\\[{level}]"
    return f"[dim]{clock}[/dim] {tag} {escape(str(e.get('message', '')))}"


def clean_logs(text: str) -> list[str]:
    out: list[str] = []
    for line in text.splitlines():
        if not line.strip():
            continue
        f = format_json_log_line(line)
        if f is _SKIP:
            continue
        if f is not None:
            out.append(f)
            continue
        plain = strip_ansi(line)
        if tqdm_line(plain) == "":
            continue  # tqdm refreshes: keep only the completed bar
        if plain.strip():
            out.append(plain)
    return out


def is_queued_404(e: Exception) -> bool:
    s = str(e).lower()
    return "404" in s and ("queued" in s or "pending" in s)


# ------------------------------------------------------------------------------------------------- run
def run_row(r: RLRun) -> dict[str, Any]:
    return {"id": r.id, "name": r.name, "status": r.status, "model": r.base_model, "environments": [e.get("id") or e.get("name") for e in r.environments],
            "max_steps": r.max_steps, "batch_size": r.batch_size, "rollouts_per_example": r.rollouts_per_example, "seq_len": r.seq_len,
            "created_at": r.created_at.isoformat(), "started_at": r.started_at.isoformat() if r.started_at else None,
            "completed_at": r.completed_at.isoformat() if r.completed_at else None, "error_message": r.error_message,
            "runs_ahead": r.runs_ahead, "team_id": r.team_id}  # fmt: skip


def run_json(r: RLRun) -> dict[str, Any]:
    """What ``-o json`` prints for a run: every field of the record under its own name (``base_model``, ``environments`` as the list
    of dicts, timestamps in ``str(datetime)`` form — the shape scripts parse) plus two convenience keys."""
    return {**r.model_dump(), "model": r.base_model, "dashboard_url": dashboard_url(r.id)}


def dashboard_url(run_id: str) -> str:
    return f"{Config(writable=False).frontend_url}/dashboard/training/{run_id}"


def check_environment_actions(client: RLClient, envs: list[EnvConfig], skip: bool) -> None:
    """Refuse to launch when an environment's latest hub build/test action failed (unless told otherwise)."""
    for e in envs:
        parts = e.id.split("/", 1)
        if len(parts) != 2:
            continue
        try:
            st = client.get_environment_status(parts[0], parts[1])
        except APIError:
            continue
        action = (st.get("latest_action") or st.get("action") or {}) if isinstance(st, dict) else {}
        if str(action.get("status", "")).upper() == "FAILED":
            msg = f"Environment {e.id}: latest action FAILED ({action.get('error') or action.get('message') or 'no details'})."
            if skip:
                console.print(f"[yellow]Warning:[/yellow] {msg} Continuing (--skip-action-check).")
            else:
                console.print(f"[red]Error:[/red] {msg}\nFix it ('prime env action logs'), or pass --skip-action-check.")
                raise typer.Exit(1)


def print_config_summary(cfg: RLConfig, secrets: dict[str, str]) -> None:
    console.print("[bold]Configuration:[/bold]")
    console.print(f"  Model: {cfg.model}")
    console.print("  Environments: " + ", ".join(e.id + (f"@{e.version}" if e.version else "") for e in cfg.env))
    console.print(f"  Max Steps: {cfg.max_steps}\n  Batch Size: {cfg.batch_size}\n  Rollouts per Example: {cfg.rollouts_per_example}")
    for label, v in (("Learning Rate", cfg.learning_rate), ("LoRA Alpha", cfg.lora_alpha), ("Oversampling Factor", cfg.oversampling_factor),
                     ("Max Async Level", cfg.max_async_level), ("Max Tokens", cfg.sampling.max_tokens), ("Temperature", cfg.sampling.temperature),
                     ("Warm start from", cfg.checkpoint_id), ("Compute Size", cfg.infrastructure.compute_size),
                     ("W&B", f"{cfg.wandb.entity or '-'}/{cfg.wandb.project}" if (cfg.wandb.project or cfg.wandb.entity) else None),
                     ("Eval", f"{len(cfg.eval.env)} env(s) every {cfg.eval.interval or '?'} steps" if cfg.eval.env else None),
                     ("Secrets", ", ".join(sorted(secrets)) if secrets else None)):  # fmt: skip
        if v is not None:
            console.print(f"  {label}: {v}")
    console.print()


@app.command("run", epilog=json_output_help({"run": {"id": "str", "status": "str", "…": "every field of the run"}, "id": "str", "status": "str",
                                                 "runs_ahead": "int|null", "dashboard_url": "str"}))
@handle_errors
def create_run(
    config_path: str = typer.Argument(..., help="TOML config (see 'prime rl init')"),
    env: Optional[List[str]] = typer.Option(None, "-e", "--env-var", help="Secret for the training container: KEY=VALUE, KEY (from $KEY) or a .env path"),
    env_file: Optional[List[str]] = typer.Option(None, "--env-file", help=".env file with secrets (${VAR} expands from your shell)"),
    output: str = OUTPUT_OPT,
    skip_action_check: bool = typer.Option(False, "--skip-action-check", help="Launch even if an environment's hub action failed"),
) -> None:
    """Start an RL training run from a config file."""
    validate_output_format(output, console)
    console.print(f"[dim]Loading config from {config_path}[/dim]\n")
    cfg = load_config(config_path)
    base = Path(config_path).parent
    files = [str(base / f) for f in cfg.env_file + cfg.env_files] + list(env_file or [])  # config files first, CLI files override
    try:
        secrets = collect_env_vars(env_args=env, env_files=files or None, on_warning=lambda m: console.print(f"[yellow]Warning:[/yellow] {m}"))
    except EnvParseError as e:
        raise fail(str(e))
    if (cfg.wandb.entity or cfg.wandb.project) and "WANDB_API_KEY" not in secrets:
        console.print("[red]Configuration Error:[/red]\n  WANDB_API_KEY is required when W&B monitoring is configured.\n")
        console.print("Provide it via:\n  prime rl run cfg.toml -e WANDB_API_KEY          (from your shell)\n  prime rl run cfg.toml --env-file secrets.env")
        raise typer.Exit(1)
    if not cfg.env:
        raise fail("Config needs at least one [[env]] entry")
    client = RLClient(api())
    team = Config(writable=False).team_id
    print_config_summary(cfg, secrets)
    check_environment_actions(client, cfg.env, skip_action_check)
    s = cfg.sampling
    run = client.create_run(
        cfg.model, [e.to_api_dict() for e in cfg.env], rollouts_per_example=cfg.rollouts_per_example, max_steps=cfg.max_steps,
        batch_size=cfg.batch_size, name=cfg.name, secrets=secrets or None, team_id=team, wandb_entity=cfg.wandb.entity,
        wandb_project=cfg.wandb.project, wandb_run_name=cfg.wandb.name, max_tokens=s.max_tokens, temperature=s.temperature,
        repetition_penalty=s.repetition_penalty, min_tokens=s.min_tokens, seed=s.seed,
        temp_scheduler=s.temp_scheduler.to_api_dict() if s.temp_scheduler else None, extra_body=s.extra_body,
        eval_config=cfg.eval.to_api_dict(), val_config=cfg.val.to_api_dict(), buffer_config=cfg.buffer.to_api_dict(),
        learning_rate=cfg.learning_rate, lora_alpha=cfg.lora_alpha, oversampling_factor=cfg.oversampling_factor,
        max_async_level=cfg.max_async_level, checkpoints_config=cfg.checkpoints.to_api_dict(), adapters_config=cfg.adapters.to_api_dict(),
        checkpoint_id=cfg.checkpoint_id, cluster_name=cfg.cluster_name, infrastructure_config=cfg.infrastructure.to_api_dict(),
    )  # fmt: skip
    url = dashboard_url(run.id)
    if output == "json":
        # {"run": <every field of the created run>} as scripts know it (timestamps in str(datetime) form), plus the summary keys
        return output_data_as_json({"run": run.model_dump(), **run_json(run)}, console)
    if run.status == "QUEUED":
        ahead = f" (~{run.runs_ahead} run(s) ahead)" if run.runs_ahead is not None else ""
        console.print(f"[yellow]Run {run.id} is QUEUED{ahead}[/yellow]")
    else:
        console.print(f"[green]✓ Run created:[/green] {run.id}")
    console.print(f"\n[bold]Dashboard:[/bold] [link={url}]{url}[/link]\n[dim]Follow logs with: prime rl logs {run.id} -f[/dim]")


# ------------------------------------------------------------------------------------------------- management
@app.command("models", epilog=list_json_help("models", {"name": "str", "at_capacity": "bool"}))
@handle_errors
def list_models(output: str = OUTPUT_OPT) -> None:
    """Models available for RL training."""
    models = RLClient(api()).list_models(team_id=Config(writable=False).team_id)
    emit(output, {"models": [m.model_dump() for m in models]}, "RL Models", [("Model", "cyan"), "Availability"],
         [[m.name, "[yellow]at capacity[/yellow]" if m.at_capacity else "[green]available[/green]"] for m in models])  # fmt: skip


def _list_runs(output: str, team: str | None, num: int, page: int) -> None:
    """The runs endpoint is not paginated: sort newest first and page client-side, like the reference
    (packages/prime/src/prime_cli/commands/rl.py:945-978)."""
    if num < 1 or page < 1:
        raise fail("--num and --page must be at least 1")
    runs = sorted(RLClient(api()).list_runs(team_id=team or Config(writable=False).team_id), key=lambda r: r.created_at, reverse=True)
    total = len(runs)
    shown = runs[(page - 1) * num : page * num]
    rows = [run_row(r) for r in shown]
    if output != "json" and not rows:
        console.print("[yellow]No more results.[/yellow]" if page > 1 else "[yellow]No RL training runs found.[/yellow]")
        return
    emit(output, {"runs": [run_json(r) for r in shown], "total": total, "total_count": total, "page": page, "per_page": num}, f"RL Runs (Total: {total})",
         [("ID", "cyan"), ("Name", "blue"), "Status", ("Model", "green"), "Environments", "Steps", ("Created", "magenta")],
         [[r["id"], r["name"] or "", colorize(r["status"], RUN_STATUS_COLORS), r["model"], ", ".join(map(str, r["environments"])), r["max_steps"],
           format_time_ago(r["created_at"])] for r in rows],
         paginate_hint(total, (page - 1) * num, num, "runs"))  # fmt: skip


_RUN_LIST_HELP = list_json_help("runs", {"id": "str", "name": "str|null", "status": "str", "model": "str", "created_at": "str"},
                                {"total": "int", "page": "int", "per_page": "int"})  # fmt: skip
_TEAM_OPT = typer.Option(None, "--team", "--team-id", "-t", help="Filter by team ID")
_NUM_OPT = typer.Option(20, "--num", "-n", help="Items per page")
_PAGE_OPT = typer.Option(1, "--page", "-p", help="Page number")


@app.command("list", epilog=_RUN_LIST_HELP)
@handle_errors
def list_runs(team: Optional[str] = _TEAM_OPT, num: int = _NUM_OPT, page: int = _PAGE_OPT, output: str = OUTPUT_OPT) -> None:
    """Your RL runs, newest first."""
    _list_runs(output, team, num, page)


@app.command("ls", hidden=True)
@handle_errors
def ls_runs(team: Optional[str] = _TEAM_OPT, num: int = _NUM_OPT, page: int = _PAGE_OPT, output: str = OUTPUT_OPT) -> None:
    """Alias of 'list'."""
    _list_runs(output, team, num, page)


@app.command("get", epilog=json_output_help({"id": "str", "status": "str", "error_message": "str|null"}))
@handle_errors
def get_run(run_id: str = typer.Argument(...), output: str = OUTPUT_OPT) -> None:
    """Details of one run."""
    run = RLClient(api()).get_run(run_id)
    row = run_row(run)
    emit(output, {"run": run.model_dump(), **run_json(run)}, f"RL Run {run_id}", [("Field", "cyan"), ("Value", "green")],
         [[k, colorize(v, RUN_STATUS_COLORS) if k == "status" else ("" if v is None else v)] for k, v in row.items()])  # fmt: skip


@app.command("stop")
@handle_errors
def stop_run(run_id: str = typer.Argument(...), yes: bool = typer.Option(False, "--yes", "-y", "--force", "-f", help="Skip confirmation")) -> None:
    """Stop a running run (checkpoints written so far are kept)."""
    if not confirm_or_skip(f"Stop run {run_id}?", yes):
        raise typer.Exit(0)
    r = RLClient(api()).stop_run(run_id)
    console.print(f"[green]✓ Run {run_id} stopping[/green] (status: {r.status})")


@app.command("delete")
@handle_errors
def delete_run(run_id: str = typer.Argument(...), yes: bool = typer.Option(False, "--yes", "-y", "--force", "-f", help="Skip confirmation")) -> None:
    """Delete a run and its records."""
    if not confirm_or_skip(f"Delete run {run_id}? This cannot be undone.", yes):
        raise typer.Exit(0)
    RLClient(api()).delete_run(run_id)
    console.print(f"[green]✓ Deleted run {run_id}[/green]")


@app.command("restart")
@handle_errors
def restart_run(run_id: str = typer.Argument(...), yes: bool = typer.Option(False, "--yes", "-y", "--force", "-f", help="Skip confirmation")) -> None:
    """Restart a RUNNING run from its latest checkpoint (server side)."""
    client = RLClient(api())
    # no status pre-check: the server refuses runs that are not RUNNING and says why (one request, as the reference does)
    if not confirm_or_skip(f"Restart run {run_id} from its latest checkpoint? (only RUNNING runs can be restarted)", yes):
        raise typer.Exit(0)
    r = client.restart_run(run_id)
    console.print(f"[green]✓ Restart requested[/green] (status: {r.status})")


@app.command("logs")
@handle_errors
def get_logs(run_id: str = typer.Argument(...), tail: int = typer.Option(1000, "--tail", "-n", help="Lines to fetch"),
             follow: bool = typer.Option(False, "--follow", "-f", help="Keep polling for new lines"),
             raw: bool = typer.Option(False, "--raw", "-r", help="No formatting")) -> None:  # fmt: skip
    """Show (or follow) a run's logs."""
    client = RLClient(api())
    render = (lambda t: t.splitlines()) if raw else clean_logs
    if not follow:
        lines = render(client.get_logs(run_id, tail_lines=tail))
        for ln in lines:
            console.print(ln, markup=not raw, highlight=False)
        if not lines:
            console.print("[yellow]No logs available yet.[/yellow]")
        return
    console.print(f"[dim]Watching logs for run {run_id}... (Ctrl+C to stop)[/dim]\n")
    shown: list[str] = []
    errors = 0
    while True:
        try:
            current = render(client.get_logs(run_id, tail_lines=tail))
            errors = 0
            if current != shown:
                # the server returns a sliding tail window: print only what extends past the overlap
                for ln in get_new_log_lines("\n".join(shown), "\n".join(current)) if shown else current:
                    console.print(ln, markup=not raw, highlight=False)
                shown = current
        except APIError as e:
            if is_queued_404(e):
                console.print("[yellow]Run is queued, waiting for it to start...[/yellow]")
                time.sleep(QUEUED_WAIT_S)
                continue
            errors += 1
            if "429" in str(e):
                if errors >= 3:
                    console.print("[yellow]Rate limited. Waiting 30s...[/yellow]")
                time.sleep(RATE_LIMIT_LONG_WAIT_S if errors >= 3 else RATE_LIMIT_WAIT_S)
                continue
            raise
        time.sleep(FOLLOW_POLL_S)


@app.command("init")
def init_config(path: str = typer.Argument("rl.toml", help="Where to write the template"),
                environment: Optional[str] = typer.Option(None, "--env", help="Environment id to pre-fill"),
                force: bool = typer.Option(False, "--force", "-f", help="Overwrite an existing file")) -> None:  # fmt: skip
    """Write a commented config template."""
    p = Path(path)
    if p.exists() and not force:
        raise fail(f"{path} already exists (use --force to overwrite)")
    p.write_text(generate_rl_config_template(environment, filename=p.name))
    console.print(f"[green]✓ Wrote {path}[/green]\n[dim]Edit it, then: prime rl run {path}[/dim]")


@app.command("metrics")
@handle_errors
def get_metrics(run_id: str = typer.Argument(...), min_step: Optional[int] = typer.Option(None, "--min-step", help="Minimum step (inclusive)"),
                max_step: Optional[int] = typer.Option(None, "--max-step", help="Maximum step (inclusive)"),
                limit: Optional[int] = typer.Option(None, "--limit", "-n", help="Maximum number of records")) -> None:  # fmt: skip
    """Training metrics as JSON."""
    output_data_as_json({"run_id": run_id, "metrics": RLClient(api()).get_metrics(run_id, min_step, max_step, limit)}, console)


@app.command("rollouts")
@handle_errors
def get_rollouts(run_id: str = typer.Argument(...), step: int = typer.Option(..., "--step", "-s", help="Step number"),
                 page: int = typer.Option(1, "--page", "-p", help="Page number (1-indexed)"),
                 num: int = typer.Option(100, "--num", "--limit", "-n", help="Items per page")) -> None:  # fmt: skip
    """Rollout samples of one step as JSON."""
    output_data_as_json(RLClient(api()).get_rollouts(run_id, step, page, num), console)


@app.command("progress")
@handle_errors
def get_progress(run_id: str = typer.Argument(...)) -> None:
    """Latest step and which steps have samples / distributions (JSON)."""
    output_data_as_json(RLClient(api()).get_progress(run_id), console)


@app.command("distributions")
@handle_errors
def get_distributions(run_id: str = typer.Argument(...), type: Optional[str] = typer.Option(None, "--type", "-t", help="reward | advantage"),
                      step: Optional[int] = typer.Option(None, "--step", "-s")) -> None:  # fmt: skip
    """Reward/advantage histogram as JSON."""
    output_data_as_json(RLClient(api()).get_distributions(run_id, type, step), console)


@app.command("checkpoints", epilog=list_json_help("checkpoints", {"id": "str", "step": "int", "status": "str", "size_bytes": "int|null", "storage_url": "str"}))
@handle_errors
def list_checkpoints(run_id: str = typer.Argument(...),
                     status: Optional[str] = typer.Option(None, "--status", "-s", help="Filter by status (READY, PENDING, UPLOADING, FAILED)"),
                     output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Checkpoints written by a run (pass an id as `checkpoint_id` to warm-start another run)."""
    from ..utils.formatters import format_size

    cps = RLClient(api()).list_checkpoints(run_id, status_filter=status)
    emit(output, {"checkpoints": [c.model_dump(mode="json") for c in cps], "total_count": len(cps)}, f"Checkpoints of {run_id}",
         [("ID", "cyan"), "Step", "Status", "Size", ("Created", "magenta")],
         [[c.id, c.step, c.status, format_size(c.size_bytes), format_time_ago(c.created_at)] for c in sorted(cps, key=lambda c: c.step)])  # fmt: skip
