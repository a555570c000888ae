"""Bridge between ``prime eval|gepa|env|lab`` and the external ``verifiers`` toolkit (always run as a subprocess).

Behavioural parity with reference packages/prime/src/prime_cli/verifiers_bridge.py:
  * help rewriting so ``python -m verifiers…`` usage reads as ``prime …``                         (:67-113)
  * environment reference resolution: ``owner/name[@ver]`` → remote; a local checkout under the env dir → local
    (with hub sync status from content hash / version); otherwise search personal → team → ``primeintellect`` (:508-622)
  * local content hash over pyproject.toml, top-level *.py, README.md and non-ignored subtrees              (:304-346)
  * inference defaults (-m/-b/-k), model validation, billing pre-flight                                     (:757-827)
  * run, then upload results unless skipped / config-driven / unpublished                                   (:871-1009)
"""

from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys
import tomllib
import uuid
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path

import httpx
import typer

from .api.inference import InferenceAPIError, InferenceClient, InferencePaymentRequiredError
from .core import APIClient, APIError, Config
from .utils.env_metadata import find_environment_metadata, get_environment_metadata
from .utils.eval_push import push_eval_results_to_hub
from .utils.plain import get_console
from .verifiers_plugin import PrimeVerifiersPlugin, load_verifiers_prime_plugin

console = get_console()
DEFAULT_MODEL = "openai/gpt-4.1-mini"
DEFAULT_ENV_DIR_PATH = "./environments"
PRIME_SLUG = "primeintellect"
INTERNAL_ENV_DISPLAY_HEADER = "X-Prime-Eval-Env-Display"
EVAL_PREFLIGHT_TIMEOUT = httpx.Timeout(connect=10.0, read=300.0, write=60.0, pool=60.0)
MODULE_TO_PRIME_COMMAND = {
    "verifiers.cli.commands.eval": "prime eval run",
    "verifiers.cli.commands.gepa": "prime gepa run",
    "verifiers.cli.commands.init": "prime env init",
    "verifiers.cli.commands.install": "prime env install",
    "verifiers.cli.commands.build": "prime env build",
    "verifiers.cli.commands.setup": "prime lab setup",
    "verifiers.cli.tui": "prime eval tui",
}
SKIP_DIRS = {"dist", "__pycache__", "build", "outputs"}
EXTRA_EVAL_HELP = [
    "  --skip-upload               Skip uploading evaluation results to the platform.",
    "  --env-path PATH             Explicit path for upstream environment metadata.",
    "  --hosted                    Run the evaluation on the platform instead of locally.",
    "  stop EVAL_ID                Cancel a running hosted evaluation.",
    "  --poll-interval FLOAT       Polling interval in seconds for hosted evaluations.",
    "  --follow                    Follow hosted evaluation logs until completion.",
    "  --timeout-minutes INTEGER   Timeout in minutes for hosted evaluations.",
    "  --allow-sandbox-access      Allow sandbox read/write access for hosted evaluations.",
    "  --allow-instances-access    Allow instance creation and management for hosted evaluations.",
    "  --custom-secrets JSON       Custom sandbox secrets for hosted evaluations.",
    "  --eval-name TEXT            Custom name for the hosted evaluation.",
]


@dataclass(frozen=True)
class ResolvedEnvironment:
    original: str
    env_name: str
    install_mode: str  # "remote" | "local" | "none"
    install_slug: str | None = None
    upstream_slug: str | None = None
    env_display_id: str | None = None
    platform_slug: str | None = None
    platform_url: str | None = None
    recommend_push: bool = False
    push_reason: str | None = None  # "ahead" | "local_only"
    local_env_path: Path | None = None


# ------------------------------------------------------------------------------------------------ argv helpers
def is_help_request(primary_arg: str, passthrough_args: list[str]) -> bool:
    return primary_arg in ("-h", "--help") or any(a in ("-h", "--help") for a in passthrough_args)


def parse_value_option(args: list[str], long_flag: str, short_flag: str) -> str | None:
    """Find ``--flag V`` / ``--flag=V`` / ``-f V`` / ``-fV`` in a pass-through argv."""
    for i, a in enumerate(args):
        if a in (long_flag, short_flag):
            return args[i + 1] if i + 1 < len(args) else None
        if a.startswith(long_flag + "="):
            return a.split("=", 1)[1]
        if short_flag and a.startswith(short_flag) and len(a) > len(short_flag) and not a.startswith("--"):
            return a[len(short_flag) :]
    return None


def has_flag(args: list[str], long_flag: str, short_flag: str) -> bool:
    return any(a in (long_flag, short_flag) or a.startswith(long_flag + "=") for a in args)


def is_config_target(raw: str) -> bool:
    return raw.endswith(".toml")


def split_version(ref: str) -> tuple[str, str | None]:
    base, sep, ver = ref.rpartition("@")
    return (base, ver) if sep else (ref, None)


def is_slug_reference(base_ref: str) -> bool:
    return "/" in base_ref and not base_ref.startswith(("./", "../", "/")) and not base_ref.endswith(".toml")


def split_owner_and_name(slug: str) -> tuple[str, str] | None:
    owner, sep, name = slug.partition("/")
    return (owner, name) if sep and owner and name else None


def environment_url_from_slug(slug: str) -> str | None:
    if split_owner_and_name(slug) is None:
        return None
    try:
        base = Config(writable=False).frontend_url
    except Exception:
        base = "https://app.primeintellect.ai"
    return f"{base}/dashboard/environments/{slug}"


def build_job_id(env_name: str, model: str) -> str:
    clean = lambda s: s.replace("/", "_").replace("-", "_")  # noqa: E731
    return f"{clean(env_name)}_{clean(model)}_{datetime.now():%Y%m%d_%H%M%S}_{uuid.uuid4().hex[:8]}"


# ------------------------------------------------------------------------------------------------ help rewriting
def sanitize_help_text(help_text: str, module_name: str, prime_command: str) -> str:
    lines = help_text.splitlines()
    for i, line in enumerate(lines):
        if line.lower().startswith("usage:"):
            rest = line.split(":", 1)[1].strip()
            m = re.match(rf"(?:python(?:\d+(?:\.\d+)?)?\s+-m\s+)?{re.escape(module_name)}(?:\s+|$)", rest)
            rest = rest[m.end() :].lstrip() if m else (rest.split(maxsplit=1)[1] if len(rest.split(maxsplit=1)) > 1 else "")
            lines[i] = f"Usage: {prime_command}" + (f" {rest}" if rest else "")
            break
    text = "\n".join(lines)
    aliases = {**MODULE_TO_PRIME_COMMAND, module_name: prime_command}
    for mod in sorted(aliases, key=len, reverse=True):
        text = text.replace(f"python -m {mod}", aliases[mod]).replace(mod, aliases[mod])
    text = re.sub(r"\bvf-[a-z0-9-]+\b", prime_command, text)
    if prime_command in ("prime eval run", "prime gepa run"):
        text = re.sub(r"\benv_id_or_config\b", "environment", text)
    return text.rstrip() + "\n"


def load_help_text(module_name: str, prime_command: str) -> str:
    cmd = load_verifiers_prime_plugin(console=console).build_module_command(module_name, ["--help"])
    r = subprocess.run(cmd, capture_output=True, text=True)
    text = r.stdout + (("" if r.stdout.endswith("\n") or not r.stdout else "\n") + r.stderr if r.stderr else "")
    if not text.strip():
        raise RuntimeError(f"Unable to load help text from {module_name}")
    return sanitize_help_text(text, module_name, prime_command)


def _print_help(module_attr: str, prime_command: str, extra: list[str] | None = None) -> None:
    try:
        text = load_help_text(getattr(load_verifiers_prime_plugin(console=console), module_attr), prime_command)
    except Exception as e:
        console.print(f"[red]Failed to load help for {prime_command}:[/red] {e}")
        raise typer.Exit(1) from e
    if extra:
        have = text.rstrip("\n").splitlines()
        text = "\n".join(have + [x for x in extra if x not in have]) + "\n"
    sys.stdout.write(text)
    sys.stdout.flush()


def print_eval_run_help() -> None:
    _print_help("eval_module", "prime eval run", EXTRA_EVAL_HELP)


def print_gepa_run_help() -> None:
    _print_help("gepa_module", "prime gepa run")


def print_env_init_help() -> None:
    _print_help("init_module", "prime env init")


def print_env_build_help() -> None:
    _print_help("build_module", "prime env build")


def print_lab_setup_help() -> None:
    _print_help("setup_module", "prime lab setup")


# ------------------------------------------------------------------------------------------------ local ↔ hub state
def should_skip_directory(name: str) -> bool:
    return name.startswith(".") or name in SKIP_DIRS or name.endswith(".egg-info")


def is_valid_hash(v: object) -> bool:
    return isinstance(v, str) and len(v) == 64 and all(c in "0123456789abcdef" for c in v.lower())


def hashed_items(env_path: Path) -> list[tuple[str, Path]]:
    """The (kind, path) sequence that defines an environment's content identity, in hashing order."""
    items: list[tuple[str, Path]] = []
    for pattern in ("pyproject.toml", "*.py", "README.md"):
        items += [("file", f) for f in env_path.glob(pattern) if f.is_file()]
    for sub in sorted(env_path.iterdir(), key=lambda p: p.name):
        if not sub.is_dir() or should_skip_directory(sub.name):
            continue
        items.append(("dir", sub))
        for f in sorted(sub.rglob("*")):
            rel = f.relative_to(env_path).parts
            if f.is_file() and not f.name.startswith(".") and not any(should_skip_directory(p) for p in rel[:-1]):
                items.append(("file", f))
    items.sort(key=lambda it: it[1].relative_to(env_path).as_posix().lower())
    return items


def compute_local_content_hash(env_path: Path) -> str | None:
    if not env_path.is_dir():
        return None
    h = hashlib.sha256()
    for kind, path in hashed_items(env_path):
        rel = path.relative_to(env_path).as_posix()
        h.update(f"{kind}:{rel}".encode())
        if kind == "file":
            try:
                h.update(path.read_bytes())
            except OSError:
                return None
    return h.hexdigest()


def find_local_env_dir(env_name: str, env_dir_path: str) -> Path | None:
    d = Path(env_dir_path) / env_name.replace("-", "_")
    return d if d.is_dir() else None


def fetch_user_slug(client: APIClient) -> str | None:
    try:
        data = client.get("/user/whoami").get("data")
    except APIError:
        return None
    slug = data.get("slug") if isinstance(data, dict) else None
    return slug if isinstance(slug, str) and slug else None


def fetch_active_team_slug(client: APIClient, team_id: str | None) -> str | None:
    if not team_id:
        return None
    try:
        teams = client.get("/user/teams").get("data")
    except APIError:
        return None
    for t in teams if isinstance(teams, list) else []:
        if isinstance(t, dict) and str(t.get("teamId")) == str(team_id) and t.get("slug"):
            return t["slug"]
    return None


def fetch_remote_env_details(client: APIClient, owner: str, name: str, version: str = "latest") -> dict | None:
    try:
        resp = client.get(f"/environmentshub/{owner}/{name}/@{version}")
    except APIError:
        return None
    details = resp.get("data", resp) if isinstance(resp, dict) else None
    return details if isinstance(details, dict) else {}


def remote_version_and_hash(details: dict | None) -> tuple[str | None, str | None]:
    if not isinstance(details, dict):
        return None, None
    block = details.get("latest_version") if isinstance(details.get("latest_version"), dict) else details
    sem, ch = block.get("semantic_version"), block.get("content_hash") or block.get("sha256")
    return (sem if isinstance(sem, str) else None, ch if isinstance(ch, str) else None)


def local_env_status(env_name: str, local_dir: Path, client: APIClient | None) -> tuple[str, str | None, str | None, bool, str | None]:
    """→ (display id, tracked slug, platform url, recommend push, reason)."""
    md = get_environment_metadata(local_dir) or {}
    owner, name = md.get("owner"), md.get("name")
    if not (isinstance(owner, str) and isinstance(name, str) and owner and name):
        return f"{env_name} (local only)", None, None, True, "local_only"
    slug = f"{owner}/{name}"
    url = environment_url_from_slug(slug)
    ahead = (f"{env_name} (local - ahead of {slug})", slug, url, True, "ahead")
    remote = fetch_remote_env_details(client, owner, name) if client else None
    if remote is None:
        return ahead
    r_ver, r_hash = remote_version_and_hash(remote)
    l_hash = compute_local_content_hash(local_dir) or (md.get("content_hash") if is_valid_hash(md.get("content_hash")) else None)
    l_ver = md.get("version") if isinstance(md.get("version"), str) else None
    if is_valid_hash(l_hash) and is_valid_hash(r_hash):
        synced = l_hash == r_hash
    else:
        synced = bool(l_ver and r_ver and l_ver == r_ver)
    return (slug, slug, url, False, None) if synced else ahead


def choose_remote_owner(env_name: str, candidates: list[tuple[str, str]]) -> tuple[str, str]:
    if len(candidates) == 1:
        return candidates[0]
    if not sys.stdin.isatty():
        console.print(f"[yellow]Warning:[/yellow] Multiple remote owners matched '{env_name}'. Non-interactive mode selected {candidates[0][1]}.")
        return candidates[0]
    console.print(f"[cyan]Multiple remote environments found for '{env_name}':[/cyan]")
    for i, (label, slug) in enumerate(candidates, 1):
        console.print(f"  [cyan]({i})[/cyan] {slug} [dim]({label})[/dim]")
    while True:
        n = typer.prompt("Select owner", type=int, default=1)
        if 1 <= n <= len(candidates):
            return candidates[n - 1]
        console.print(f"[red]Invalid selection.[/red] Enter 1-{len(candidates)}.")


def _remote(original: str, owner: str, name: str, version: str | None) -> ResolvedEnvironment:
    slug = f"{owner}/{name}"
    return ResolvedEnvironment(original=original, env_name=name, install_mode="remote", install_slug=slug + (f"@{version}" if version else ""),
                               upstream_slug=slug, env_display_id=slug, platform_slug=slug, platform_url=environment_url_from_slug(slug))  # fmt: skip


def resolve_environment_reference(env_reference: str, env_dir_path: str, client: APIClient | None = None,
                                  config: Config | None = None) -> ResolvedEnvironment:  # fmt: skip
    base, version = split_version(env_reference)
    if is_slug_reference(base):
        parts = split_owner_and_name(base)
        if parts is None:
            console.print(f"[red]Invalid environment reference: {env_reference}[/red]")
            raise typer.Exit(1)
        return _remote(env_reference, parts[0], parts[1], version)
    env_name = base
    if client is None:
        try:
            client = APIClient(require_auth=False)
        except Exception:
            client = None
    if config is None:
        try:
            config = Config(writable=False)
        except Exception:
            config = None
    local = find_local_env_dir(env_name, env_dir_path)
    if local is not None:
        display, slug, url, push, why = local_env_status(env_name, local, client)
        return ResolvedEnvironment(original=env_reference, env_name=env_name, install_mode="local", upstream_slug=slug if (slug and not push) else None,
                                   env_display_id=display, platform_slug=slug, platform_url=url, recommend_push=push, push_reason=why,
                                   local_env_path=local)  # fmt: skip
    nothing = ResolvedEnvironment(original=env_reference, env_name=env_name, install_mode="none", env_display_id=env_name)
    if client is None or config is None:
        return nothing
    owners: list[tuple[str, str]] = []
    me = fetch_user_slug(client)
    if me:
        owners.append(("personal", me))
    team = fetch_active_team_slug(client, config.team_id)
    if team and team != me:
        owners.append(("team", team))
    found = [(label, o) for label, o in owners if fetch_remote_env_details(client, o, env_name) is not None]
    if not found and fetch_remote_env_details(client, PRIME_SLUG, env_name) is not None:
        found = [("official", PRIME_SLUG)]
    if not found:
        return nothing
    label, owner = choose_remote_owner(env_name, found)
    console.print(f"[dim]Using remote environment {owner}/{env_name} ({label})[/dim]")
    return _remote(env_reference, owner, env_name, version)


# ------------------------------------------------------------------------------------------------ running things
def run_command(command: list[str], env: dict[str, str] | None = None) -> None:
    rc = subprocess.run(command, env=env).returncode
    if rc != 0:
        raise typer.Exit(rc)


def prepare_environment(plugin: PrimeVerifiersPlugin, env_reference: str, env_dir_path: str) -> ResolvedEnvironment:
    from .commands.env import install_single_environment, is_environment_installed

    r = resolve_environment_reference(env_reference, env_dir_path)
    if r.env_display_id and r.install_mode != "none":
        console.print(f"[dim]Resolved source: {r.env_display_id}[/dim]")
    if r.install_mode == "local":
        console.print(f"[dim]Using local environment '{r.env_name}'[/dim]")
        run_command(plugin.build_module_command(plugin.install_module, [r.env_name, "--path", env_dir_path]))
    elif r.install_mode == "remote":
        base, ver = split_version(r.install_slug or "")
        if not is_environment_installed(base.split("/", 1)[1], ver) and not install_single_environment(r.install_slug):
            raise typer.Exit(1)
    elif not is_environment_installed(r.env_name, None):
        console.print(f"[yellow]Warning:[/yellow] No local checkout or matching remote environment found for '{r.env_name}'. "
                      "Continuing with installed package resolution.")  # fmt: skip
    return r


def _load_toml(path: Path, what: str) -> dict:
    try:
        return tomllib.loads(path.read_text())
    except Exception as e:
        console.print(f"[yellow]Warning:[/yellow] Could not parse {what} config {path}: {e}. Skipping pre-install.")
        return {}


def collect_eval_config_envs(config_path: Path, fallback_env_dir: str) -> list[tuple[str, str]]:
    raw = _load_toml(config_path, "eval")
    default_dir = raw.get("env_dir_path") if isinstance(raw.get("env_dir_path"), str) else fallback_env_dir
    out: list[tuple[str, str]] = []
    for e in raw.get("eval", []) if isinstance(raw.get("eval"), list) else []:
        if isinstance(e, dict) and isinstance(e.get("env_id"), str) and e["env_id"]:
            key = (e["env_id"], e["env_dir_path"] if isinstance(e.get("env_dir_path"), str) else default_dir)
            if key not in out:
                out.append(key)
    return out


def collect_gepa_config_env(config_path: Path, fallback_env_dir: str) -> tuple[str, str] | None:
    raw = _load_toml(config_path, "GEPA")
    env = raw.get("env")
    if not (isinstance(env, dict) and isinstance(env.get("env_id"), str) and env["env_id"]):
        return None
    return env["env_id"], raw["env_dir_path"] if isinstance(raw.get("env_dir_path"), str) else fallback_env_dir


def with_inference_defaults(passthrough: list[str], config: Config) -> tuple[list[str], dict[str, str], str, str]:
    """Fill in -b (inference URL) and -k (API key env var) when the user did not; → (args, env, model, base_url)."""
    args, env = list(passthrough), os.environ.copy()
    model = parse_value_option(args, "--model", "-m") or DEFAULT_MODEL
    base = (parse_value_option(args, "--api-base-url", "-b") or "").rstrip("/")
    if not base:
        base = (config.inference_url or "").strip().rstrip("/")
        if not base:
            console.print("[red]Inference URL not configured.[/red] Check [bold]prime config view[/bold].")
            raise typer.Exit(1)
        args += ["-b", base]
    if parse_value_option(args, "--api-key-var", "-k") is None:
        env["PRIME_API_KEY"] = config.api_key
        args += ["-k", "PRIME_API_KEY"]
    return args, env, model, base


def preflight(model: str, base_url: str, configured_base_url: str, make_client=None) -> None:
    """Against our own inference service only, two independent checks, each with its own client and its own time-out handling
    (reference: verifiers_bridge.py — model validation, then the billing probe): (1) the model exists; (2) one tiny completion, so a
    402 surfaces now and not after the environment was installed.  A time-out in either is a warning — slow-warming models still run —
    and does not skip the other.  ``make_client`` defaults to this module's ``InferenceClient`` looked up at call time."""
    if base_url != configured_base_url:
        return
    make = make_client or InferenceClient

    def step(what: str, call) -> None:
        try:
            call(make(timeout=EVAL_PREFLIGHT_TIMEOUT))
        except httpx.TimeoutException:
            console.print(f"[yellow]Timed out during {what} for '{model}'.[/yellow] Continuing: some thinking models warm up slowly.")
        except InferencePaymentRequiredError as e:
            console.print(f"[red]{e}[/red]")
            raise typer.Exit(1) from e
        except InferenceAPIError as e:
            console.print(f"[red]Invalid model:[/red] {e} \n\n[b]Use 'prime inference models' to see available models.[/b]")
            raise typer.Exit(1) from e

    step("model validation", lambda c: c.retrieve_model(model))
    step("the billing pre-flight", lambda c: c.chat_completion({"model": model, "messages": [{"role": "user", "content": "Reply with OK."}]}))


def format_push_command(r: ResolvedEnvironment) -> str:
    cmd = f"prime env push --path {r.local_env_path}" if r.local_env_path is not None else "prime env push"
    parts = split_owner_and_name(r.platform_slug) if r.platform_slug else None
    return cmd + (f" --owner {parts[0]}" if parts else "")


def print_source_footer(r: ResolvedEnvironment | None) -> None:
    if r is None:
        return
    if r.platform_url:
        console.print(f"[dim]Environment URL: {r.platform_url}[/dim]")
    if r.recommend_push:
        msg = {"ahead": f"Local environment is ahead of {r.platform_slug}.", "local_only": "Local environment is not linked to an upstream."}
        console.print(f"[yellow]{msg.get(r.push_reason or '', 'Local environment differs from the current platform version.')}[/yellow]")
        console.print(f"[dim]Publish the current local version with:[/dim] {format_push_command(r)}")


def _require_key(config: Config) -> None:
    if not config.api_key:
        console.print("[red]No API key configured.[/red] Run [bold]prime login[/bold] or [bold]prime config set-api-key[/bold].")
        raise typer.Exit(1)


def run_eval_tui(env_dir: str | None, outputs_dir: str | None) -> None:
    plugin = load_verifiers_prime_plugin(console=console)
    env = {**os.environ, "VF_ENV_DIR": env_dir or "./environments", "VF_OUTPUTS_DIR": outputs_dir or "./outputs"}
    run_command(plugin.build_module_command(plugin.tui_module), env=env)


def run_eval_passthrough(environment: str, passthrough_args: list[str], *, skip_upload: bool, env_path: str | None) -> None:
    plugin, config = load_verifiers_prime_plugin(console=console), Config()
    _require_key(config)
    args, env, model, base_url = with_inference_defaults(passthrough_args, config)
    preflight(model, base_url, (config.inference_url or "").strip().rstrip("/"))
    env_dir = parse_value_option(args, "--env-dir-path", "-p") or DEFAULT_ENV_DIR_PATH
    resolved: ResolvedEnvironment | None = None
    target = environment
    if is_config_target(environment):
        for ref, d in collect_eval_config_envs(Path(environment), env_dir):
            prepare_environment(plugin, ref, d)
    else:
        resolved = _prepare_single_environment(plugin, environment, env_dir)
        target = resolved.env_name
        if resolved.env_display_id:
            args += ["--header", f"{INTERNAL_ENV_DISPLAY_HEADER}: {resolved.env_display_id}"]
    if not skip_upload and not has_flag(args, "--save-results", "-s"):
        args.append("-s")
    job_id = build_job_id(resolved.env_name if resolved else Path(environment).stem, model)
    args += ["--header", f"X-PI-Job-Id: {job_id}"]
    if config.team_id:
        args += ["--header", f"X-Prime-Team-ID: {config.team_id}"]
    console.print(f"[dim]Eval job_id: {job_id}[/dim]")
    run_command(plugin.build_module_command(plugin.eval_module, [target, *args]), env=env)  # ← all model traffic happens in the child

    if skip_upload:
        print_source_footer(resolved)
        console.print("[dim]Skipped uploading evaluation results[/dim]")
        return
    if resolved is None:
        console.print("[yellow]Evaluation completed. Automatic upload is skipped for config-driven runs.[/yellow]")
        return
    if resolved.recommend_push:
        print_source_footer(resolved)
        console.print("[yellow]Evaluation completed. Automatic upload is skipped until the local environment is published.[/yellow]")
        return
    upstream = resolved.upstream_slug
    if upstream is None:
        md = find_environment_metadata(env_name=resolved.env_name, env_path=Path(env_path) if env_path else Path.cwd(),
                                       module_name=resolved.env_name.replace("-", "_"))  # fmt: skip
        if md and md.get("owner") and md.get("name"):
            upstream = f"{md['owner']}/{md['name']}"
            console.print(f"[dim]Using upstream environment {upstream}[/dim]")
    if upstream is None:
        print_source_footer(resolved)
        console.print("[dim]No upstream environment found. Skipped uploading evaluation results to platform.\n"
                      "Use `prime env push` to set an upstream, or use `--env-path` to specify the correct environment path.[/dim]")  # fmt: skip
        return
    if resolved.platform_url:
        console.print(f"[dim]Environment URL: {resolved.platform_url}[/dim]")
    try:
        push_eval_results_to_hub(env_name=resolved.env_name, model=model, job_id=job_id, env_path=Path(env_path) if env_path else None,
                                 upstream_slug=upstream)  # fmt: skip
    except Exception as e:
        console.print(f"[red]Failed to push results to hub:[/red] {e}\n[yellow]Evaluation completed but results were not pushed.[/yellow]")
        raise typer.Exit(1) from e


def run_gepa_passthrough(environment_or_config: str, passthrough_args: list[str]) -> None:
    plugin, config = load_verifiers_prime_plugin(console=console), Config()
    _require_key(config)
    args, env, _model, _base = with_inference_defaults(passthrough_args, config)
    env_dir = parse_value_option(args, "--env-dir-path", "-p") or DEFAULT_ENV_DIR_PATH
    target = environment_or_config
    if is_config_target(environment_or_config):
        ce = collect_gepa_config_env(Path(environment_or_config), env_dir)
        if ce is not None:
            prepare_environment(plugin, *ce)
    else:
        target = prepare_environment(plugin, environment_or_config, env_dir).env_name
    run_command(plugin.build_module_command(plugin.gepa_module, [target, *args]), env=env)


# the reference's private spellings: its tests import the first and replace the second to stop a run after the pre-flight
_sanitize_help_text = sanitize_help_text
_prepare_single_environment = prepare_environment
