"""``prime env`` — the Environments Hub: browse, push (wheel + source archive through signed URLs), pull, install
(simple index → wheel URL → private pull-and-build cache), versions, CI actions; secrets/variables live in env_secrets.py
(reference: packages/prime/src/prime_cli/commands/env.py:48-3659)."""

from __future__ import annotations

import hashlib
import json
import re
import shutil
import subprocess
import tarfile
import tempfile
import time
from datetime import datetime
from pathlib import Path
from typing import Any, Optional

import httpx
import toml
import typer

from ..core import APIClient, APIError
from ..utils.display import colorize, output_data_as_json, validate_output_format
from ..utils.env_metadata import find_environment_metadata, get_environment_metadata, write_environment_metadata
from ..utils.formatters import format_size, strip_ansi
from ..utils.time_utils import format_time_ago, iso_timestamp
from ..verifiers_bridge import is_help_request, print_env_build_help, print_env_init_help
from ..verifiers_plugin import load_verifiers_prime_plugin
from . import env_packaging as pk
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app
from .env_packaging import compute_content_hash, is_environment_installed  # noqa: F401  (public re-exports)

app = make_app("Manage verifiers environments")
action_app = make_app("Manage environment actions (CI jobs)")
version_app = make_app("Manage environment versions")
app.add_typer(action_app, name="action", rich_help_panel="Manage")
app.add_typer(version_app, name="version", rich_help_panel="Manage")

DEFAULT_LIST_LIMIT = 20
MAX_FILES_TO_SHOW = 10
ACTION_COLORS = {"SUCCESS": "green", "FAILED": "red", "RUNNING": "yellow", "PENDING": "dim", "CANCELLED": "dim"}
PASSTHROUGH = {"allow_extra_args": True, "ignore_unknown_options": True, "help_option_names": []}


def _data(resp: dict[str, Any]) -> Any:
    return resp.get("data", resp) if isinstance(resp, dict) else resp


def parse_slug(environment: str) -> tuple[str, str]:
    try:
        return pk.parse_environment_slug(environment)
    except ValueError as e:
        raise fail(str(e))


def resolve_environment(environment: Optional[str]) -> tuple[str, str]:
    """Explicit slug, else the upstream recorded in the current directory's metadata."""
    if environment:
        return parse_slug(environment)
    md = find_environment_metadata() or {}
    if md.get("owner") and md.get("name"):
        console.print(f"[dim]Using environment: {md['owner']}/{md['name']}[/dim]")
        return md["owner"], md["name"]
    raise fail("No environment specified and none detected in current directory")


def get_environment_id(client: APIClient, owner: str, name: str) -> str:
    env_id = (client.get(f"/environmentshub/{owner}/{name}/@latest").get("data") or {}).get("id")
    if not env_id:
        raise APIError(f"Environment {owner}/{name} not found")
    return env_id


def display_upstream_environment_info(env_path: Path | None = None, environment_name: str | None = None) -> bool:
    md = find_environment_metadata(env_name=environment_name, env_path=env_path,
                                   module_name=environment_name.replace("-", "_") if environment_name else None)  # fmt: skip
    if md and md.get("owner") and md.get("name"):
        console.print(f"[dim]Using upstream environment {md['owner']}/{md['name']}[/dim]\n")
        return True
    console.print("[dim]No upstream environment found.[/dim]\n")
    return False


def _download(url: str, dest: Path, api_key: str | None) -> None:
    if not pk.is_valid_url(url):
        raise ValueError(f"Invalid download URL: {url}")
    headers = {"Authorization": f"Bearer {api_key}"} if api_key else {}
    with httpx.stream("GET", url, headers=headers, timeout=60.0, follow_redirects=True) as r:
        r.raise_for_status()
        with open(dest, "wb") as f:
            for chunk in r.iter_bytes(chunk_size=1 << 16):
                f.write(chunk)


def _upload(url: str, path: Path) -> None:
    httpx.put(url, content=path.read_bytes(), headers={"Content-Type": "application/octet-stream"}, timeout=300.0).raise_for_status()


# ======================================================================================================= explore
@app.command("list", rich_help_panel="Explore")
@handle_errors
def list_cmd(
    num: int = typer.Option(DEFAULT_LIST_LIMIT, "--num", "-n", help="Items per page"),
    page: int = typer.Option(1, "--page", "-p", help="Page number"),
    owner: Optional[str] = typer.Option(None, "--owner", help="Filter by owner name"),
    visibility: Optional[str] = typer.Option(None, "--visibility", help="PUBLIC or PRIVATE"),
    output: str = typer.Option("table", "--output", help="Output format: table or json"),
    search: Optional[str] = typer.Option(None, "--search", "-s", help="Search by name or description"),
    tag: Optional[list[str]] = typer.Option(None, "--tag", "-t", help="Filter by tag (repeatable)"),
    action_status: Optional[str] = typer.Option(None, "--action-status", help="SUCCESS/FAILED/RUNNING/PENDING"),
    sort: str = typer.Option("created_at", "--sort", help="name, created_at, updated_at, stars"),
    order: str = typer.Option("desc", "--order", help="asc or desc"),
    show_actions: bool = typer.Option(False, "--show-actions", help="Show action status column"),
    starred: bool = typer.Option(False, "--starred", help="Only environments you starred"),
    mine: bool = typer.Option(False, "--mine", help="Only your own environments (personal + team)"),
) -> None:
    """List environments from the hub (public ones, plus private ones you can see when logged in)."""
    validate_output_format(output, console)
    if num < 1 or page < 1:
        raise fail("--num and --page must be at least 1")
    if sort not in ("name", "created_at", "updated_at", "stars"):
        raise fail("--sort must be one of: name, created_at, updated_at, stars")
    if order.lower() not in ("asc", "desc"):
        raise fail("--order must be one of: asc, desc")
    with_ci = show_actions or bool(action_status)
    params: dict[str, Any] = {"include_teams": True, "limit": num, "offset": (page - 1) * num, "sort_by": sort, "sort_order": order}
    for key, val in (("owner", owner), ("visibility", visibility), ("search", search), ("tags", tag), ("ci_status", action_status)):
        if val:
            params[key] = val
    for key, on in (("include_ci_status", with_ci), ("starred_only", starred), ("mine_only", mine)):
        if on:
            params[key] = True
    res = api(require_auth=starred or mine).get("/environmentshub/", params=params)
    envs = res.get("data", res.get("environments", []))
    total = res.get("total_count", res.get("total", 0))
    if not envs and output != "json":
        console.print("[yellow]No more results.[/yellow]" if page > 1 else "[yellow]No environments found.[/yellow]")
        return
    entries = []
    for e in envs:
        item = {"environment": f"{e['owner']['name']}/{e['name']}", "description": e.get("description", ""),
                "visibility": e.get("visibility", ""), "version": e.get("latest_version"), "stars": e.get("stars", 0),
                "updated_at": e.get("updated_at")}  # fmt: skip
        if with_ci:
            item["action_status"] = e.get("latest_ci_status")
        if e.get("tags"):
            item["tags"] = e["tags"]
        entries.append(item)
    cols: list[Any] = [("Environment", "cyan"), ("Description", "green"), ("Version", "blue"), ("Stars", "yellow"), ("Updated", "dim")]
    if with_ci:
        cols.append("Action Status")
    rows = [[it["environment"], it["description"], it["version"] or "-", str(it["stars"]), (it["updated_at"] or "")[:10],
             *([colorize(it.get("action_status") or "-", ACTION_COLORS)] if with_ci else [])] for it in entries]  # fmt: skip
    footer = (f"\n[yellow]Showing page {page} of results. Use --page {page + 1} to see more.[/yellow]" if total > page * num
              else f"\n[dim]Total: {total} environment(s)[/dim]")  # fmt: skip
    emit(output, {"environments": entries, "total": total, "page": page, "per_page": num}, f"Environments (Total: {total})", cols, rows, footer)


@app.command("status", rich_help_panel="Explore")
@handle_errors
def status_cmd(env_id: str = typer.Argument(..., help="Environment ID (owner/name)"),
               output: str = typer.Option("table", "--output", help="Output format: table or json")) -> None:  # fmt: skip
    """Show the latest version and CI action status of an environment."""
    validate_output_format(output, console)
    owner, name = parse_slug(env_id)
    data = _data(api(require_auth=False).get(f"/environmentshub/{owner}/{name}/status"))
    if output == "json":
        output_data_as_json(data, console)
        return
    console.print(f"\n[bold cyan]Environment:[/bold cyan] {owner}/{data.get('name', name)}")
    if data.get("description"):
        console.print(f"[dim]Description:[/dim] {data['description']}")
    console.print(f"[dim]Visibility:[/dim] {data.get('visibility', 'UNKNOWN')}")
    console.print("\n[bold]Latest Version:[/bold]")
    lv = data.get("latest_version")
    if lv:
        ch = lv.get("content_hash") or ""
        console.print(f"  Version: {lv.get('semantic_version') or ch[:8]}")
        console.print(f"  Hash: {(ch or '-')[:12]}")
        console.print(f"  Created: {format_time_ago(lv.get('created_at'))}")
    else:
        console.print("  [dim]No versions found[/dim]")
    act = data.get("action")
    if act:
        console.print("\n[bold]Action Status:[/bold]")
        console.print(f"  Status: {colorize(act.get('status') or '-', ACTION_COLORS)}")
        if act.get("job_id"):
            console.print(f"  Job ID: [dim]{act['job_id']}[/dim]")
    console.print()


@app.command(no_args_is_help=True, rich_help_panel="Explore")
@handle_errors
def info(env_id: str = typer.Argument(..., help="Environment ID (owner/name)"),
         version: str = typer.Option("latest", "--version", "-v", help="Version to show")) -> None:  # fmt: skip
    """Show environment details and the ways to install it."""
    try:
        base, parsed = pk.validate_env_id(env_id)
    except ValueError as e:
        raise fail(str(e))
    target = parsed if parsed != "latest" else version
    owner, name = base.split("/")
    console.print(f"Fetching {base}@{target}...")
    d = fetch_environment_details(api(require_auth=False), owner, name, target)
    wheel_url, index = pk.process_wheel_url(d.get("wheel_url")), d.get("simple_index_url")
    console.print(f"\n[bold cyan]{owner}/{name}[/bold cyan][dim]@{target}[/dim]")
    if (d.get("metadata") or {}).get("description"):
        console.print(f"[dim]{d['metadata']['description']}[/dim]")
    console.print()
    pkg = pk.normalize_package_name(name)
    sh = "  [green]$[/green] "
    if wheel_url or index:
        console.print("[bold yellow]Install (choose one)[/bold yellow]")
        console.print(f"{sh}prime env install {owner}/{name}@{target}")
        if index:
            spec = f"{pkg}=={target}" if target != "latest" else pkg
            console.print(f"{sh}uv pip install {spec} --extra-index-url {index}")
            console.print(f"{sh}uv add {spec} --index {index}")
            console.print(f"{sh}pip install {spec} --extra-index-url {index}")
        else:
            console.print(f"{sh}uv pip install {wheel_url}")
            console.print(f"{sh}uv add {pkg}@{wheel_url}")
            console.print(f"{sh}pip install {wheel_url}")
        console.print("\n[bold yellow]Usage[/bold yellow]")
        console.print("  [blue]>>>[/blue] from verifiers import load_environment")
        console.print(f"  [blue]>>>[/blue] env = load_environment('{name}')")
    elif d.get("visibility") == "PRIVATE":
        console.print("[bold yellow]Install (private environment)[/bold yellow]")
        console.print(f"{sh}prime env pull {owner}/{name}@{target}")
        console.print("  [dim]Note: direct uv/pip install is not available for private environments[/dim]")
        console.print("\n[bold yellow]After pulling[/bold yellow]")
        console.print(f"{sh}cd <target_directory>")
        console.print(f"{sh}uv pip install -e .")
    else:
        console.print("[yellow]No wheel available for this version[/yellow]")
    console.print()


def fetch_environment_details(client: APIClient, owner: str, name: str, version: str) -> dict[str, Any]:
    return _data(client.get(f"/environmentshub/{owner}/{name}/@{version}"))


# ======================================================================================================= push
def _resolve_push_path(path: str | None, env_id: str | None) -> Path:
    """Where ``prime env push [ENV_ID]`` looks: ``--path`` (default ``.``), or with an id ``<--path | ./environments>/<id with _>``."""
    if env_id:
        return (Path(path or "./environments") / env_id.split("/")[-1].replace("-", "_")).resolve()
    return Path(path or ".").resolve()


def _resolve_pull_path(target: str | None, env_name: str) -> Path:
    """Where ``prime env pull`` unpacks: ``--target`` verbatim, else the importable folder name (``-`` → ``_``) under ``./environments``
    when the project has that directory, else under the working directory (reference: commands/env.py:845-853)."""
    if target:
        return Path(target)
    cwd = Path.cwd()
    return (cwd / "environments" if (cwd / "environments").is_dir() else cwd) / env_name.replace("-", "_")


def _resolve_with_username(client: APIClient, payload: dict[str, Any]) -> dict[str, Any]:
    """POST /resolve; if the account has no public username yet, ask for one (set once), then retry."""
    try:
        return _data(client.post("/environmentshub/resolve", json=payload))
    except APIError as e:
        if "missing a username" not in str(e).lower():
            raise fail(f"Failed to resolve environment: {e}")
    console.print("[yellow]Your user profile is missing a username.[/yellow] You must choose one to publish environments.")
    console.print("[dim]Note: the username can only be chosen once and will be public.[/dim]")
    while True:
        try:
            chosen = typer.prompt("Enter your desired username").strip().lower()
        except typer.Abort:
            raise fail("Cancelled by user")
        if not re.match(r"^[a-z0-9-]{3,30}$", chosen):
            console.print("[red]Invalid username.[/red] Use 3-30 chars: lowercase letters, numbers and '-'.")
            continue
        try:
            client.patch("/user/slug", json={"slug": chosen})
            console.print(f"[green]✓ Username set to {chosen}[/green]")
            break
        except APIError as se:
            if "409" in str(se) or "already taken" in str(se).lower():
                console.print("[red]That username is already taken.[/red] Please choose another.")
                continue
            raise fail(f"Failed to set username: {se}")
    try:
        return _data(client.post("/environmentshub/resolve", json=payload))
    except APIError as e2:
        raise fail(f"Failed to resolve environment after setting username: {e2}")


def _duplicate_hash_hint(what: str, e: APIError) -> "typer.Exit":
    console.print(f"[red]Failed to prepare {what} upload: {e}[/red]")
    if "content hash" in str(e).lower() and "already exists" in str(e):
        console.print("[yellow]Tip: the content hash covers your source files (*.py, pyproject.toml, README.md, sub-packages); "
                      "nothing changed since the last push.[/yellow]")  # fmt: skip
        console.print("[dim]Use --auto-bump to publish a new version without content changes.[/dim]")
    return typer.Exit(1)


def _save_push_metadata(env_path: Path, env_id: str, owner_name: str, env_name: str, wheel_sha256: str) -> None:
    existing = get_environment_metadata(env_path) or {}
    changed = bool(existing) and (existing.get("owner") != owner_name or existing.get("name") != env_name)
    target = write_environment_metadata(env_path, {**existing, "environment_id": env_id, "owner": owner_name, "name": env_name,
                                                   "pushed_at": datetime.now().isoformat(), "wheel_sha256": wheel_sha256})  # fmt: skip
    console.print(f"[dim]{'Updated environment metadata in' if existing else 'Saved environment metadata to'} {target}[/dim]", soft_wrap=True)
    if changed:
        console.print(f"[dim]Upstream set to {owner_name}/{env_name}[/dim]")


@app.command(rich_help_panel="Manage")
def push(
    env_id: Optional[str] = typer.Argument(None, help="Environment ID used as the local folder name (hyphens → underscores)"),
    path: Optional[str] = typer.Option(None, "--path", "-p", help="Environment directory ('.' without ENV_ID; parent './environments' with it)"),
    name: Optional[str] = typer.Option(None, "--name", "-n", help="Override environment name (default: pyproject name)"),
    owner: Optional[str] = typer.Option(None, "--owner", "-o", help="Owner slug (user or team) to push to"),
    team: Optional[str] = typer.Option(None, "--team", "-t", help="Team slug for team ownership"),
    visibility: Optional[str] = typer.Option(None, "--visibility", "-v", help="PUBLIC or PRIVATE"),
    auto_bump: bool = typer.Option(False, "--auto-bump", help="Bump the patch version before pushing"),
    rc: bool = typer.Option(False, "--rc", help="Bump or create an rc pre-release"),
    post: bool = typer.Option(False, "--post", help="Bump or create a .post release"),
) -> None:
    """Build the environment's wheel and publish wheel + source archive to the hub."""
    env_path = _resolve_push_path(path, env_id)
    display_upstream_environment_info(env_path)
    pyproject = env_path / "pyproject.toml"
    if not pyproject.exists():
        raise fail("pyproject.toml not found")
    try:
        project = toml.load(pyproject).get("project", {})
    except Exception as e:
        raise fail(f"Failed to parse pyproject.toml: {e}")
    env_name = name or project.get("name")
    if not env_name:
        raise fail("No name found in pyproject.toml and no --name provided")
    if auto_bump or rc or post:
        if sum((auto_bump, rc, post)) > 1:
            raise fail("--auto-bump, --rc, and --post are mutually exclusive")
        cur = project.get("version")
        if not cur:
            raise fail("No version found in pyproject.toml for auto-bump")
        try:
            new = pk.bump_version(cur) if auto_bump else pk.bump_rc_version(cur) if rc else pk.bump_post_version(cur)
            console.print(f"Auto-bumping version: {cur} → {new}")
            pk.update_pyproject_version(pyproject, new)
            project = toml.load(pyproject).get("project", {})
            console.print("[green]✓ Updated version in pyproject.toml[/green]")
        except Exception as e:
            raise fail(f"Failed to update version in pyproject.toml: {e}")
    console.print(f"Environment name: {env_name}")
    if not pk.has_environment_code(env_path):
        raise fail("No environment Python file found")

    console.print(f"Building environment package at {env_path}...")
    try:
        wheel = pk.build_wheel(env_path)
    except subprocess.CalledProcessError as e:
        console.print("[red]Build failed![/red]")
        console.print(e.stderr or "")
        raise typer.Exit(1)
    except FileNotFoundError as e:
        raise fail(f"Build tool or wheel not found ({e}). Install 'uv' or 'build'.")
    console.print(f"[green]✓ Built {wheel.name} ({wheel.stat().st_size:,} bytes)[/green]")

    console.print("\nUploading to Prime Intellect Hub...")
    try:
        client = api()
        payload: dict[str, Any] = {"name": env_name}
        if visibility:
            payload["visibility"] = visibility
        if owner:
            payload["owner_slug"] = owner
        elif team:
            payload["team_slug"] = team
        elif client.config.team_id:
            payload["team_id"] = client.config.team_id
        console.print("Resolving environment...")
        resolved = _resolve_with_username(client, payload)
        hub_id, owner_name = resolved["id"], resolved["owner"]["name"]
        console.print(f"[green]✓ {'Created' if resolved.get('created') else 'Found existing'} environment: {owner_name}/{env_name}[/green]")

        console.print("Uploading wheel ...")
        wheel_sha = hashlib.sha256(wheel.read_bytes()).hexdigest()
        content_hash = compute_content_hash(env_path)
        version = project.get("version")
        meta = {"description": project.get("description", ""), "tags": project.get("tags", []), "license": project.get("license", ""),
                "dependencies": project.get("dependencies", []), "python_requires": project.get("requires-python", ">=3.8"),
                "original_filename": wheel.name, "requires_dist": pk.extract_requires_dist_from_wheel(wheel)}  # fmt: skip
        try:
            w = client.post(f"/environmentshub/{hub_id}/wheels", json={"content_hash": content_hash, "filename": wheel.name, "sha256": wheel_sha,
                            "size": wheel.stat().st_size, "semantic_version": version, "metadata": meta})["data"]  # fmt: skip
        except APIError as e:
            raise _duplicate_hash_hint("wheel", e)
        if w.get("upload_url"):
            _upload(w["upload_url"], wheel)
            client.post(f"/environmentshub/{hub_id}/wheels/{w['wheel_id']}/finalize")

        console.print("Creating source archive...")
        with tempfile.TemporaryDirectory(prefix="prime_push_") as td:
            archive = Path(td) / "source.tar.gz"
            size = pk.build_source_archive(env_path, archive)
            console.print(f"Source archive size: {format_size(size)}")
            if size > pk.MAX_TARBALL_SIZE_LIMIT:
                console.print(f"\n[yellow]⚠ Warning: the archive ({format_size(size)}) exceeds the recommended {format_size(pk.MAX_TARBALL_SIZE_LIMIT)}. "
                              "Exclude data files, model weights and build artefacts (a root .gitignore is honoured).[/yellow]\n")  # fmt: skip
            try:
                v = client.post(f"/environmentshub/{hub_id}/versions", json={
                    "content_hash": content_hash, "filename": f"{env_name}-{version}-{content_hash[:8]}.tar.gz",
                    "sha256": hashlib.sha256(archive.read_bytes()).hexdigest(), "semantic_version": version,
                    "metadata": {**meta, "original_filename": f"{env_name}-{version}.tar.gz"}})["data"]  # fmt: skip
            except APIError as e:
                raise _duplicate_hash_hint("source", e)
            _upload(v["upload_url"], archive)
            fin = client.post(f"/environmentshub/{hub_id}/versions/{v['version_id']}/finalize")["data"]
        if not fin.get("success"):
            raise fail(f"Error finalizing: {fin.get('message')}")
        console.print(f"\n[green]✓ Successfully pushed {owner_name}/{env_name}[/green]")
        console.print(f"Wheel: {wheel.name}")
        console.print(f"SHA256: {wheel_sha}")
        try:
            _save_push_metadata(env_path, hub_id, owner_name, env_name, wheel_sha)
        except Exception as e:
            console.print(f"[yellow]Warning: Could not save environment metadata: {e}[/yellow]")
        hub_url = f"{client.config.frontend_url.rstrip('/')}/dashboard/environments/{owner_name}/{env_name}"
        console.print("\n[cyan]View on Environments Hub:[/cyan]")
        console.print(f"  [link={hub_url}]{hub_url}[/link]", soft_wrap=True)
        console.print("\n[cyan]Install with:[/cyan]")
        console.print(f"  prime env install {owner_name}/{env_name}")
    except typer.Exit:
        raise
    except APIError as e:
        raise fail(f"API Error: {e}")
    except (httpx.HTTPError, OSError, tarfile.TarError, KeyError) as e:
        raise fail(f"Upload failed: {e}")


# ======================================================================================================= init / build (verifiers)
def _delegate(ctx: typer.Context, first: Optional[str], module_attr: str, what: str, example: str, help_fn) -> None:
    extra = list(ctx.args)
    if is_help_request(first or "", extra):
        help_fn()
        raise typer.Exit(0)
    if first is None or first.startswith("-"):
        console.print(f"[red]Error:[/red] {'Missing argument ' + repr(what.upper()) if first is None else what + ' must be the first argument'}.")
        console.print(f"[dim]Example: {example}[/dim]")
        raise typer.Exit(2)
    plugin = load_verifiers_prime_plugin(console=console)
    rc = subprocess.run(plugin.build_module_command(getattr(plugin, module_attr), [first, *extra])).returncode
    if rc != 0:
        raise typer.Exit(rc)


@app.command(no_args_is_help=True, rich_help_panel="Manage", context_settings=PASSTHROUGH)
def init(ctx: typer.Context, name: Optional[str] = typer.Argument(None, help="Name of the new environment")) -> None:
    """Initialize a new environment from the verifiers template."""
    _delegate(ctx, name, "init_module", "name", "prime env init my-env --path ./environments", print_env_init_help)


@app.command(no_args_is_help=True, rich_help_panel="Manage", context_settings=PASSTHROUGH)
def build(ctx: typer.Context, env_id: Optional[str] = typer.Argument(None, help="Environment ID (e.g. openenv-echo)")) -> None:
    """Build an OpenEnv-backed environment image."""
    _delegate(ctx, env_id, "build_module", "env_id", "prime env build openenv-echo --path ./environments", print_env_build_help)


# ======================================================================================================= pull
@app.command(no_args_is_help=True, rich_help_panel="Manage")
@handle_errors
def pull(env_id: str = typer.Argument(..., help="Environment ID (owner/name or owner/name@version)"),
         target: Optional[str] = typer.Option(None, "--target", "-t", help="Target directory"),
         version: str = typer.Option("latest", "--version", "-v", help="Version to pull")) -> None:  # fmt: skip
    """Download an environment's source for local inspection."""
    if "@" in env_id:
        env_id, version = env_id.rsplit("@", 1)
    parts = env_id.split("/")
    if len(parts) != 2:
        raise fail("Invalid environment ID format. Expected: owner/name")
    owner, name = parts
    client = api(require_auth=False)
    console.print(f"Pulling {env_id}@{version}...")
    details = fetch_environment_details(client, owner, name, version)
    url = details.get("package_url")
    if not url:
        raise fail("No downloadable package found")
    base = _resolve_pull_path(target, name)
    dest = base
    if not target and dest.exists():
        i = 1
        while dest.exists():
            dest = base.parent / f"{base.name}-{i}"
            i += 1
        console.print(f"[yellow]Directory {base} already exists. Using {dest} instead.[/yellow]")
    try:
        dest.mkdir(parents=True, exist_ok=True)
        console.print(f"Downloading to {dest}...")
        with tempfile.TemporaryDirectory(prefix="prime_pull_") as td:
            tmp = Path(td) / "src.tar.gz"
            _download(url, tmp, client.api_key)
            with tarfile.open(tmp, "r:gz") as tar:
                pk.safe_tar_extract(tar, dest)
    except (httpx.HTTPError, OSError, tarfile.TarError, ValueError) as e:
        raise fail(f"Pull failed: {e}")
    console.print(f"[green]✓ Environment pulled to {dest}[/green]")
    try:
        md = write_environment_metadata(dest, {"environment_id": details.get("id"), "owner": owner, "name": name})
        console.print(f"[dim]Created environment metadata at {md}[/dim]", soft_wrap=True)
    except OSError as e:
        console.print(f"[yellow]Warning: Could not create metadata file: {e}[/yellow]")
    files = sorted(f.name for f in dest.iterdir() if f.name not in (".prime", ".env-metadata.json"))
    if files:
        console.print("\nExtracted files:")
        for f in files[:MAX_FILES_TO_SHOW]:
            console.print(f"  - {f}")
        if len(files) > MAX_FILES_TO_SHOW:
            console.print(f"  ... and {len(files) - MAX_FILES_TO_SHOW} more files")


# ======================================================================================================= install / uninstall
def _version_from_pyproject(env_path: Path) -> str | None:
    try:
        return toml.load(env_path / "pyproject.toml").get("project", {}).get("version")
    except Exception:
        return None


def _cached_wheel(cache_dir: Path, owner: str, name: str, version: str) -> tuple[Path, Path | None]:
    slot = cache_dir / owner / name / version
    if not slot.resolve().is_relative_to(cache_dir.resolve()):
        raise ValueError("Cache path escapes cache directory")
    wheels = sorted((slot / "dist").glob("*.whl")) if (slot / "dist").exists() else []
    return slot, (wheels[0] if wheels else None)


def pull_and_build_private_env(client: APIClient, owner: str, name: str, version: str, details: dict[str, Any]) -> tuple[Path, str]:
    """Private environments have no index/wheel URL: download the source with the user's key, build the wheel once into
    ``~/.prime/wheel_cache/<owner>/<name>/<version>`` and return (wheel, resolved version)."""
    for comp, what in ((owner, "owner"), (name, "name"), (version, "version")):
        pk.validate_path_component(comp, what)
    url = details.get("package_url")
    if not url:
        raise ValueError("No downloadable package found for private environment")
    cache = pk.env_cache_dir()
    if version != "latest":
        _, hit = _cached_wheel(cache, owner, name, version)
        if hit:
            console.print(f"[dim]Using cached wheel at {hit}[/dim]", soft_wrap=True)
            return hit, version
    with tempfile.TemporaryDirectory(prefix="prime_env_") as td:
        tmp, extracted = Path(td) / "src.tar.gz", Path(td) / "x"
        extracted.mkdir()
        _download(url, tmp, client.api_key)
        with tarfile.open(tmp, "r:gz") as tar:
            pk.safe_tar_extract(tar, extracted)
        actual = _version_from_pyproject(extracted) or version
        pk.validate_path_component(actual, "version")
        slot, hit = _cached_wheel(cache, owner, name, actual)
        if hit:
            console.print(f"[dim]Using cached wheel at {hit}[/dim]", soft_wrap=True)
            return hit, actual
        slot.mkdir(parents=True, exist_ok=True)
        for item in extracted.iterdir():
            shutil.move(str(item), str(slot / item.name))
    console.print("[dim]Building wheel...[/dim]")
    try:
        wheel = pk.build_wheel(slot)
    except subprocess.CalledProcessError as e:
        raise RuntimeError(f"Failed to build wheel: {e.stderr}") from e
    try:
        write_environment_metadata(slot, {"environment_id": details.get("id"), "owner": owner, "name": name, "version": actual,
                                          "cached_at": datetime.now().isoformat(), "wheel_path": str(wheel)})  # fmt: skip
    except OSError:
        pass
    return wheel, actual


def execute_install_command(cmd: list[str], env_id: str, version: str, tool: str) -> None:
    console.print(f"\n[cyan]Installing {env_id}@{version} with {tool}...[/cyan]")
    console.print(f"[dim]Running: {' '.join(cmd)}[/dim]", soft_wrap=True)
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, bufsize=1)
    assert proc.stdout is not None
    for line in proc.stdout:
        console.print(line.rstrip(), markup=False, highlight=False)
    if proc.wait() != 0:
        raise RuntimeError(f"installation failed with exit code {proc.returncode}")
    console.print(f"[green]✓ Successfully installed {env_id}@{version}[/green]")


def _private_install_cmd(wheel: Path, name: str, tool: str, no_upgrade: bool) -> list[str]:
    if tool == "uv":
        return pk.uv_pip_command("install", *([] if no_upgrade else ["-P", pk.normalize_package_name(name)]), str(wheel))
    return ["pip", "install", str(wheel), *([] if no_upgrade else ["--upgrade"])]


def _resolve_hub_install(client: APIClient, slug: str, tool: str, no_upgrade: bool) -> tuple[list[str], str, str, str]:
    """→ (command, owner/name, version, package name); raises ValueError (skip) / APIError / RuntimeError (fail)."""
    base, version = pk.validate_env_id(slug)
    owner, name = base.split("/")
    d = fetch_environment_details(client, owner, name, version)
    index, wheel_url = d.get("simple_index_url"), pk.process_wheel_url(d.get("wheel_url"))
    if not index and not wheel_url:
        if d.get("visibility") != "PRIVATE":
            raise ValueError("No installation method available")
        console.print("[dim]Private environment detected, pulling and building...[/dim]")
        wheel, version = pull_and_build_private_env(client, owner, name, version, d)
        return _private_install_cmd(wheel, name, tool, no_upgrade), base, version, name
    cmd = pk.build_install_command(name, version, index, wheel_url, tool, no_upgrade, d.get("url_dependencies", []))
    if not cmd:
        raise ValueError("No installation method available")
    return cmd, base, version, name


def install_single_environment(env_slug: str, tool: str = "uv") -> bool:
    """Used by ``prime eval run`` / ``prime lab`` to make a hub environment importable; True on success."""
    try:
        cmd, base, version, _ = _resolve_hub_install(api(require_auth=False), env_slug, tool, no_upgrade=False)
        execute_install_command(cmd, base, version, tool)
        return True
    except Exception as e:
        console.print(f"[red]Failed to install {env_slug}: {e}[/red]")
        return False


@app.command(no_args_is_help=True, rich_help_panel="Manage")
@handle_errors
def install(
    env_ids: list[str] = typer.Argument(..., help="Environment ID(s): owner/name[@version] or a local name"),
    with_tool: str = typer.Option("uv", "--with", help="Package manager to use (uv or pip)"),
    path: str = typer.Option("./environments", "--path", "-p", help="Local environments directory (for local installs)"),
    no_upgrade: bool = typer.Option(False, "--no-upgrade", help="Don't upgrade existing packages (useful with uv.lock)"),
) -> None:
    """Install verifiers environments from the hub or from ./environments."""
    if with_tool not in ("uv", "pip"):
        raise fail(f"Unsupported package manager '{with_tool}'. Use 'uv' or 'pip'.")
    if not shutil.which(with_tool):
        raise fail(f"{with_tool} is not installed.")
    client = api(require_auth=False)
    plugin = load_verifiers_prime_plugin(console=console)
    env_ids = list(dict.fromkeys(env_ids))
    plural = lambda n: "s" if n != 1 else ""  # noqa: E731
    todo: list[tuple[list[str], str, str, str]] = []
    console.print(f"[bold]Resolving {len(env_ids)} environment{plural(len(env_ids))}...[/bold]")
    for slug in env_ids:
        local = slug.split("@")[0]
        if "/" not in local:
            if not local.strip():
                console.print("[yellow]⚠ Skipping: Empty environment name[/yellow]")
                continue
            folder = local.replace("-", "_")
            env_path = Path(path) / folder
            if env_path.exists():
                cmd = (plugin.build_module_command(plugin.install_module, [local, "--path", path]) if with_tool == "uv"
                       else ["pip", "install", "-e", str(env_path)])  # fmt: skip
                todo.append((cmd, local, "local", local))
                console.print(f"[green]✓ Found local environment: {env_path}[/green]")
            else:
                console.print(f"[red]✗ Local environment not found: {env_path}[/red]")
                if "-" in local and (Path(path) / local).exists():
                    console.print(f"[yellow]  Hint: found '{Path(path) / local}'; Python packages use underscores — rename it to '{folder}'[/yellow]")
            continue
        try:
            todo.append(_resolve_hub_install(client, slug, with_tool, no_upgrade))
            console.print(f"[green]✓ Found {todo[-1][1]}@{todo[-1][2]}[/green]")
        except ValueError as e:
            console.print(f"[yellow]⚠ Skipping {slug}: {e}[/yellow]")
            console.print("[dim]  Use 'prime env info' to see available options or 'pull' to download source.[/dim]")
        except (APIError, RuntimeError, httpx.HTTPError, OSError) as e:
            console.print(f"[red]✗ Failed to resolve {slug}: {e}[/red]")
    if not todo:
        raise fail("Unable to resolve installable environments")
    done, failed = [], []
    console.print(f"\n[bold]Installing {len(todo)} environment{plural(len(todo))}...[/bold]")
    for cmd, base, version, name in todo:
        try:
            execute_install_command(cmd, base, version, with_tool)
            done.append(f"{base}@{version}")
            console.print("\n[dim]Use in Python:[/dim]\n  from verifiers import load_environment")
            console.print(f"  env = load_environment('{name}')")
        except FileNotFoundError:
            failed.append((f"{base}@{version}", f"{cmd[0]} command not found"))
        except Exception as e:
            failed.append((f"{base}@{version}", str(e)))
    if done:
        console.print(f"\n[bold]Installed {len(done)} environment{plural(len(done))}:[/bold]")
        for d in done:
            console.print(f"[green]✓ {d}[/green]")
    if failed:
        console.print(f"\n[bold]Failed to install {len(failed)} environment{plural(len(failed))}:[/bold]")
        for d, why in failed:
            console.print(f"[red]✗ {d} - {why}[/red]")
        if not done:
            raise typer.Exit(1)


@app.command(no_args_is_help=True, rich_help_panel="Manage")
def uninstall(env_name: str = typer.Argument(..., help="Environment name to uninstall"),
              with_tool: str = typer.Option("uv", "--with", help="Package manager to use (uv or pip)")) -> None:  # fmt: skip
    """Uninstall a verifiers environment."""
    if with_tool not in ("uv", "pip"):
        raise fail(f"Unsupported package manager '{with_tool}'. Use 'uv' or 'pip'.")
    pkg = pk.normalize_package_name(env_name.split("/", 1)[-1])
    cmd = pk.uv_pip_command("uninstall", pkg) if with_tool == "uv" else ["pip", "uninstall", "-y", pkg]
    if not shutil.which(cmd[0]):
        raise fail(f"{cmd[0]} is not installed.")
    console.print(f"[cyan]Uninstalling {pkg} with {with_tool}...[/cyan]")
    r = subprocess.run(cmd, capture_output=True, text=True)
    out = (r.stdout or "") + (r.stderr or "")
    if r.returncode != 0:
        raise fail(f"Uninstall failed (exit {r.returncode}): {out.strip()}")
    if "not installed" in out.lower() or "no packages to uninstall" in out.lower():
        console.print(f"[yellow]{pkg} is not installed.[/yellow]")
    else:
        console.print(f"[green]✓ Successfully uninstalled {pkg}[/green]")


# ======================================================================================================= versions / delete
@version_app.command("list", no_args_is_help=True)
@handle_errors
def list_versions(env_id: str = typer.Argument(..., help="Environment ID (owner/name)"),
                  full_hashes: bool = typer.Option(False, "--full-hashes", help="Show full content hashes")) -> None:  # fmt: skip
    """List all versions of an environment (newest first)."""
    owner, name = parse_slug(env_id)
    console.print(f"Fetching versions for {env_id}...")
    data = _data(api(require_auth=False).get(f"/environmentshub/{owner}/{name}/versions"))
    versions = data if isinstance(data, list) else (data or {}).get("versions", [])
    if not versions:
        console.print("No versions found.")
        return
    rows = []
    for v in versions:
        created = v.get("created_at", "")
        try:
            created = iso_timestamp(created) if "T" in created else created
        except Exception:
            pass
        h, n = v.get("sha256", ""), v.get("size", 0)
        rows.append([v.get("version", "unknown"), created, h if full_hashes or v.get("version") is None else h[:8], f"{n} artifact{'s' if n != 1 else ''}"])
    latest = versions[0].get("version", "unknown")
    emit("table", None, f"Versions for {env_id}", [("Version", "cyan"), ("Created", "green"), ("Content Hash", "yellow"), ("Artifacts", "magenta")],
         rows, f"\n[dim]Latest version: {latest}[/dim]\n[dim]Install with: prime env install {env_id}@{latest}[/dim]")  # fmt: skip


def _confirm(message: str, force: bool) -> None:
    if force:
        return
    try:
        ok = typer.confirm(message)
    except typer.Abort:
        ok = False
    if not ok:
        console.print("Deletion cancelled.")
        raise typer.Exit()


@version_app.command("delete", no_args_is_help=True)
@handle_errors
def delete_version(env_id: str = typer.Argument(..., help="Environment ID (owner/name)"),
                   content_hash: str = typer.Argument(..., help="Content hash of the version to delete"),
                   force: bool = typer.Option(False, "--force", "-f", help="Skip confirmation")) -> None:  # fmt: skip
    """Delete one version (by content hash) from the hub."""
    if len(content_hash) < 8:
        console.print("[yellow]Use 'prime env version list' to see available content hashes[/yellow]")
        raise fail("Please provide a valid content hash (at least 8 characters)")
    owner, name = parse_slug(env_id)
    _confirm(f"Are you sure you want to permanently delete version with content hash '{content_hash}' from '{env_id}' on the environments hub?", force)
    console.print(f"Deleting version {content_hash} from {env_id}...")
    try:
        api().delete(f"/environmentshub/{owner}/{name}/@{content_hash}")
    except APIError as e:
        raise fail(f"Version with content hash '{content_hash}' not found in environment '{env_id}'" if getattr(e, "status_code", None) == 404
                   or "404" in str(e) else f"Failed to delete version: {e}")  # fmt: skip
    console.print(f"[green]✓ Version {content_hash} deleted successfully from {env_id}[/green]")


@app.command(no_args_is_help=True, rich_help_panel="Manage")
@handle_errors
def delete(env_id: str = typer.Argument(..., help="Environment ID to delete"),
           force: bool = typer.Option(False, "--force", "-f", help="Skip confirmation")) -> None:  # fmt: skip
    """Delete an entire environment and ALL its versions from the hub."""
    _confirm(f"Are you sure you want to permanently delete entire environment '{env_id}' and ALL its versions from the environments hub?", force)
    console.print(f"Deleting {env_id} from remote hub...")
    api().delete(f"/environmentshub/{env_id}")
    console.print(f"[green]✓ Environment {env_id} deleted successfully[/green]")


# ======================================================================================================= actions (CI)
@action_app.command("list")
@handle_errors
def actions_list(environment: str = typer.Argument(..., help="Environment slug (owner/name)"),
                 version_id: Optional[str] = typer.Option(None, "--version-id", "-v", help="Filter by version ID"),
                 num: int = typer.Option(20, "--num", "-n", help="Items per page"),
                 page: int = typer.Option(1, "--page", "-p", help="Page number"), output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List actions (CI jobs) for an environment."""
    validate_output_format(output, console)
    if num < 1 or page < 1:
        raise fail("--num and --page must be at least 1")
    owner, name = parse_slug(environment)
    params: dict[str, Any] = {"limit": num, "offset": (page - 1) * num}
    if version_id:
        params["version_id"] = version_id
    data = api().get(f"/environmentshub/{owner}/{name}/actions", params=params).get("data", {})
    actions, total = data.get("actions", []), data.get("total", 0)
    if not actions and output != "json":
        console.print("[yellow]No more results.[/yellow]" if page > 1 else "[yellow]No actions found for this environment.[/yellow]")
        return
    rows = []
    for a in actions:
        ver = a.get("version") or {}
        rows.append([a.get("id", ""), a.get("name") or a.get("job_type", ""), colorize(a.get("status", ""), ACTION_COLORS),
                     ver.get("semantic_version") or (ver.get("content_hash") or "")[:8], a.get("trigger", ""),
                     format_time_ago(a["created_at"]) if a.get("created_at") else ""])  # fmt: skip
    footer = (f"\n[yellow]Showing page {page} of results. Use --page {page + 1} to see more.[/yellow]" if total > page * num
              else f"\n[dim]Total: {total} action(s)[/dim]")  # fmt: skip
    emit(output, data, f"Actions for {owner}/{name}", [("ID", "cyan"), ("Name", "blue"), ("Status", "yellow"), ("Version", "dim"),
                                                       ("Trigger", "dim"), ("Created", "dim")], rows, footer)  # fmt: skip


def new_log_lines(previous: str, current: str) -> list[str]:
    """Lines of ``current`` not already shown: the longest suffix of ``previous`` that prefixes ``current`` is the overlap
    (the server returns a sliding tail window)."""
    old, new = previous.splitlines(), current.splitlines()
    if not old:
        return new
    overlap = 0
    for i in range(1, min(len(old), len(new)) + 1):
        if old[-i:] == new[:i]:
            overlap = i
    return new[overlap:]


@action_app.command("logs")
@handle_errors
def actions_logs(environment: str = typer.Argument(..., help="Environment slug (owner/name)"),
                 action_id: str = typer.Argument(..., help="Action/job ID"),
                 tail: int = typer.Option(1000, "--tail", "-n", help="Number of lines to show"),
                 follow: bool = typer.Option(False, "--follow", "-f", help="Follow log output")) -> None:  # fmt: skip
    """Get logs for a specific action."""
    owner, name = parse_slug(environment)
    client = api()
    fetch = lambda: strip_ansi((client.get(f"/environmentshub/{owner}/{name}/actions/{action_id}/logs",
                                           params={"tail_lines": tail}).get("data") or {}).get("logs") or "")  # noqa: E731  # fmt: skip
    if not follow:
        logs = fetch()
        console.print(logs, markup=False, highlight=False) if logs else console.print("[yellow]No logs available yet.[/yellow]")
        return
    console.print(f"[dim]Watching logs for action {action_id}... (Ctrl+C to stop)[/dim]\n")
    shown, errors = "", 0
    try:
        while True:
            try:
                logs = fetch()
                errors = 0
            except APIError as e:
                errors += 1
                if getattr(e, "status_code", None) == 429 or "429" in str(e):
                    if errors >= 3:
                        console.print("[yellow]Rate limited. Waiting 30s...[/yellow]")
                    time.sleep(30 if errors >= 3 else 10)
                    continue
                raise
            if logs != shown:
                for line in new_log_lines(shown, logs):
                    console.print(line, markup=False, highlight=False)
                shown = logs
            time.sleep(5)
    except KeyboardInterrupt:
        console.print("\n[dim]Stopped watching logs.[/dim]")


@action_app.command("retry")
@handle_errors
def actions_retry(environment: str = typer.Argument(..., help="Environment slug (owner/name)"),
                  action_id: Optional[str] = typer.Argument(None, help="Action ID to retry (default: the latest)"),
                  output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Retry an action (integration test) for an environment."""
    validate_output_format(output, console)
    owner, name = parse_slug(environment)
    data = api().post(f"/environmentshub/{owner}/{name}/actions/retry", json={"action_id": action_id} if action_id else {}).get("data", {})
    if output == "json":
        output_data_as_json(data, console)
        return
    if not data.get("success"):
        raise fail(f"Retry failed: {data.get('message', 'Unknown error')}")
    console.print("[green]Successfully triggered retry[/green]")
    console.print(f"[dim]Job ID: {data.get('job_id')}[/dim]\n[dim]Version: {data.get('version_id')}[/dim]")
    console.print(f"\n[dim]Use 'prime env action logs {environment} {data.get('job_id')}' to view logs[/dim]")


# the reference's private spellings of helpers that its white-box tests (and a few downstream scripts) import by name
_resolve_push_environment_path = _resolve_push_path
_resolve_pull_environment_path = _resolve_pull_path
_collect_archive_files = pk.collect_archive_files
_safe_tar_extract = pk.safe_tar_extract
_validate_path_component = pk.validate_path_component

from . import env_secrets as _env_secrets  # noqa: E402  (registers `secret` and `var` sub-apps on `app`)

_ = (json, _env_secrets)
