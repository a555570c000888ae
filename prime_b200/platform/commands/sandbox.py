"""``prime sandbox {list|ls,get,create,delete,logs,run,upload,download,reset-cache,expose,unexpose,list-ports,ssh}``
(reference: packages/prime/src/prime_cli/commands/sandbox.py:146-1426)."""

from __future__ import annotations

import os
import random
import shutil
import string
import subprocess
import tempfile
import time
from typing import Any, List, Optional

import httpx
import typer

from ..core import Config
from ..sandboxes import CommandTimeoutError, CreateSandboxRequest, Sandbox, SandboxClient, SandboxNotRunningError
from ..utils.display import SANDBOX_STATUS_COLORS, colorize, output_data_as_json
from ..utils.formatters import format_resources, obfuscate_env_vars
from ..utils.json_help import json_output_help, list_json_help
from ..utils.prompt import confirm_or_skip
from ..utils.time_utils import human_age, iso_timestamp, sort_by_created
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app

app = make_app("Remote code-execution sandboxes")
BULK_BATCH = 100
_LIST_FIELDS = {"id": "str", "name": "str", "image": "str", "status": "str", "resources": "str", "labels": ["str"], "created_at": "str"}


def client() -> SandboxClient:
    return SandboxClient(api())


def resources_of(s: Sandbox) -> str:
    base = f"{s.cpu_cores:g} CPU, {s.memory_gb:g}GB RAM, {s.disk_size_gb:g}GB disk"
    return base + (f", {s.gpu_type} x{s.gpu_count}" if s.gpu_count else "")


def list_row(s: Sandbox) -> dict[str, Any]:
    return {"id": s.id, "name": s.name, "image": s.docker_image, "status": s.status,
            "resources": format_resources(s.cpu_cores, s.memory_gb, s.gpu_count), "labels": s.labels,
            "created_at": iso_timestamp(s.created_at), "age": human_age(s.created_at), "type": "VM" if s.vm else "Container",
            "user_id": s.user_id, "team_id": s.team_id}  # fmt: skip


def detail_row(s: Sandbox) -> dict[str, Any]:
    d = list_row(s)
    # the raw fields under the names `prime sandbox get -o json` has always printed them (reference: commands/sandbox.py, get)
    d.update(docker_image=s.docker_image, cpu_cores=s.cpu_cores, memory_gb=s.memory_gb, disk_size_gb=s.disk_size_gb, disk_mount_path=s.disk_mount_path,
             gpu_count=s.gpu_count, gpu_type=s.gpu_type, vm=s.vm, resources_detail=resources_of(s))
    d.update(start_command=s.start_command, network_access=s.network_access, timeout_minutes=s.timeout_minutes,
             environment_vars=obfuscate_env_vars(s.environment_vars), secrets=sorted((s.secrets or {}).keys()),
             started_at=iso_timestamp(s.started_at) if s.started_at else None,
             terminated_at=iso_timestamp(s.terminated_at) if s.terminated_at else None, exit_code=s.exit_code,
             error_type=s.error_type, error_message=s.error_message, registry_credentials_id=s.registry_credentials_id)  # fmt: skip
    return d


def guard_vm_unsupported(s: Sandbox, feature: str) -> None:
    if s.vm:
        raise fail(f"{feature} is not yet supported for VM sandboxes.")


def parse_pairs(items: list[str] | None, what: str) -> dict[str, str]:
    out: dict[str, str] = {}
    for item in items or []:
        key, sep, value = item.partition("=")
        if not sep:
            raise fail(f"{what} must be in KEY=VALUE format")
        out[key] = value
    return out


def slug(text: str) -> str:
    return "-".join(filter(None, "".join(c if c.isalnum() else "-" for c in text.lower()).split("-")))


def auto_name(docker_image: str, vm: bool, gpu_count: int, gpu_type: str | None, rng=random) -> str:
    if gpu_count > 0 and gpu_type:
        base = f"gpu-{slug(gpu_type)}"
    else:
        image = slug(docker_image.split("/")[-1].split(":")[0])
        base = f"vm-{image}" if vm else image
    return f"{base}-{''.join(rng.choices(string.ascii_lowercase + string.digits, k=4))}"


def parse_ids(raw: list[str] | None) -> list[str]:
    out: list[str] = []
    for chunk in raw or []:
        for sid in chunk.split(","):
            sid = sid.strip()
            if sid and sid not in out:
                out.append(sid)
    return out


# --------------------------------------------------------------------------------------------------- list / get
def _list(team_id, status, labels, page, num, all, output) -> None:
    if num < 1 or page < 1:
        raise fail("--num and --page must be at least 1")
    # an explicit --status already says which sandboxes are wanted: only the unfiltered default view hides terminated ones
    resp = client().list(team_id=team_id, status=status, labels=labels, page=page, per_page=num, exclude_terminated=(not all and status is None))
    rows = [list_row(s) for s in sort_by_created(resp.sandboxes)]
    footer = f"\n[yellow]More results available. Use --page {page + 1}[/yellow]" if resp.has_next else None
    emit(output, {"sandboxes": [{k: r[k] for k in _LIST_FIELDS} for r in rows], "total": resp.total, "page": resp.page, "per_page": num,
                  "has_next": resp.has_next}, f"Code Sandboxes (Total: {resp.total})",
         [("ID", "cyan"), ("Name", "blue"), ("Image", "green"), "Status", "Type", "Resources", "Labels", "Age"],
         [[r["id"], r["name"], r["image"], colorize(r["status"], SANDBOX_STATUS_COLORS), r["type"], r["resources"], ", ".join(r["labels"]), r["age"]] for r in rows],
         footer)  # fmt: skip


_TEAM = typer.Option(None, help="Team ID (defaults to the configured team)")
_STATUS = typer.Option(None, help="Filter by status")
_LABELS = typer.Option(None, "--label", "-l", help="Filter by label (repeatable; all must match)")
_PAGE = typer.Option(1, "--page", "-p", help="Page number")
_NUM = typer.Option(50, "--num", "-n", help="Items per page")
_ALL = typer.Option(False, "--all", help="Include terminated sandboxes")


@app.command("list", epilog=list_json_help("sandboxes", _LIST_FIELDS, {"total": "int", "page": "int", "has_next": "bool"}))
@handle_errors
def list_sandboxes_cmd(team_id: Optional[str] = _TEAM, status: Optional[str] = _STATUS, labels: Optional[List[str]] = _LABELS, page: int = _PAGE,
                       num: int = _NUM, all: bool = _ALL, output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List sandboxes."""
    _list(team_id, status, labels, page, num, all, output)


@app.command("ls", hidden=True)
@handle_errors
def ls(team_id: Optional[str] = _TEAM, status: Optional[str] = _STATUS, labels: Optional[List[str]] = _LABELS, page: int = _PAGE,
       num: int = _NUM, all: bool = _ALL, output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Alias of 'list'."""
    _list(team_id, status, labels, page, num, all, output)


@app.command(no_args_is_help=True, epilog=json_output_help({**_LIST_FIELDS, "exit_code": "int|null", "error_type": "str|null"}))
@handle_errors
def get(sandbox_id: str = typer.Argument(...), output: str = OUTPUT_OPT) -> None:
    """Show one sandbox."""
    d = detail_row(client().get(sandbox_id))
    rows = [[k.replace("_", " ").title(), colorize(v, SANDBOX_STATUS_COLORS) if k == "status" else ("" if v is None else v)] for k, v in d.items()]
    emit(output, d, f"Sandbox {sandbox_id}", [("Property", "cyan"), ("Value", "green")], rows)


# --------------------------------------------------------------------------------------------------- create / delete
@app.command(no_args_is_help=True)
@handle_errors
def create(
    docker_image: Optional[str] = typer.Argument(None, help="Docker image, e.g. python:3.12-slim"),
    name: Optional[str] = typer.Option(None, help="Sandbox name (generated when omitted)"),
    start_command: Optional[str] = typer.Option("tail -f /dev/null", help="Command that keeps the container alive"),
    cpu_cores: float = typer.Option(1.0, help="CPU cores"),
    memory_gb: float = typer.Option(2.0, help="Memory in GB"),
    disk_size_gb: float = typer.Option(10.0, help="Disk size in GB"),
    gpu_count: int = typer.Option(0, help="Number of GPUs (needs --vm and --gpu-type)"),
    gpu_type: Optional[str] = typer.Option(None, help="GPU type, e.g. H100_80GB"),
    vm: bool = typer.Option(False, "--vm", help="Run as a VM instead of a container"),
    network_access: bool = typer.Option(True, "--network-access/--no-network-access", help="Outbound network access"),
    timeout_minutes: int = typer.Option(60, help="Lifetime in minutes"),
    team_id: Optional[str] = typer.Option(None, help="Team ID (defaults to the configured team)"),
    registry_credentials_id: Optional[str] = typer.Option(None, help="Credentials for a private image"),
    env: Optional[List[str]] = typer.Option(None, "--env", "-e", help="KEY=VALUE environment variable (repeatable)"),
    secret: Optional[List[str]] = typer.Option(None, "--secret", help="KEY=VALUE secret (repeatable; never echoed)"),
    labels: Optional[List[str]] = typer.Option(None, "--label", "-l", help="Label (repeatable)"),
    yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation"),
) -> None:
    """Create a sandbox."""
    env_vars, secrets = parse_pairs(env, "Environment variables"), parse_pairs(secret, "Secrets")
    if gpu_count > 0 and not gpu_type:
        raise fail("GPU type is required when requesting GPUs. Provide --gpu-type with --gpu-count > 0.")
    if gpu_count > 0 and not vm:
        raise fail("GPUs require VM sandboxes. Pass --vm whenever using --gpu-count.")
    if gpu_count == 0 and gpu_type:
        raise fail("GPU type provided without GPUs. Set --gpu-count > 0 when using --gpu-type.")
    if not docker_image:
        raise fail("Docker image is required. Provide a DOCKER_IMAGE positional argument.")
    name = name or auto_name(docker_image, vm, gpu_count, gpu_type)
    req = CreateSandboxRequest(name=name, docker_image=docker_image, start_command=start_command, cpu_cores=cpu_cores, memory_gb=memory_gb,
                               disk_size_gb=disk_size_gb, gpu_count=gpu_count, gpu_type=gpu_type, vm=vm, network_access=network_access,
                               timeout_minutes=timeout_minutes, environment_vars=env_vars or None, secrets=secrets or None,
                               labels=labels or [], team_id=team_id, registry_credentials_id=registry_credentials_id)  # fmt: skip
    console.print("\n[bold]Sandbox Configuration:[/bold]")
    lines = [("Name", name), ("Docker Image", docker_image), ("Start Command", start_command or "N/A"),
             ("Resources", f"{cpu_cores} CPU, {memory_gb}GB RAM, {disk_size_gb}GB disk"), ("VM", "Enabled" if vm else "Disabled"),
             ("GPUs", f"{gpu_type} x{gpu_count}" if gpu_count else None),
             ("Network Access", "[green]Enabled[/green]" if network_access else "[yellow]Disabled[/yellow]"),
             ("Timeout", f"{timeout_minutes} minutes"), ("Team", team_id or Config(writable=False).team_id or "Personal"),
             ("Registry Credentials", registry_credentials_id), ("Labels", ", ".join(labels) if labels else None),
             ("Environment Variables", obfuscate_env_vars(env_vars) if env_vars else None),
             ("Secrets", {k: "***" for k in secrets} if secrets else None)]  # fmt: skip
    for label, v in lines:
        if v is not None:
            console.print(f"{label}: {v}")
    if not confirm_or_skip("\nDo you want to create this sandbox?", yes, default=True):
        console.print("\nSandbox creation cancelled")
        return
    with console.status("[bold blue]Creating sandbox..."):
        sb = client().create(req)
    console.print(f"\n[green]Successfully created sandbox {sb.id}[/green]\n[blue]Use 'prime sandbox get {sb.id}' to check the sandbox status[/blue]")


@app.command(no_args_is_help=True)
@handle_errors
def delete(
    sandbox_ids: Optional[List[str]] = typer.Argument(None, help="Sandbox ID(s), space or comma separated"),
    all: bool = typer.Option(False, "--all", "-a", help="Delete all sandboxes"),
    labels: Optional[List[str]] = typer.Option(None, "--label", "-l", help="Delete every sandbox carrying these labels"),
    yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation"),
    only_mine: bool = typer.Option(True, "--only-mine/--all-users", "-m/-A", help="With --all: only sandboxes you own", show_default=True),
) -> None:
    """Delete sandboxes by id, by label, or all of them."""
    if sum(map(bool, (all, sandbox_ids, labels))) > 1:
        raise fail("Cannot specify more than one of: sandbox IDs, --all, or --label")
    if not (all or sandbox_ids or labels):
        raise fail("Must specify either sandbox IDs, --all flag, or --label")
    c = client()
    if labels:
        if not confirm_or_skip(f"Are you sure you want to delete ALL sandboxes with labels: {', '.join(labels)}? This action cannot be undone.", yes):
            console.print("Delete cancelled")
            return
        with console.status("[bold blue]Deleting sandboxes by labels..."):
            r = c.bulk_delete(labels=labels)
        console.print(f"\n[green]{r.message}[/green]")
        for sid in r.succeeded:
            console.print(f"  ✓ {sid}")
        return
    if all:
        me = Config(writable=False).user_id
        if only_mine and not me:
            raise fail("Cannot filter by user - no user_id configured. Use --all-users, or run 'prime whoami' first.")
        with console.status("[bold blue]Fetching all sandboxes..."):
            found, page = [], 1
            while True:
                resp = c.list(per_page=100, page=page, exclude_terminated=True)
                found += resp.sandboxes
                if not resp.has_next:
                    break
                page += 1
        ids = [s.id for s in found if not only_mine or s.user_id == me]
        if not ids:
            console.print("[yellow]No sandboxes to delete[/yellow]" + ("\n\n[dim]Note: --all only deletes your own sandboxes; use --all-users for the whole team.[/dim]" if only_mine and found else ""))
            return
    else:
        ids = parse_ids(sandbox_ids)
    if len(ids) == 1 and not all:
        if not confirm_or_skip(f"Are you sure you want to delete sandbox {ids[0]}?", yes):
            console.print("Delete cancelled")
            return
        with console.status("[bold blue]Deleting sandbox..."):
            c.delete(ids[0])
        console.print(f"[green]Successfully deleted sandbox {ids[0]}[/green]")
        return
    question = f"Are you sure you want to delete ALL {len(ids)} sandbox(es)? This action cannot be undone." if all else f"Are you sure you want to delete {len(ids)} sandbox(es)?"
    if not confirm_or_skip(question, yes):
        console.print("Delete all cancelled" if all else "Bulk delete cancelled")
        return
    ok, bad = [], []
    batches = [ids[i : i + BULK_BATCH] for i in range(0, len(ids), BULK_BATCH)]
    for n, batch in enumerate(batches, 1):
        console.print(f"[dim]Processing batch {n}/{len(batches)} ({len(batch)} sandboxes)...[/dim]")
        r = c.bulk_delete(sandbox_ids=batch)
        ok += r.succeeded
        bad += r.failed
    console.print(f"\n[green]Processed {len(ok) + len(bad)} sandbox(es)[/green]")
    for sid in ok:
        console.print(f"  ✓ {sid}")
    for f in bad:
        console.print(f"  ✗ {f.get('sandbox_id', 'unknown')}: {f.get('error', 'unknown error')}")
    if bad:
        raise typer.Exit(1)


# --------------------------------------------------------------------------------------------------- logs / run / files
@app.command(no_args_is_help=True)
@handle_errors
def logs(sandbox_id: str = typer.Argument(...)) -> None:
    """Print the sandbox's container logs."""
    with console.status("[bold blue]Fetching logs..."):
        text = client().get_logs(sandbox_id)
    console.print(text, markup=False) if text else console.print("[yellow]No logs available[/yellow]")


@app.command(no_args_is_help=True)
@handle_errors
def run(
    sandbox_id: str = typer.Argument(...),
    command: List[str] = typer.Argument(..., help="Command to run (quote it or put it after --)"),
    working_dir: Optional[str] = typer.Option(None, "--working-dir", "-w", help="Working directory"),
    env: Optional[List[str]] = typer.Option(None, "--env", "-e", help="KEY=VALUE environment variable (repeatable)"),
    timeout: Optional[int] = typer.Option(None, "--timeout", "-t", help="Timeout in seconds (default 300)"),
) -> None:
    """Execute a command in a sandbox; exits with the command's exit code."""
    cmd = " ".join(command)
    t0 = time.time()
    try:
        with console.status(f"[bold blue]Running: {cmd[:60]}"):
            r = client().execute_command(sandbox_id, cmd, working_dir=working_dir, env=parse_pairs(env, "Environment variables") or None, timeout=timeout)
    except (CommandTimeoutError, SandboxNotRunningError) as e:
        raise fail(str(e))
    if r.stdout:
        console.print(r.stdout.rstrip("\n"), markup=False, highlight=False)
    if r.stderr:
        console.print(f"[red]{'stderr:'}[/red]")
        console.print(r.stderr.rstrip("\n"), markup=False, highlight=False)
    console.print(f"[dim]exit code {r.exit_code} in {time.time() - t0:.2f}s[/dim]")
    if r.exit_code != 0:
        raise typer.Exit(r.exit_code)


@app.command("upload", no_args_is_help=True)
@handle_errors
def upload_file(sandbox_id: str = typer.Argument(...), local_file: str = typer.Argument(..., help="Local file"),
                remote_path: str = typer.Argument(..., help="Destination path inside the sandbox")) -> None:  # fmt: skip
    """Upload a file."""
    if not os.path.isfile(local_file):
        raise fail(f"Local file not found: {local_file}")
    with console.status("[bold blue]Uploading..."):
        r = client().upload_file(sandbox_id, remote_path, local_file)
    console.print(f"[green]✓ Uploaded {local_file} → {r.path} ({r.size} bytes)[/green]")


@app.command("download", no_args_is_help=True)
@handle_errors
def download_file(sandbox_id: str = typer.Argument(...), remote_path: str = typer.Argument(..., help="File inside the sandbox"),
                  local_file: str = typer.Argument(..., help="Where to save it")) -> None:  # fmt: skip
    """Download a file."""
    with console.status("[bold blue]Downloading..."):
        client().download_file(sandbox_id, remote_path, local_file)
    console.print(f"[green]✓ Downloaded {remote_path} → {local_file} ({os.path.getsize(local_file)} bytes)[/green]")


@app.command("reset-cache")
@handle_errors
def reset_cache(yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:
    """Forget cached per-sandbox gateway tokens (they are re-issued on demand)."""
    if not confirm_or_skip("Clear the local sandbox auth-token cache?", yes, default=True):
        return
    client().clear_auth_cache()
    console.print("[green]✓ Sandbox auth cache cleared[/green]")


# --------------------------------------------------------------------------------------------------- ports
_PORT_FIELDS = {"exposure_id": "str", "sandbox_id": "str", "port": "int", "name": "str|null", "url": "str", "protocol": "str|null", "external_endpoint": "str|null"}


@app.command("expose", no_args_is_help=True, epilog=json_output_help(_PORT_FIELDS))
@handle_errors
def expose_port(sandbox_id: str = typer.Argument(...), port: int = typer.Argument(..., help="Port inside the sandbox"),
                name: Optional[str] = typer.Option(None, help="Friendly name"),
                protocol: str = typer.Option("HTTP", "--protocol", "-p", help="HTTP (public URL) or TCP (host:port endpoint)"), output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """Expose a sandbox port to the internet."""
    proto = protocol.upper()
    if proto not in ("HTTP", "TCP"):
        raise fail("--protocol must be HTTP or TCP")
    c = client()
    guard_vm_unsupported(c.get(sandbox_id), "Port exposure")
    e = c.expose(sandbox_id, port, name=name, protocol=proto)
    if output == "json":
        return output_data_as_json(e.model_dump(), console)
    console.print(f"[green]✓ Exposed port {port}[/green]\n[bold]Exposure ID:[/bold] {e.exposure_id}")
    console.print(f"[bold]Endpoint:[/bold] {e.external_endpoint}" if proto == "TCP" and e.external_endpoint else f"[bold]URL:[/bold] {e.url}")


@app.command("unexpose", no_args_is_help=True)
@handle_errors
def unexpose_port(sandbox_id: str = typer.Argument(...), exposure_id: str = typer.Argument(...),
                  yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:  # fmt: skip
    """Remove a port exposure."""
    if not confirm_or_skip(f"Remove exposure {exposure_id} from sandbox {sandbox_id}?", yes):
        console.print("Cancelled")
        return
    c = client()
    guard_vm_unsupported(c.get(sandbox_id), "Port unexpose")
    c.unexpose(sandbox_id, exposure_id)
    console.print(f"[green]✓ Removed exposure {exposure_id}[/green]")


@app.command("list-ports", epilog=list_json_help("exposures", _PORT_FIELDS))
@handle_errors
def list_ports(sandbox_id: Optional[str] = typer.Argument(None, help="Sandbox ID (all sandboxes when omitted)"), output: str = OUTPUT_OPT) -> None:
    """List exposed ports."""
    c = client()
    if sandbox_id:
        guard_vm_unsupported(c.get(sandbox_id), "Port listing")
    exps = (c.list_exposed_ports(sandbox_id) if sandbox_id else c.list_all_exposed_ports()).exposures
    emit(output, {"exposures": [e.model_dump() for e in exps], "total_count": len(exps)}, "Exposed Ports",
         [("Exposure ID", "cyan"), "Sandbox", "Port", "Name", "Protocol", ("URL / endpoint", "green")],
         [[e.exposure_id, e.sandbox_id, e.port, e.name or "", e.protocol or "HTTP", e.external_endpoint or e.url] for e in exps])  # fmt: skip


# --------------------------------------------------------------------------------------------------- ssh
def sandbox_ssh_command(session_id: str, host: str, port: int, key_path: str | None, shell: str | None, extra: list[str] | None) -> list[str]:
    cmd = ["ssh", f"{session_id}@{host}", "-p", str(port), "-o", "StrictHostKeyChecking=no", "-o", "UserKnownHostsFile=/dev/null", "-o", "LogLevel=ERROR"]
    if key_path:
        cmd += ["-i", key_path]
    if shell:
        cmd.append("-t")  # a remote command needs an explicit PTY
    cmd += extra or []
    if shell:
        cmd.append(shell)
    return cmd


@app.command("ssh", no_args_is_help=True)
def ssh_connect(sandbox_id: str = typer.Argument(...), ssh_args: Optional[List[str]] = typer.Argument(None, help="Extra ssh args after --"),
                shell: Optional[str] = typer.Option(None, "--shell", "-s", help="Shell to start (bash, zsh, sh …)")) -> None:  # fmt: skip
    """Open an interactive shell: ephemeral ed25519 key → SSH session → authorize key at the gateway → ssh → clean up."""
    for tool in ("ssh", "ssh-keygen"):
        if not shutil.which(tool):
            raise fail(f"{tool} not found. Please install OpenSSH.")
    c = client()
    session_id, tmp = None, tempfile.mkdtemp(prefix="prime-ssh-")

    def cleanup() -> None:
        if session_id:
            try:
                c.close_ssh_session(sandbox_id, session_id)
                console.print("[green]✓[/green] SSH session closed")
            except Exception:
                pass
        shutil.rmtree(tmp, ignore_errors=True)

    try:
        sb = c.get(sandbox_id)
        guard_vm_unsupported(sb, "SSH")
        if sb.status != "RUNNING":
            raise fail(f"Sandbox is not running (status: {sb.status}). Check with: prime sandbox get {sandbox_id}")
        key = os.path.join(tmp, "id_ed25519")
        subprocess.run(["ssh-keygen", "-t", "ed25519", "-N", "", "-f", key], check=True, capture_output=True)
        with console.status("[bold blue]Setting up SSH session..."):
            sess = c.create_ssh_session(sandbox_id)
        session_id = sess.session_id
        url = f"{sess.gateway_url.rstrip('/')}/{sess.user_ns}/{sess.job_id}/authorize"
        body = {"session_id": sess.session_id, "public_key": open(key + ".pub").read().strip(), "ttl_seconds": sess.ttl_seconds}
        try:
            with httpx.Client(timeout=30) as h:
                h.post(url, json=body, headers={"Authorization": f"Bearer {sess.token}"}).raise_for_status()
        except Exception as e:
            raise fail(f"Failed to authorize SSH key: {e}")
        console.print(f"[green]✓[/green] SSH session ready: {sess.session_id}@{sess.host} port {sess.port}")
        with console.status("[bold blue]Waiting for connection to be ready..."):
            time.sleep(5)  # the TCP exposure needs a moment to propagate through the load balancer
        console.print("[dim]Press Ctrl+D or type 'exit' to disconnect[/dim]\n")
        rc = subprocess.run(sandbox_ssh_command(sess.session_id, sess.host, sess.port, key, shell, ssh_args)).returncode
        if rc not in (0, 255):
            console.print(f"\n[yellow]SSH connection exited with code {rc}[/yellow]")
    except KeyboardInterrupt:
        console.print("\n[yellow]SSH connection interrupted[/yellow]")
        raise typer.Exit(130)
    finally:
        cleanup()
