"""``prime pods {list,status,create,terminate,history,connect|ssh}``
(reference: packages/prime/src/prime_cli/commands/pods.py:148-1143).

The interactive create flow (offer → provider → count → name → disk/vCPU/RAM → image → sharing → confirm) is a small
state object (`PodWizard`) whose every question can be pre-answered from the command line; the request body it
produces is built by one pure function (`build_pod_config`).
"""

from __future__ import annotations

import hashlib
import json
import os
import subprocess
import time
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional

import typer

from ..api.availability import AvailabilityClient, CountSpec, GPUAvailability
from ..api.pods import HistoryObj, Pod, PodsClient, PodStatus
from ..core import Config
from ..helper.short_id import generate_short_id
from ..utils.display import POD_STATUS_COLORS, colorize, validate_output_format
from ..utils.formatters import format_ip_display
from ..utils.json_help import json_output_help, list_json_help
from ..utils.plain import is_plain_mode
from ..utils.prompt import confirm_or_skip
from ..utils.time_utils import human_age, iso_timestamp, sort_by_created
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app, paginate_hint
from .teams import fetch_team_members

app = make_app("Manage GPU pods")
DEFAULT_IMAGE = "ubuntu_22_cuda_12"
_LIST_FIELDS = {"id": "str", "name": "str|null", "gpu": "str", "status": "str", "created_at": "str", "price_hr": "float|null", "team_id": "str|null"}


# ----------------------------------------------------------------------------------------------- formatting
def pod_row(p: Pod) -> dict[str, Any]:
    return {"id": p.id, "name": p.name, "gpu": f"{p.gpu_type} x{p.gpu_count}", "status": p.status, "created_at": iso_timestamp(p.created_at),
            "age": human_age(p.created_at), "price_hr": p.price_hr, "team_id": p.team_id, "provider": p.provider_type}  # fmt: skip


def status_row(s: PodStatus, p: Pod) -> dict[str, Any]:
    """Keys and value forms of ``prime pods status -o json`` as scripts know them (reference: commands/pods.py:74-99 — ``ssh`` and
    ``ip`` are display strings, the optional keys appear only when set), plus the structured extras of this CLI."""
    ssh = s.ssh_connection if isinstance(s.ssh_connection, list) else ([s.ssh_connection] if s.ssh_connection else [])
    status = s.status
    if p.status == "ACTIVE" and s.installation_progress is not None and s.installation_progress < 100 and not s.installation_failure:
        status = "INSTALLING"
    row: dict[str, Any] = {"id": p.id, "status": status, "name": p.name, "team_id": p.team_id, "provider": s.provider_type,
                           "gpu": f"{p.gpu_type} x{p.gpu_count}", "image": p.environment_type, "created_at": iso_timestamp(p.created_at),
                           "ip": format_ip_display(s.ip), "ssh": format_ip_display(s.ssh_connection)}  # fmt: skip
    if s.cost_per_hr:
        row["cost_per_hour"] = s.cost_per_hr
    if p.installation_status:
        row["installation_status"] = p.installation_status
    if s.installation_progress is not None:
        row["installation_progress"] = s.installation_progress
    if s.installation_failure:
        row["installation_error"] = s.installation_failure
    row.update(ssh_connections=[c for c in ssh if c], port_mappings=[m.model_dump() for m in (s.prime_port_mapping or [])],
               attached_resources=[r.model_dump() for r in (p.attached_resources or [])])
    return row


def _duration(created: Any, terminated: Any) -> str | None:
    """``42m`` below an hour, ``3.5h`` below a day, ``2.1d`` above."""
    from ..utils.time_utils import parse_dt

    if not created or not terminated:
        return None
    hours = (parse_dt(terminated) - parse_dt(created)).total_seconds() / 3600
    return f"{int(hours * 60)}m" if hours < 1 else (f"{hours:.1f}h" if hours < 24 else f"{hours / 24:.1f}d")


def history_row(h: HistoryObj) -> dict[str, Any]:
    return {"id": h.id, "name": h.name, "provider": h.provider_type, "gpu": f"{h.gpu_name} x{h.count}", "created_at": iso_timestamp(h.created_at),
            "terminated_at": iso_timestamp(h.terminated_at) if h.terminated_at else None, "duration": _duration(h.created_at, h.terminated_at),
            "price_per_hour": h.price_hr, "total_cost": h.total_billed_price, "team_id": h.team_id, "type": h.type}  # fmt: skip


# ----------------------------------------------------------------------------------------------- list / status / history
@app.command("list", epilog=list_json_help("pods", _LIST_FIELDS, {"total_count": "int", "offset": "int", "limit": "int"}))
@handle_errors
def list_(limit: int = typer.Option(100, help="Maximum number of pods"), offset: int = typer.Option(0, help="Pods to skip"),
          watch: bool = typer.Option(False, "--watch", "-w", help="Refresh when something changes"), output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List your pods (oldest first)."""
    validate_output_format(output, console)
    if watch and output == "json":
        raise fail("--watch mode is not compatible with --output=json")
    client = PodsClient(api())
    seen = None
    while True:
        page = client.list(offset=offset, limit=limit)
        digest = hashlib.md5(json.dumps([p.model_dump(mode="json") for p in page.data], sort_keys=True).encode()).hexdigest()
        if digest != seen:
            if watch and not is_plain_mode():
                os.system("cls" if os.name == "nt" else "clear")
            rows = [pod_row(p) for p in sort_by_created(page.data)]
            emit(output, {"pods": [{k: r[k] for k in _LIST_FIELDS} for r in rows], "total_count": page.total_count, "offset": offset, "limit": limit},
                 f"Compute Pods (Total: {page.total_count})", [("ID", "cyan"), ("Name", "blue"), ("GPU", "green"), "Status", "Age", ("$/hr", "magenta"), "Team"],
                 [[r["id"], r["name"] or "", r["gpu"], colorize(r["status"], POD_STATUS_COLORS), r["age"],
                   "" if r["price_hr"] is None else f"{r['price_hr']:.2f}", r["team_id"] or "Personal"] for r in rows],
                 paginate_hint(page.total_count, offset, limit, "pods") or "\n[blue]Use 'prime pods status <pod-id>' for details[/blue]")  # fmt: skip
            seen = digest
        if not watch:
            return
        time.sleep(5)


@app.command(no_args_is_help=True, epilog=json_output_help({"id": "str", "status": "str", "ssh": "str", "ip": "str", "ssh_connections": ["str"], "installation_progress": "int (when installing)"}))
@handle_errors
def status(pod_id: str = typer.Argument(..., help="Pod ID"), output: str = OUTPUT_OPT) -> None:
    """Detailed status of one pod (state, SSH endpoint, install progress, ports, attached disks)."""
    client = PodsClient(api())
    statuses = client.get_status([pod_id])
    if not statuses:
        raise fail(f"No status found for pod {pod_id}")
    row = status_row(statuses[0], client.get(pod_id))
    rows = [["Status", colorize(row["status"], POD_STATUS_COLORS)], ["Name", row["name"] or "N/A"], ["Team", row["team_id"] or "Personal"],
            ["Provider", row["provider"]], ["GPU", row["gpu"]], ["Image", row["image"] or "N/A"],
            ["Cost per Hour", f"${row['cost_per_hour']:.3f}" if "cost_per_hour" in row else "N/A"], ["Created", row["created_at"]],
            ["IP", row["ip"]], ["SSH", "\n".join(row["ssh_connections"]) or "N/A"]]  # fmt: skip
    if "installation_status" in row:
        rows.append(["Installation Status", row["installation_status"]])
    if "installation_progress" in row:
        rows.append(["Installation Progress", f"{row['installation_progress']}%"])
    if "installation_error" in row:
        rows.append(["Installation Error", f"[red]{row['installation_error']}[/red]"])
    for m in row["port_mappings"]:
        rows.append([f"Port {m.get('internal')}", f"{m.get('external')} ({m.get('protocol')}) {m.get('description') or ''}"])
    for r in row["attached_resources"]:
        rows.append([f"Disk {r.get('id')}", f"{r.get('size')} GB at {r.get('mount_path')} [{r.get('status')}]"])
    emit(output, row, f"Pod Status: {pod_id}", [("Property", "cyan"), ("Value", "green")], rows)


@app.command(epilog=list_json_help("history", {"id": "str", "name": "str", "gpu": "str", "created_at": "str", "terminated_at": "str|null", "duration": "str|null",
                                                "price_per_hour": "float", "total_cost": "float"}))
@handle_errors
def history(limit: int = typer.Option(100, help="Maximum number of entries"), offset: int = typer.Option(0), output: str = OUTPUT_OPT) -> None:
    """Terminated pods and what they cost."""
    page = PodsClient(api()).history(offset=offset, limit=limit)
    rows = [history_row(h) for h in page.data]
    emit(output, {"history": rows, "total_count": page.total_count, "offset": offset, "limit": limit}, f"Pods History (Total: {page.total_count})",
         [("ID", "cyan"), ("Name", "blue"), ("GPU", "green"), "Provider", "Created", "Terminated", "Duration", ("$/hr", "magenta"), ("Billed", "magenta")],
         [[r["id"], r["name"], r["gpu"], r["provider"], r["created_at"], r["terminated_at"] or "", r["duration"] or "", f"{r['price_per_hour']:.2f}",
           f"${r['total_cost']:.2f}"] for r in rows],
         paginate_hint(page.total_count, offset, limit, "entries"))  # fmt: skip


# ----------------------------------------------------------------------------------------------- create
def parse_env_pairs(pairs: list[str] | None) -> list[dict[str, str]]:
    out = []
    for item in pairs or []:
        key, sep, value = item.partition("=")
        if not sep or not key:
            raise ValueError(f"--env expects KEY=VALUE, got {item!r}")
        out.append({"key": key, "value": value})
    return out


def offers_supporting_env_vars(offers: dict[str, list[GPUAvailability]]) -> dict[str, list[GPUAvailability]]:
    """Env vars need an image with an init system: drop runpod and the bare ubuntu image."""
    kept: dict[str, list[GPUAvailability]] = {}
    for gtype, gpus in offers.items():
        ok = []
        for g in gpus:
            images = [i for i in (g.images or []) if i != DEFAULT_IMAGE]
            if g.provider != "runpod" and images:
                g.images = images
                ok.append(g)
        if ok:
            kept[gtype] = ok
    return kept


def valid_pod_name(name: str) -> bool:
    return bool(name) and any(c.isalpha() for c in name) and all(c.isalnum() or c == "-" for c in name)


def cheapest_per_provider(configs: list[GPUAvailability]) -> list[GPUAvailability]:
    seen, out = set(), []
    for g in sorted(configs, key=lambda x: x.prices.price):
        if (g.provider, g.is_spot) not in seen:
            seen.add((g.provider, g.is_spot))
            out.append(g)
    return out


def price_label(p: float) -> str:
    return "N/A" if p == float("inf") else f"${round(float(p), 2)}/hr"


def build_pod_config(gpu: GPUAvailability, *, name: str | None, cloud_id: str | None, disk_size: int | None, vcpus: int | None,
                     memory: int | None, image: str | None, custom_template_id: str | None, env_vars: list[dict[str, str]],
                     disks: list[str] | None, team_id: str | None, shared_with_team: bool, team_member_ids: list[str]) -> dict[str, Any]:  # fmt: skip
    cfg: dict[str, Any] = {
        "pod": {"name": name or None, "cloudId": cloud_id or gpu.cloud_id, "gpuType": gpu.gpu_type, "socket": gpu.socket,
                "gpuCount": gpu.gpu_count, "diskSize": disk_size, "vcpus": vcpus, "memory": memory, "image": image,
                "dataCenterId": gpu.data_center, "maxPrice": None, "country": None, "security": None, "jupyterPassword": None,
                "autoRestart": False, "customTemplateId": custom_template_id, "envVars": env_vars},
        "provider": {"type": gpu.provider} if gpu.provider else {},
        "disks": disks,
        "team": {"teamId": team_id} if team_id else None,
    }  # fmt: skip
    if shared_with_team:
        cfg["sharedWithTeam"] = True
    if team_member_ids:
        cfg["teamMemberIds"] = team_member_ids
    return cfg


@dataclass
class PodWizard:
    offers: dict[str, list[GPUAvailability]]
    ask: Callable[..., Any] = typer.prompt
    out: Any = field(default_factory=lambda: console)

    def _pick(self, title: str, labels: list[str], question: str) -> int:
        self.out.print(f"\n[bold]{title}[/bold]")
        for i, lab in enumerate(labels, 1):
            self.out.print(f"{i}. {lab}")
        n = self.ask(question, type=int, default=1, show_default=False)
        if not 1 <= n <= len(labels):
            raise fail(f"Invalid selection: {n}")
        return n - 1

    def find(self, short_id: str | None, cloud_id: str | None) -> GPUAvailability | None:
        for gpus in self.offers.values():
            for g in gpus:
                if (short_id and generate_short_id(g) == short_id) or (not short_id and cloud_id and g.cloud_id == cloud_id):
                    return g
        return None

    def choose_offer(self, gpu_type: str | None, gpu_count: int | None) -> GPUAvailability:
        if not gpu_type:
            types = sorted(t for t, g in self.offers.items() if g)
            if not types:
                raise fail("No GPU offers available")
            gpu_type = types[self._pick("Available GPU Types:", types, "Select GPU type number")]
        configs = self.offers.get(str(gpu_type), [])
        if not gpu_count:
            best: dict[int, float] = {}
            for g in configs:
                best[g.gpu_count] = min(best.get(g.gpu_count, float("inf")), g.prices.price)
            counts = sorted(best)
            if not counts:
                raise fail(f"No configuration found for {gpu_type}")
            gpu_count = counts[self._pick(f"Available {gpu_type} Configurations:", [f"{c}x {gpu_type} ({price_label(best[c])})" for c in counts],
                                          "Select configuration number")]  # fmt: skip
        matching = [g for g in configs if g.gpu_count == gpu_count]
        if not matching:
            raise fail(f"No configuration found for {gpu_count}x {gpu_type}")
        unique = cheapest_per_provider(matching)
        if len(unique) == 1:
            return unique[0]
        labels = [f"{g.provider}{' (spot)' if g.is_spot else ''} ({price_label(g.prices.price)})" for g in unique]
        return unique[self._pick("Available Providers:", labels, "Select provider number")]

    def choose_name(self, gpu: GPUAvailability) -> str:
        default = f"{gpu.gpu_type.lower().split('_')[0]}-{gpu.gpu_count}"
        while True:
            name = self.ask("Pod name (alphanumeric and dashes only, must contain at least 1 letter)", default=default)
            if valid_pod_name(name):
                return name
            self.out.print("[red]Invalid name format. Use only letters, numbers and dashes. Must contain at least 1 letter.[/red]")

    def choose_amount(self, spec: CountSpec, what: str, unit: str) -> int | None:
        if spec.min_count is None or spec.max_count is None or (what != "Disk size" and spec.default_count is None):
            return spec.default_count
        v = self.ask(f"{what}{unit} (min: {spec.min_count}, max: {spec.max_count})", default=spec.default_count or spec.min_count, type=int)
        if v is None or not spec.min_count <= v <= spec.max_count:
            raise fail(f"{what} must be between {spec.min_count} and {spec.max_count}")
        return v

    def choose_image(self, gpu: GPUAvailability) -> str | None:
        images = gpu.images or []
        if not images:
            return None
        return images[0] if len(images) == 1 else images[self._pick("Available Images:", images, "Select image number")]

    def choose_members(self, members: list[dict], me: str | None) -> tuple[bool, list[str], str | None]:
        """→ (share with whole team, explicit member ids, summary)."""
        others = [m for m in members if m.get("userId") != me]
        if not others:
            self.out.print("[yellow]No other team members to share with.[/yellow]")
            return False, [], None
        self.out.print("\n[bold]Team Members:[/bold]")
        for i, m in enumerate(others, 1):
            self.out.print(f"  {i}. {m.get('userName') or 'N/A'} ({m.get('userEmail') or 'N/A'}) - {m.get('role', '')}")
        answer = self.ask("\nSelect members (comma-separated numbers, or 'all' for everyone)", default="all")
        if answer.strip().lower() == "all":
            return True, [], "All team members"
        picked = []
        for part in (p.strip() for p in answer.split(",")):
            if not part.isdigit() or not 1 <= int(part) <= len(others):
                raise fail(f"Invalid selection: {part}. Must be between 1 and {len(others)}")
            picked.append(others[int(part) - 1])
        return False, [m["userId"] for m in picked], ", ".join(m.get("userName") or m["userId"] for m in picked)


@app.command()
@handle_errors
def create(
    id: Optional[str] = typer.Option(None, help="Short ID from 'prime availability list'"),
    cloud_id: Optional[str] = typer.Option(None, help="Cloud ID from the provider"),
    gpu_type: Optional[str] = typer.Option(None, help="GPU type (e.g. H100_80GB)"),
    gpu_count: Optional[int] = typer.Option(None, help="Number of GPUs"),
    name: Optional[str] = typer.Option(None, help="Pod name"),
    disk_size: Optional[int] = typer.Option(None, help="Disk size in GB"),
    vcpus: Optional[int] = typer.Option(None, help="Number of vCPUs"),
    memory: Optional[int] = typer.Option(None, help="Memory in GB"),
    image: Optional[str] = typer.Option(None, help="Image name, or 'custom_template' with --custom-template-id"),
    custom_template_id: Optional[str] = typer.Option(None, help="Custom template ID"),
    team_id: Optional[str] = typer.Option(None, help="Team ID (defaults to the configured team)"),
    disks: Optional[List[str]] = typer.Option(None, help="Disk IDs to attach (repeatable)"),
    env: Optional[List[str]] = typer.Option(None, help="KEY=VALUE environment variables (repeatable)"),
    share_with_team: bool = typer.Option(False, "--share-with-team", help="Share with every team member"),
    add_members: bool = typer.Option(False, "--add-members", help="Pick team members to share with"),
    yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation"),
) -> None:
    """Create a pod; anything not given on the command line is asked interactively."""
    cfg = Config(writable=False)
    team = team_id or cfg.team_id
    try:
        env_vars = parse_env_pairs(env)
    except ValueError as e:
        raise fail(str(e))
    if share_with_team and add_members:
        raise fail("--share-with-team and --add-members are mutually exclusive.")
    if (share_with_team or add_members) and not team:
        raise fail("--share-with-team and --add-members require a team. Use --team-id or 'prime switch <team>'.")
    if custom_template_id and image != "custom_template":
        raise fail("Must set image='custom_template' when using custom_template_id")
    if image == "custom_template" and not custom_template_id:
        raise fail("Must provide custom_template_id when image='custom_template'")

    client = api()
    with console.status("[bold blue]Loading available GPU configurations..."):
        offers = AvailabilityClient(client, on_error=console.print).get()
    if env_vars:
        offers = offers_supporting_env_vars(offers)
    wiz = PodWizard(offers)
    gpu = wiz.find(id, cloud_id) if (id or cloud_id) else wiz.choose_offer(gpu_type, gpu_count)
    if gpu is None:
        raise fail("No valid GPU configuration found")
    name = name or wiz.choose_name(gpu)
    disk_size = disk_size or wiz.choose_amount(gpu.disk, "Disk size", " in GB")
    vcpus = vcpus or wiz.choose_amount(gpu.vcpu, "Number of vCPUs", "")
    memory = memory or wiz.choose_amount(gpu.memory, "Memory", " in GB")
    image = image or wiz.choose_image(gpu)

    shared, member_ids, sharing = share_with_team, [], None
    if add_members and team:
        shared, member_ids, sharing = wiz.choose_members(fetch_team_members(client, team), cfg.user_id)
    elif not share_with_team and team and cfg.share_resources_with_team:
        shared, sharing = True, "All team members (from config default)"
    if shared and not sharing:
        sharing = "All team members"

    pod_config = build_pod_config(gpu, name=name, cloud_id=gpu.cloud_id if id else (cloud_id or gpu.cloud_id), disk_size=disk_size, vcpus=vcpus,
                                  memory=memory, image=image, custom_template_id=custom_template_id, env_vars=env_vars, disks=disks,
                                  team_id=team, shared_with_team=shared, team_member_ids=member_ids)  # fmt: skip
    console.print("\n[bold]Pod Configuration Summary:[/bold]")
    for k, v in pod_config["pod"].items():
        if v is not None:
            console.print(f"{k}: {v}")
    if pod_config["provider"].get("type"):
        console.print(f"provider: {pod_config['provider']['type']}")
    console.print(f"team: {team}")
    if disks:
        console.print(f"disks: {', '.join(disks)}")
    if sharing:
        console.print(f"sharing: {sharing}")
    if not confirm_or_skip("\nDo you want to create this pod?", yes, default=True):
        console.print("\nPod creation cancelled")
        raise typer.Exit(0)
    with console.status("[bold blue]Creating pod..."):
        pod = PodsClient(client).create(pod_config)
    console.print(f"\n[green]Successfully created pod {pod.id}[/green]\n\n[blue]Use 'prime pods status {pod.id}' to check the pod status[/blue]")


@app.command(no_args_is_help=True)
@handle_errors
def terminate(pod_id: str = typer.Argument(..., help="Pod ID"), yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:
    """Terminate a pod."""
    if not confirm_or_skip(f"Are you sure you want to terminate pod {pod_id}?", yes):
        console.print("Termination cancelled")
        raise typer.Exit(0)
    with console.status("[bold blue]Terminating pod..."):
        PodsClient(api()).delete(pod_id)
    console.print(f"[green]Successfully terminated pod {pod_id}[/green]")


# ----------------------------------------------------------------------------------------------- ssh
def split_ssh_target(connection: str) -> tuple[str, str]:
    """``"root@1.2.3.4 -p 2222"`` → ("root@1.2.3.4", "2222")."""
    host, sep, port = connection.partition(" -p ")
    return host.strip(), (port.strip() if sep else "22")


def ssh_command(key_path: str, connection: str) -> list[str]:
    host, port = split_ssh_target(connection)
    return ["ssh", "-i", key_path, "-o", "StrictHostKeyChecking=no", "-p", port, host]


def _connect(pod_id: str) -> None:
    client = PodsClient(api())
    with console.status("[bold blue]Waiting for SSH connection to become available..."):
        while True:
            statuses = client.get_status([pod_id])
            if not statuses:
                raise fail(f"No status found for pod {pod_id}")
            if statuses[0].ssh_connection:
                break
            time.sleep(5)
    key = Config(writable=False).ssh_key_path
    if not os.path.exists(key):
        raise fail(f"SSH key not found at {key}")
    console.print(f"[blue]Using SSH key:[/blue] {key}\n[dim]To change SSH key path, use: prime config set-ssh-key-path[/dim]")
    conn = statuses[0].ssh_connection
    options = [str(c) for c in conn if c] if isinstance(conn, list) else [str(conn)]
    if not options:
        raise fail("No valid SSH connections available")
    pick = 0
    if len(options) > 1:
        console.print("\nMultiple nodes available. Please select one:")
        for i, c in enumerate(options, 1):
            console.print(f"[blue]{i}[/blue]) {c}")
        pick = typer.prompt("Enter node number", type=int, default=1, show_default=False) - 1
        if not 0 <= pick < len(options):
            raise fail("Invalid selection")
    subprocess.run(ssh_command(key, options[pick]))


@app.command("connect", no_args_is_help=True)
@handle_errors
def connect(pod_id: str = typer.Argument(..., help="Pod ID")) -> None:
    """SSH into a pod with the configured key (waits until the endpoint exists)."""
    _connect(pod_id)


@app.command("ssh", no_args_is_help=True)
@handle_errors
def ssh(pod_id: str = typer.Argument(..., help="Pod ID")) -> None:
    """Alias of 'connect'."""
    _connect(pod_id)
