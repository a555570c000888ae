"""``prime disks {list,get,create,update,terminate}`` (reference: packages/prime/src/prime_cli/commands/disks.py:107-509)."""

from __future__ import annotations

import hashlib
import json
import os
import time
from typing import Any, Optional

import typer

from ..api.availability import AvailabilityClient
from ..api.disks import Disk, DisksClient
from ..helper.short_id import generate_short_id_disk
from ..utils.display import POD_STATUS_COLORS, colorize, validate_output_format
from ..utils.json_help import json_output_help, list_json_help
from ..utils.plain import is_plain_mode
from ..utils.prompt import confirm_or_skip
from ..utils.time_utils import human_age, iso_timestamp, sort_by_created
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app, paginate_hint

app = make_app("Manage persistent disks")
_LIST_FIELDS = {"id": "str", "name": "str", "size": "int (GB)", "status": "str", "provider": "str", "location": "str", "created_at": "str", "price_hr": "float|null"}


def disk_row(d: Disk) -> dict[str, Any]:
    info = d.info or {}
    # "US (dc-1)": country and data centre, "N/A" for whichever is unknown (the form scripts see in `disks list -o json`)
    location = f"{info.get('country', 'N/A')} ({info.get('dataCenterId', 'N/A')})"
    return {"id": d.id, "name": d.name, "size": d.size, "status": d.status, "provider": d.provider_type, "location": location,
            "created_at": iso_timestamp(d.created_at), "age": human_age(d.created_at), "price_hr": d.price_hr, "pods": d.pods, "clusters": d.clusters}  # fmt: skip


def disk_detail(d: Disk) -> dict[str, Any]:
    """`disks get -o json`: the list row plus the raw record's timestamps, prices, owners and provider info."""
    r = disk_row(d)
    r.update(updated_at=iso_timestamp(d.updated_at), terminated_at=iso_timestamp(d.terminated_at) if d.terminated_at else None,
             stopped_price_hr=d.stopped_price_hr, user_id=d.user_id, team_id=d.team_id, wallet_id=d.wallet_id, info=d.info)
    return r


def build_disk_config(size: int, name: str | None, team_id: str | None, *, offer=None, country: str | None = None,
                      cloud_id: str | None = None, data_center_id: str | None = None, provider_type: str | None = None) -> dict[str, Any]:  # fmt: skip
    if offer is not None:
        disk = {"size": size, "name": name, "country": offer.country, "cloudId": offer.cloud_id, "dataCenterId": offer.data_center}
        provider = {"type": offer.provider}
    else:
        disk = {"size": size, "name": name, "country": country, "cloudId": cloud_id, "dataCenterId": data_center_id}
        provider = {"type": provider_type} if provider_type else {}
    return {"disk": disk, "provider": provider, "team": {"teamId": team_id} if team_id else None}


@app.command("list", epilog=list_json_help("disks", _LIST_FIELDS, {"total_count": "int", "offset": "int", "limit": "int"}))
@handle_errors
def list_(limit: int = typer.Option(100, help="Maximum number of disks"), offset: int = typer.Option(0, help="Disks to skip"),
          watch: bool = typer.Option(False, "--watch", "-w", help="Refresh when something changes"), output: str = OUTPUT_OPT) -> None:  # fmt: skip
    """List your persistent disks (oldest first)."""
    validate_output_format(output, console)
    if watch and output == "json":
        raise fail("--watch mode is not compatible with --output=json")
    client = DisksClient(api())
    seen = None
    while True:
        page = client.list(offset=offset, limit=limit)
        digest = hashlib.md5(json.dumps([d.model_dump(mode="json") for d in page.data], sort_keys=True).encode()).hexdigest()
        if digest != seen:
            if watch and not is_plain_mode():
                os.system("cls" if os.name == "nt" else "clear")
            rows = [disk_row(d) for d in sort_by_created(page.data)]
            emit(output, {"disks": [{k: r[k] for k in _LIST_FIELDS} for r in rows], "total_count": page.total_count, "offset": offset, "limit": limit},
                 f"Disks (Total: {page.total_count})", [("ID", "cyan"), ("Name", "blue"), "Size (GB)", "Status", "Provider", ("Location", "green"), "Age", ("$/hr", "magenta")],
                 [[r["id"], r["name"], r["size"], colorize(r["status"], POD_STATUS_COLORS), r["provider"], r["location"], r["age"],
                   "" if r["price_hr"] is None else f"{r['price_hr']:.4f}"] for r in rows], paginate_hint(page.total_count, offset, limit, "disks"))  # fmt: skip
            seen = digest
        if not watch:
            return
        time.sleep(5)


@app.command(no_args_is_help=True, epilog=json_output_help({**_LIST_FIELDS, "pods": ["str"], "clusters": ["str"]}))
@handle_errors
def get(disk_id: str = typer.Argument(..., help="Disk ID"), output: str = OUTPUT_OPT) -> None:
    """Show one disk."""
    r = disk_detail(DisksClient(api()).get(disk_id))
    emit(output, r, f"Disk {r['id']}", [("Field", "cyan"), ("Value", "green")], [[k, v] for k, v in r.items()])


@app.command()
@handle_errors
def create(
    id: Optional[str] = typer.Option(None, help="Short ID from 'prime availability disks'"),
    size: int = typer.Option(..., help="Size in GB"),
    name: Optional[str] = typer.Option(None, help="Disk name"),
    country: Optional[str] = typer.Option(None), cloud_id: Optional[str] = typer.Option(None), data_center_id: Optional[str] = typer.Option(None),
    team_id: Optional[str] = typer.Option(None, help="Team ID (defaults to the configured team)"),
    provider_type: Optional[str] = typer.Option(None, help="Provider (e.g. lambda, runpod)"),
    yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation"),
) -> None:  # fmt: skip
    """Create a disk, either from an availability ID or from explicit location fields."""
    if size <= 0:
        raise fail("Disk size must be greater than 0")
    client = api()
    offer = None
    if id:
        with console.status("[bold blue]Loading available disks..."):
            offer = next((d for d in AvailabilityClient(client).get_disks() if generate_short_id_disk(d) == id), None)
        if offer is None:
            raise fail(f"No disk offer with ID {id}; see 'prime availability disks'")
    cfg = build_disk_config(size, name, team_id, offer=offer, country=country, cloud_id=cloud_id, data_center_id=data_center_id, provider_type=provider_type)
    if not cfg["provider"]:
        raise fail("Invalid disk configuration: give --id or --provider-type with location fields")
    console.print("\n[bold]Disk Configuration Summary:[/bold]")
    for label, v in (("Size", f"{size}GB"), ("Name", name), ("Country", cfg["disk"].get("country")), ("Cloud ID", cfg["disk"].get("cloudId")),
                     ("Data Center ID", cfg["disk"].get("dataCenterId")), ("Provider", cfg["provider"].get("type")), ("Team", team_id)):  # fmt: skip
        if v:
            console.print(f"{label}: {v}")
    if not confirm_or_skip("\nDo you want to create this disk?", yes, default=True):
        console.print("\nDisk creation cancelled")
        raise typer.Exit(0)
    with console.status("[bold blue]Creating disk..."):
        disk = DisksClient(client).create(cfg)
    console.print(f"\n[green]Successfully created disk {disk.id}[/green]\n\n[blue]Use 'prime disks get {disk.id}' to check the disk status[/blue]")


@app.command(no_args_is_help=True)
@handle_errors
def update(disk_id: str = typer.Argument(...), name: str = typer.Option(..., "--name", "-n", help="New name")) -> None:
    """Rename a disk."""
    DisksClient(api()).update(disk_id, name)
    console.print(f"[green]Disk {disk_id} renamed to '{name}'[/green]")


@app.command(no_args_is_help=True)
@handle_errors
def terminate(disk_id: str = typer.Argument(...), yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:
    """Delete a disk (data is lost)."""
    if not confirm_or_skip(f"Are you sure you want to terminate disk {disk_id}? All data will be lost.", yes):
        console.print("Termination cancelled")
        raise typer.Exit(0)
    r = DisksClient(api()).delete(disk_id)
    console.print(f"[green]Disk {disk_id}: {r.status}[/green]")
