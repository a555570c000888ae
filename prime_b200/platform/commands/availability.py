"""``prime availability {gpu-types,list,disks}`` (reference: packages/prime/src/prime_cli/commands/availability.py:81-416)."""

from __future__ import annotations

from typing import Any, List, Optional

import typer

from ..api.availability import AvailabilityClient, GPUAvailability
from ..helper.short_id import generate_short_id, generate_short_id_disk
from ..utils.display import status_color
from ..utils.json_help import list_json_help
from ..utils.plain import get_console
from ._common import OUTPUT_OPT, api, emit, handle_errors, make_app

_err = get_console(stderr=True)


def _warn(message: str) -> None:
    """Partial-failure notes (one endpoint of several down) go to stderr so that ``--output json`` stays parseable."""
    _err.print(message)


app = make_app("Check GPU and disk availability")
STOCK_COLORS = {"AVAILABLE": "green", "HIGH": "green", "MEDIUM": "yellow", "LOW": "yellow", "UNAVAILABLE": "red"}
_GPU_FIELDS = {"id": "str", "cloud_id": "str", "gpu_type": "str", "gpu_count": "int", "socket": "str", "provider": "str",
               "location": "str", "stock_status": "str", "price_per_hour": "str", "price_value": "float|null", "security": "str",
               "vcpus": "str", "memory_gb": "str", "disk_gb": "str", "gpu_memory": "int", "is_spot": "bool|null"}  # fmt: skip


def offer_row(g: GPUAvailability, gpu_type: str) -> dict[str, Any]:
    price = g.prices.price
    disk = str(g.disk.default_count)
    if g.disk.max_count is not None and g.disk.max_count != g.disk.default_count:
        disk += "+"
    # display forms scripts already parse: "B200 180GB" (underscores → spaces), security "community" | "datacenter"
    return {"id": generate_short_id(g), "cloud_id": g.cloud_id, "gpu_type": gpu_type.replace("_", " "), "gpu_count": g.gpu_count,
            "socket": g.socket or "N/A", "provider": g.provider or "N/A", "location": g.country or "N/A",
            "stock_status": g.stock_status, "price_per_hour": f"${price:.2f}" if price != float("inf") else "N/A",
            "price_value": None if price == float("inf") else price, "security": "community" if g.security == "community_cloud" else "datacenter",
            "vcpus": str(g.vcpu.default_count), "memory_gb": str(g.memory.default_count), "disk_gb": disk,
            "gpu_memory": g.gpu_memory, "is_spot": g.is_spot}  # fmt: skip


def _sort_key(r: dict[str, Any]):
    return (float("inf") if r["price_value"] is None else r["price_value"], r["id"])


def dedupe(rows: list[dict[str, Any]]) -> list[dict[str, Any]]:
    seen, out = set(), []
    for r in sorted(rows, key=_sort_key):
        if r["id"] not in seen:
            seen.add(r["id"])
            out.append(r)
    return out


def group_similar_rows(rows: list[dict[str, Any]]) -> list[dict[str, Any]]:
    """Collapse offers identical in everything a buyer cares about; show vCPU/RAM as ranges."""
    groups: dict[tuple, list[dict[str, Any]]] = {}
    for r in sorted(rows, key=_sort_key):
        key = tuple(r[k] for k in ("provider", "gpu_type", "gpu_count", "socket", "location", "security", "price_per_hour"))
        groups.setdefault(key, []).append(r)
    out = []
    for members in groups.values():
        rep = dict(members[0])
        for field in ("vcpus", "memory_gb"):
            vals = sorted({int(m[field]) for m in members if str(m[field]).isdigit()})
            if len(vals) > 1:
                rep[field] = f"{vals[0]}-{vals[-1]}"
        out.append(rep)
    return out


@app.command("gpu-types", epilog=list_json_help("gpu_types", {"gpu_type": "str"}))
@handle_errors
def gpu_types(output: str = OUTPUT_OPT) -> None:
    """List GPU types that currently have offers."""
    types = sorted(AvailabilityClient(api()).get_available_gpu_types())
    emit(output, {"gpu_types": [{"gpu_type": t} for t in types]}, "Available GPU Types", [("GPU Type", "cyan")], [[t] for t in types])


@app.command("list", epilog=list_json_help("gpu_resources", _GPU_FIELDS, {"total_count": "int", "filters": "object"}))
@handle_errors
def list_(
    gpu_type: Optional[str] = typer.Option(None, help="GPU type (e.g. H100_80GB, B200_180GB)"),
    gpu_count: Optional[int] = typer.Option(None, help="Number of GPUs required"),
    regions: Optional[List[str]] = typer.Option(None, help="Regions (repeatable or comma separated), e.g. united_states,eu_west"),
    socket: Optional[str] = typer.Option(None, help="Socket: PCIe, SXM4, SXM5, SXM6 …"),
    provider: Optional[str] = typer.Option(None, help="Only this provider"),
    disks: Optional[List[str]] = typer.Option(None, help="Only offers that can attach these disk ids"),
    group_similar: bool = typer.Option(True, "--group-similar/--no-group-similar", help="Collapse near-identical offers"),
    output: str = OUTPUT_OPT,
) -> None:
    """List available GPU offers, cheapest first."""
    data = AvailabilityClient(api(), on_error=_warn).get(gpu_type=gpu_type, gpu_count=gpu_count, regions=regions, disks=disks)
    rows = [offer_row(g, t) for t, gs in data.items() for g in gs
            if (not provider or g.provider == provider) and (not socket or g.socket == socket)]  # fmt: skip
    rows = group_similar_rows(rows) if group_similar else dedupe(rows)
    payload = {"gpu_resources": rows, "total_count": len(rows),
               "filters": {"gpu_type": gpu_type, "gpu_count": gpu_count, "regions": regions, "socket": socket,
                           "provider": provider, "group_similar": group_similar}}  # fmt: skip
    cols = [("ID", "cyan"), ("GPU Type", "cyan"), ("GPUs", "cyan"), ("Socket", "blue"), ("Provider", "blue"), ("Location", "green"),
            ("Stock", "yellow"), ("Price/Hr", "magenta"), ("Security", "white"), ("vCPUs", "blue"), ("RAM (GB)", "blue"), ("Disk (GB)", "blue")]  # fmt: skip
    table_rows = [[r["id"], r["gpu_type"], r["gpu_count"], r["socket"], r["provider"], r["location"],
                   f"[{status_color(r['stock_status'], STOCK_COLORS)}]{r['stock_status']}[/]", r["price_per_hour"], r["security"],
                   r["vcpus"], r["memory_gb"], r["disk_gb"]] for r in rows]  # fmt: skip
    emit(output, payload, "Available GPU Resources", cols, table_rows,
         "\n[bold blue]To deploy a pod with one of these configurations:[/bold blue]\n  [green]prime pods create --id <ID>[/green]   (interactive setup follows)")  # fmt: skip


@app.command("disks", epilog=list_json_help("disks", {"id": "str", "provider": "str", "location": "str", "price_per_gb_month": "str"}))
@handle_errors
def disks(
    regions: Optional[List[str]] = typer.Option(None, help="Regions (repeatable or comma separated)"),
    data_center_id: Optional[str] = typer.Option(None, help="Only this data center"),
    output: str = OUTPUT_OPT,
) -> None:
    """List persistent-disk offers."""
    offers = AvailabilityClient(api(), on_error=_warn).get_disks(regions=regions, data_center_id=data_center_id)
    rows = [{"id": generate_short_id_disk(d), "cloud_id": d.cloud_id, "provider": d.provider or "N/A", "data_center": d.data_center or "N/A",
             "location": d.country or d.region or "N/A", "stock_status": d.stock_status or "N/A",
             "price_per_gb_month": f"${d.spec.price_per_unit:.4f}" if d.spec.price_per_unit is not None else "N/A",
             "min_gb": d.spec.min_count, "max_gb": d.spec.max_count, "is_multinode": d.is_multinode} for d in offers]  # fmt: skip
    emit(output, {"disks": rows, "total_count": len(rows), "filters": {"regions": regions, "data_center_id": data_center_id}}, "Available Disks",
         [("ID", "cyan"), "Provider", "Data center", ("Location", "green"), ("Stock", "yellow"), ("Price/GB", "magenta"), "Min Size (GB)", "Max Size (GB)", "Is Multinode"],
         [[r["id"], r["provider"], r["data_center"], r["location"], r["stock_status"], r["price_per_gb_month"], r["min_gb"], r["max_gb"],
           r["is_multinode"]] for r in rows], "\n[bold blue]To create a disk with one of these configurations:[/bold blue]\n  [green]prime disks create --id <ID> --size <GB>[/green]")  # fmt: skip
