"""``prime images {push,list,delete}`` — tar the build context, PUT it to a signed URL, start a remote build
(reference: packages/prime/src/prime_cli/commands/images.py:32-457)."""

from __future__ import annotations

import tarfile
import tempfile
from pathlib import Path
from typing import Optional

import click
import httpx
import typer

from ..core import APIClient, APIError, Config, UnauthorizedError  # noqa: F401  (APIClient: the injection point `api()` honours)
from ..utils.display import colorize
from ..utils.json_help import list_json_help
from ._common import OUTPUT_OPT, api, console, emit, fail, handle_errors, make_app

app = make_app("Build and manage container images")
PACKAGED_DOCKERFILE_PATH = ".__prime_dockerfile__"  # wire contract with the build service
BUILD_COLORS = {"COMPLETED": "green", "BUILDING": "yellow", "PENDING": "yellow", "UPLOADING": "yellow", "FAILED": "red", "CANCELLED": "dim"}


def split_reference(ref: str, require_tag: bool = False) -> tuple[str, str]:
    if ":" in ref:
        name, tag = ref.rsplit(":", 1)
        return name, tag
    if require_tag:
        raise ValueError("Image reference must include a tag (e.g. myapp:latest)")
    return ref, "latest"


def parse_delete_reference(ref: str, default_team: str | None) -> tuple[str, str, str | None]:
    """``name:tag`` or ``team-<teamId>/name:tag`` → (name, tag, team_id)."""
    team = default_team
    if "/" in ref:
        ns, ref = ref.split("/", 1)
        if not ns.startswith("team-"):
            raise ValueError(f"Unrecognized image namespace '{ns}'. Use 'imagename:tag' or 'team-{{teamId}}/imagename:tag'.")
        team = ns[5:]
        if not team:
            raise ValueError("Invalid team image reference. Expected format: team-{teamId}/imagename:tag")
    name, tag = split_reference(ref, require_tag=True)
    return name, tag, team


def package_context(context: Path, dockerfile: Path, dest: Path) -> float:
    """tar.gz of the context with the Dockerfile copied to a fixed in-archive path; returns size in MB."""
    with tarfile.open(dest, "w:gz") as tar:
        tar.add(context, arcname=".")
        tar.add(dockerfile, arcname=PACKAGED_DOCKERFILE_PATH)
    return dest.stat().st_size / (1 << 20)


@app.command("push")
def push_image(
    image_reference: str = typer.Argument(..., help="e.g. 'myapp:v1.0.0' (tag defaults to latest)"),
    context: str = typer.Option(".", "--context", "-c", help="Build context directory"),
    dockerfile: Optional[str] = typer.Option(None, "--dockerfile", "-f", help="Dockerfile path", show_default="<context>/Dockerfile"),
    platform: str = typer.Option("linux/amd64", "--platform", click_type=click.Choice(["linux/amd64", "linux/arm64"])),
) -> None:
    """Build an image remotely and push it to the platform registry."""
    cfg = Config(writable=False)
    name, tag = split_reference(image_reference)
    if "/" in name:
        raise fail("Image name cannot contain '/'. Use simple names like 'myapp:v1.0.0'.")
    ctx = Path(context).resolve()
    df = Path(dockerfile).resolve() if dockerfile else ctx / "Dockerfile"
    if not ctx.is_dir():
        raise fail(f"Build context must be an existing directory: {ctx}")
    if not df.is_file():
        raise fail(f"Dockerfile not found at {df}")
    console.print(f"[bold blue]Building and pushing image:[/bold blue] {name}:{tag}" + (f"\n[dim]Team: {cfg.team_id}[/dim]" if cfg.team_id else ""))
    client = api()
    with tempfile.TemporaryDirectory() as tmp:
        tar_path = Path(tmp) / "context.tar.gz"
        console.print("[cyan]Preparing build context...[/cyan]")
        console.print(f"[green]✓[/green] Build context packaged ({package_context(ctx, df, tar_path):.2f} MB)")
        body = {"image_name": name, "image_tag": tag, "dockerfile_path": PACKAGED_DOCKERFILE_PATH, "platform": platform}
        if cfg.team_id:
            body["team_id"] = cfg.team_id
        try:
            build = client.request("POST", "/images/build", json=body)
        except UnauthorizedError:
            raise fail("Not authenticated. Please run 'prime login' first.")
        except APIError as e:
            raise fail(f"Failed to initiate build: {e}")
        build_id, upload_url = build.get("build_id"), build.get("upload_url")
        if not build_id or not upload_url:
            raise fail("Invalid response from server (missing build_id or upload_url)")
        console.print("[green]✓[/green] Build initiated\n[cyan]Uploading build context...[/cyan]")
        try:
            with open(tar_path, "rb") as f:
                httpx.put(upload_url, content=f, headers={"Content-Type": "application/octet-stream"}, timeout=600.0).raise_for_status()
        except httpx.HTTPError as e:
            raise fail(f"Upload failed: {e}")
        try:
            client.request("POST", f"/images/build/{build_id}/start", json={"context_uploaded": True})
        except APIError as e:
            raise fail(f"Failed to start build: {e}")
    full = build.get("fullImagePath") or f"{name}:{tag}"
    console.print(f"[green]✓[/green] Build context uploaded\n[green]✓[/green] Build started\n\n[bold]Build ID:[/bold] {build_id}\n[bold]Image:[/bold] {full}")
    console.print(f"\nCheck status with [bold]prime images list[/bold]; once complete: prime sandbox create {full}")


@app.command("list", epilog=list_json_help("images", {"imageName": "str", "imageTag": "str", "status": "str", "fullImagePath": "str", "createdAt": "str"}))
@handle_errors
def list_images(output: str = OUTPUT_OPT, all_images: bool = typer.Option(False, "--all", "-a", help="[Deprecated] no effect")) -> None:
    """List images and builds of the active account."""
    cfg = Config(writable=False)
    if all_images and output != "json":
        console.print("[yellow]Warning: --all is deprecated; images are scoped to the active account.[/yellow]\n")
    images = api().request("GET", "/images", params={"teamId": cfg.team_id} if cfg.team_id else None).get("data", [])
    if not images and output != "json":
        console.print("[yellow]No images or builds found.[/yellow]")
        return
    emit(output, {"images": images, "total_count": len(images)}, "Images",
         [("Image", "cyan"), ("Status", "white"), "Size", ("Created", "magenta")],
         [[i.get("fullImagePath") or f"{i.get('imageName')}:{i.get('imageTag')}", colorize(i.get("status"), BUILD_COLORS),
           i.get("sizeBytes") or "", i.get("createdAt") or ""] for i in images])  # fmt: skip


@app.command("delete")
@handle_errors
def delete_image(image_reference: str = typer.Argument(..., help="'myapp:v1' or 'team-{teamId}/myapp:v1'"),
                 yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation")) -> None:  # fmt: skip
    """Delete an image (creator or team admin only for team images)."""
    try:
        name, tag, team = parse_delete_reference(image_reference, Config(writable=False).team_id)
    except ValueError as e:
        raise fail(str(e))
    if not yes and not typer.confirm(f"Are you sure you want to delete {name}:{tag}" + (f" (team: {team})" if team else "") + "?"):
        console.print("[yellow]Cancelled[/yellow]")
        raise typer.Exit(0)
    api().request("DELETE", f"/images/{name}/{tag}", params={"teamId": team} if team else None)
    console.print(f"[green]✓ Deleted {name}:{tag}[/green]")
