"""``prime tunnel {start,list,status,stop}`` (reference: packages/prime/src/prime_cli/commands/tunnel.py:22-349)."""

from __future__ import annotations

import asyncio
import signal
from typing import Any, Callable, Coroutine, List, Optional

import typer

from ..tunnel import Tunnel, TunnelClient, TunnelConnectionError, TunnelLimitReachedError, TunnelTimeoutError
from ..utils.display import build_table
from ..utils.prompt import confirm_or_skip
from ._common import console, fail, make_app

app = make_app("Expose local ports through managed tunnels")
STATUS_MARKUP = {"CONNECTED": "[green]connected[/green]", "PENDING": "[yellow]pending[/yellow]",
                 "DISCONNECTED": "[red]disconnected[/red]", "EXPIRED": "[dim]expired[/dim]"}  # fmt: skip
HEALTH_POLL_S = 2.0


def with_client(fn: Callable[[TunnelClient], Coroutine[Any, Any, Any]]) -> Any:
    async def go():
        c = TunnelClient()
        try:
            return await fn(c)
        finally:
            await c.close()

    try:
        return asyncio.run(go())
    except typer.Exit:
        raise
    except Exception as e:
        raise fail(str(e))


def parse_ids(raw: list[str] | None) -> list[str]:
    """Space- and/or comma-separated ids → unique, order-preserving list."""
    out: list[str] = []
    for chunk in raw or []:
        for tid in chunk.split(","):
            tid = tid.strip()
            if tid and tid not in out:
                out.append(tid)
    return out


async def supervise(tunnel: Tunnel, port: int, stop: asyncio.Event) -> None:
    url = await tunnel.start()
    console.print(f"\n[green]Tunnel started successfully![/green]\n[bold]URL:[/bold] {url}\n[bold]Tunnel ID:[/bold] {tunnel.tunnel_id}")
    console.print(f"\n[dim]Forwarding to localhost:{port}[/dim]\n[dim]Press Ctrl+C to stop the tunnel[/dim]\n")
    while not stop.is_set():
        if not tunnel.is_running:  # frpc died: surface its last words
            out = "\n".join(tunnel.recent_output) or "(no output captured)"
            raise TunnelConnectionError(f"Tunnel process exited unexpectedly\n--- frpc output ---\n{out}", tunnel_id=tunnel.tunnel_id)
        try:
            await asyncio.wait_for(stop.wait(), timeout=HEALTH_POLL_S)
        except asyncio.TimeoutError:
            pass


@app.command("start")
def start_tunnel(port: int = typer.Option(8765, "--port", "-p", help="Local port to expose"),
                 name: Optional[str] = typer.Option(None, "--name", "-n", help="Friendly name"),
                 team_id: Optional[str] = typer.Option(None, "--team-id", help="Team ID (defaults to the configured team)")) -> None:  # fmt: skip
    """Start a tunnel and keep it alive until Ctrl+C."""

    async def main() -> None:
        tunnel, stop = Tunnel(local_port=port, name=name, team_id=team_id), asyncio.Event()

        def on_signal() -> None:
            console.print("\n[yellow]Shutting down tunnel...[/yellow]")
            stop.set()

        loop = asyncio.get_running_loop()
        for sig in (signal.SIGINT, signal.SIGTERM):
            try:
                loop.add_signal_handler(sig, on_signal)
            except NotImplementedError:  # Windows
                pass
        try:
            await supervise(tunnel, port, stop)
        except TunnelConnectionError as e:
            console.print(f"\n[red]Tunnel error:[/red] {e}" + (f"\n[dim]Tunnel ID: {e.tunnel_id}[/dim]" if e.tunnel_id else ""))
            raise typer.Exit(1)
        except TunnelLimitReachedError as e:
            console.print(f"\n[red]Tunnel limit reached:[/red] {e}\n[dim]Delete an existing tunnel before creating a new one.[/dim]")
            raise typer.Exit(1)
        except TunnelTimeoutError as e:
            console.print(f"\n[red]Connection timed out:[/red] {e}\n[dim]{e.hint}[/dim]")
            raise typer.Exit(1)
        except Exception as e:
            hint = getattr(e, "hint", None)
            console.print(f"[red]Error:[/red] {e}" + (f"\n[dim]{hint}[/dim]" if hint else ""))
            raise typer.Exit(1)
        finally:
            await tunnel.stop()
            console.print("[green]Tunnel stopped[/green]")

    try:
        asyncio.run(main())
    except KeyboardInterrupt:
        pass


@app.command("list")
def list_tunnels(team_id: Optional[str] = typer.Option(None, "--team-id", help="Include this team's tunnels")) -> None:
    """List active tunnels."""
    tunnels = with_client(lambda c: c.list_tunnels(team_id=team_id))
    if not tunnels:
        console.print("[dim]No active tunnels[/dim]")
        return
    console.print(build_table("Active Tunnels", [("Tunnel ID", "cyan"), ("User ID", "magenta"), ("URL", "green"), "Status", "Expires At"],
                              [[t.tunnel_id, t.user_id or "", t.url, STATUS_MARKUP.get(t.status or "", t.status or "unknown"), t.expires_at] for t in tunnels]))  # fmt: skip


@app.command("status")
def tunnel_status(tunnel_id: str = typer.Argument(..., help="Tunnel ID")) -> None:
    """Show one tunnel."""
    t = with_client(lambda c: c.get_tunnel(tunnel_id))
    if not t:
        console.print(f"[red]Tunnel not found:[/red] {tunnel_id}")
        raise typer.Exit(1)
    for label, v in (("Tunnel ID", t.tunnel_id), ("URL", t.url), ("Hostname", t.hostname), ("Status", t.status or "unknown"), ("Expires At", t.expires_at)):
        console.print(f"[bold]{label}:[/bold] {v}")


@app.command("stop")
def stop_tunnel(
    tunnel_ids: Optional[List[str]] = typer.Argument(None, help="Tunnel ID(s), space or comma separated"),
    all: bool = typer.Option(False, "--all", "-a", help="Stop all tunnels"),
    team_id: Optional[str] = typer.Option(None, "--team-id", help="Team whose tunnels --all should include"),
    yes: bool = typer.Option(False, "--yes", "-y", help="Skip confirmation"),
    only_mine: bool = typer.Option(True, "--only-mine/--all-users", "-m/-A", help="With --all: only tunnels you own", show_default=True),
) -> None:
    """Stop and delete tunnels."""
    if all and tunnel_ids:
        raise fail("Cannot specify tunnel IDs with --all")
    if not all and not tunnel_ids:
        raise fail("Must specify at least one tunnel ID or --all")
    if all:

        async def mine(c: TunnelClient) -> list[str]:
            ts = await c.list_tunnels(team_id=team_id)
            if only_mine:
                if not c.config.user_id:
                    raise fail("Cannot filter by user - no user_id configured. Use --all-users, or run 'prime whoami' first.")
                ts = [t for t in ts if t.user_id == c.config.user_id]
            return [t.tunnel_id for t in ts]

        ids = with_client(mine)
        if not ids:
            console.print("[yellow]No active tunnels to stop[/yellow]" + ("\n\n[dim]Note: --all only touches your own tunnels; add --all-users for the whole team.[/dim]" if only_mine else ""))
            return
        question, cancelled = f"Are you sure you want to stop ALL {len(ids)} tunnel(s)? This action cannot be undone.", "Stop all cancelled"
    else:
        ids = parse_ids(tunnel_ids)
        if not ids:
            raise fail("No valid tunnel IDs provided")
        question = f"Are you sure you want to stop tunnel {ids[0]}?" if len(ids) == 1 else f"Are you sure you want to stop {len(ids)} tunnel(s)?"
        cancelled = "Stop cancelled" if len(ids) == 1 else "Bulk stop cancelled"
    if not confirm_or_skip(question, yes):
        console.print(cancelled)
        return

    async def delete(c: TunnelClient):
        # one id or many: the bulk endpoint (`DELETE /tunnel` with the id list), as the reference's `tunnel stop` does
        r = await c.bulk_delete_tunnels(ids)
        return r.get("succeeded", []), r.get("not_found", []), r.get("failed", [])

    ok, missing, failed = with_client(delete)
    for tid in ok:
        console.print(f"[green]✓ Stopped {tid}[/green]")
    for m in missing:
        console.print(f"[yellow]Not found: {m.get('tunnel_id', m) if isinstance(m, dict) else m}[/yellow]")
    for f in failed:
        console.print(f"[red]Failed: {f}[/red]")
    if failed:
        raise typer.Exit(1)
