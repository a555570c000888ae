"""``prime login`` — browser challenge flow with an ephemeral RSA-2048 key pair: the server encrypts the new API
key with our public key (OAEP/SHA-256), we poll every 5 s and decrypt locally, so the key never travels in clear
(reference: packages/prime/src/prime_cli/commands/login.py:88-246)."""

from __future__ import annotations

import base64
import time
import webbrowser
from typing import Callable

import httpx
import typer
from cryptography.hazmat.primitives import hashes, serialization
from cryptography.hazmat.primitives.asymmetric import padding, rsa

from ..core import APIClient, Config
from ._common import console, make_app
from .teams import fetch_teams

app = make_app("Log in through the browser", invoke_without_command=True)
POLL_S = 5.0


def new_keypair() -> tuple[rsa.RSAPrivateKey, str]:
    key = rsa.generate_private_key(public_exponent=65537, key_size=2048)
    pem = key.public_key().public_bytes(serialization.Encoding.PEM, serialization.PublicFormat.SubjectPublicKeyInfo)
    return key, pem.decode()


def decrypt_result(key: rsa.RSAPrivateKey, blob: bytes) -> bytes:
    oaep = padding.OAEP(mgf=padding.MGF1(algorithm=hashes.SHA256()), algorithm=hashes.SHA256(), label=None)
    return key.decrypt(blob, oaep)


def run_challenge(base_url: str, frontend_url: str, *, http=httpx, announce: Callable[[str, str], None], sleep=time.sleep,
                  max_polls: int | None = None) -> str | None:  # fmt: skip
    """Returns the decrypted API key, or None if the challenge expired."""
    key, pem = new_keypair()
    try:
        r = http.post(f"{base_url}/api/v1/auth_challenge/generate", json={"encryptionPublicKey": pem})
        if r.status_code != 200:
            try:
                detail = r.json().get("detail", "Unknown error")
            except Exception:
                detail = r.text
            raise RuntimeError(f"Failed to generate challenge: {detail}")
        ch = r.json()
        announce(f"{frontend_url}/dashboard/tokens/challenge?code={ch['challenge']}", ch["challenge"])
        headers = {"Authorization": f"Bearer {ch['status_auth_token']}"}
        polls = 0
        while max_polls is None or polls < max_polls:
            polls += 1
            try:
                s = http.get(f"{base_url}/api/v1/auth_challenge/status", params={"challenge": ch["challenge"]}, headers=headers)
            except httpx.RequestError:
                console.print("[red]Failed to connect to server. Retrying...[/red]")
                sleep(POLL_S)
                continue
            if s.status_code == 404:
                return None
            result = s.json().get("result")
            if result:
                return decrypt_result(key, base64.b64decode(result)).decode()
            sleep(POLL_S)
        return None
    finally:
        del key  # the private key only ever lives in this frame


def pick_team(client: APIClient, config: Config) -> None:
    """After login: personal account, or choose one of the user's teams."""
    try:
        teams = fetch_teams(client)
        if not teams:
            config.set_team(None)
        else:
            console.print("\n[bold]Select account:[/bold]\n  [cyan](1)[/cyan] Personal")
            for i, t in enumerate(teams, 2):
                console.print(f"  [cyan]({i})[/cyan] {t.get('name', 'Unknown')} [dim]({t.get('role', 'member')})[/dim]")
            n = typer.prompt("Select", type=int, default=1)
            if 2 <= n <= len(teams) + 1:
                t = teams[n - 2]
                config.set_team(t.get("teamId"), team_name=t.get("name"), team_role=t.get("role"))
                console.print(f"[green]Using team '{t.get('name')}'.[/green]")
            else:
                config.set_team(None)
    except (typer.Abort, Exception):
        config.set_team(None)
    config.update_current_environment_file()


@app.callback(invoke_without_command=True)
def login(headless: bool = typer.Option(False, "--headless", help="Do not try to open a browser")) -> None:
    """Log in to Prime Intellect."""
    config = Config()
    if not config.base_url:
        console.print("Base URL not configured. Run 'prime config set-base-url' first.")
        raise typer.Exit(1)

    def announce(url: str, code: str) -> None:
        console.print("\n[bold blue]Login required[/bold blue]\n\n[bold]To authenticate:[/bold]\n")
        console.print(f"[bold yellow]1.[/bold yellow] Open this link in your browser:\n[link={url}]{url}[/link]")
        console.print(f"[bold yellow]2.[/bold yellow] The code should be pre-filled. Code:\n\n[bold green]{code}[/bold green]\n")
        console.print("[dim]Waiting for authentication...[/dim]")
        if not headless:
            try:
                webbrowser.open(url, new=2)
            except Exception:
                pass

    try:
        api_key = run_challenge(config.base_url, config.frontend_url, announce=announce)
    except KeyboardInterrupt:
        console.print("\n[yellow]Login cancelled by user[/yellow]")
        raise typer.Exit(1)
    except Exception as e:
        console.print(f"[red]{e}[/red]")
        raise typer.Exit(1)
    if api_key is None:
        console.print("[red]Challenge expired[/red]")
        raise typer.Exit(1)
    config.set_api_key(api_key)
    config.update_current_environment_file()
    client = APIClient(api_key=api_key, config=config)
    try:
        data = client.get("/user/whoami").get("data")
        if isinstance(data, dict) and data.get("id"):
            config.set_user_id(data["id"])
            config.update_current_environment_file()
    except Exception:
        console.print("[yellow]Logged in, but failed to fetch user id[/yellow]")
    console.print("[green]Successfully logged in![/green]")
    pick_team(client, config)
