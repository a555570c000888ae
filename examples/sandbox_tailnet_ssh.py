"""Put a sandbox on your tailnet and SSH into it: Tailscale in userspace-networking mode (a sandbox has no TUN device), the auth key
handed over as a sandbox SECRET (never on a command line or in the sandbox record), every step checked by exit code.

    export TS_AUTHKEY=tskey-auth-…          # an ephemeral, pre-approved key is the right kind
    python examples/sandbox_tailnet_ssh.py [--keep]

Without ``--keep`` the sandbox is deleted at the end (an ephemeral key then removes the node from the tailnet by itself).
"""

import argparse
import os
import sys
import time

from prime_b200.platform.sandboxes import APIClient, CommandTimeoutError, CreateSandboxRequest, SandboxClient

ap = argparse.ArgumentParser()
ap.add_argument("--keep", action="store_true", help="leave the sandbox running and print how to reach it")
ap.add_argument("--image", default="ubuntu:22.04")
args = ap.parse_args()
key = os.environ.get("TS_AUTHKEY")
if not key:
    sys.exit("TS_AUTHKEY is not set: create an auth key in the Tailscale admin console and export it")

client = SandboxClient(APIClient())
sb = client.create(CreateSandboxRequest(name="tailnet-ssh", docker_image=args.image, start_command="sleep infinity", cpu_cores=1, memory_gb=2,
                                        timeout_minutes=120, secrets={"TS_AUTHKEY": key}))  # fmt: skip


def sh(command: str, what: str, timeout: int = 300) -> str:
    r = client.execute_command(sb.id, command, timeout=timeout)
    if r.exit_code != 0:
        raise RuntimeError(f"{what} failed with exit code {r.exit_code}: {(r.stderr or r.stdout).strip()[-400:]}")
    return r.stdout.strip()


SOCK = "/tmp/tailscaled.sock"  # explicit socket: the same path for the daemon and every client call
TS = f"tailscale --socket={SOCK}"
keep = args.keep
try:
    client.wait_for_creation(sb.id, max_attempts=120)
    print("sandbox", sb.id, "is running; installing tailscale …")
    sh("command -v curl >/dev/null || (apt-get update -qq && apt-get install -y -qq curl ca-certificates)", "installing curl")
    sh("command -v tailscaled >/dev/null || curl -fsSL https://tailscale.com/install.sh | sh", "installing tailscale", timeout=600)
    # the daemon outlives the request that started it: a background job, not a foreground command
    daemon = client.start_background_job(sb.id, f"tailscaled --tun=userspace-networking --socks5-server=localhost:1055 --state=mem: --socket={SOCK}")
    for _ in range(30):  # wait for its socket instead of sleeping a fixed time
        if client.execute_command(sb.id, f"test -S {SOCK}", timeout=10).exit_code == 0:
            break
        st = client.get_background_job(sb.id, daemon)
        if st.completed:
            raise RuntimeError(f"tailscaled exited with {st.exit_code}: {(st.stderr or '').strip()[-400:]}")
        time.sleep(1)
    else:
        raise RuntimeError("tailscaled did not open its socket within 30 s")
    try:
        sh(TS + ' up --ssh --hostname "sandbox-${SANDBOX_ID:-prime}" --authkey="$TS_AUTHKEY"', "tailscale up", timeout=60)
    except CommandTimeoutError:
        sys.exit("tailscale up timed out: is the key valid, and is the node pre-approved?")
    user, ip = sh("whoami", "whoami"), sh(TS + " ip -4 | head -n1", "reading the tailnet address")
    print(f"on the tailnet as {ip}\nconnect:  ssh -o StrictHostKeyChecking=accept-new {user}@{ip}")
    if keep:
        print(f"left running (2 h timeout); delete with:  prime-b200 sandbox delete {sb.id}")
except BaseException:
    keep = False
    raise
finally:
    if not keep:
        client.delete(sb.id)
        print("sandbox deleted")
