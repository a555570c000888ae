"""Sandbox in five calls (sync client). Needs PRIME_API_KEY."""

from prime_b200.platform.sandboxes import APIClient, CreateSandboxRequest, SandboxClient

client = SandboxClient(APIClient())
sb = client.create(CreateSandboxRequest(name="quickstart", docker_image="python:3.11-slim", cpu_cores=1, memory_gb=2, timeout_minutes=15,
                                        labels=["example"], environment_vars={"GREETING": "hello from the sandbox"}))  # fmt: skip
print("created", sb.id, sb.status)
try:
    client.wait_for_creation(sb.id)
    for cmd in ("echo $GREETING", "python -c 'import platform; print(platform.platform())'", "df -h / | tail -1"):
        r = client.execute_command(sb.id, cmd, timeout=60)
        print(f"$ {cmd}\n{r.stdout.rstrip()}" + (f"\n[stderr] {r.stderr.rstrip()}" if r.stderr else "") + f"   (exit {r.exit_code})")
    print("--- container logs ---")
    print(client.get_logs(sb.id)[-500:])
finally:
    client.delete(sb.id)
    print("deleted", sb.id)
