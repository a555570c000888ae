"""Serve HTTP from inside a sandbox and reach it from the outside through an exposed port."""

import time

import httpx

from prime_b200.platform.sandboxes import APIClient, CreateSandboxRequest, SandboxClient

client = SandboxClient(APIClient())
sb = client.create(CreateSandboxRequest(name="expose-demo", docker_image="python:3.11-slim", timeout_minutes=15))
try:
    client.wait_for_creation(sb.id)
    client.execute_command(sb.id, "mkdir -p /srv && echo 'served from the sandbox' > /srv/index.html")
    job = client.start_background_job(sb.id, "python -m http.server 8000 --directory /srv")
    exposed = client.expose(sb.id, 8000, name="demo-http")
    print("exposed:", exposed.url)
    for attempt in range(20):
        try:
            r = httpx.get(exposed.url, timeout=5)
            print(r.status_code, r.text.strip())
            break
        except httpx.HTTPError:
            time.sleep(1.5)
    print("ports:", [(p.port, p.url) for p in client.list_exposed_ports(sb.id).exposures])
    client.unexpose(sb.id, exposed.exposure_id)
    print("server job still running:", not client.get_background_job(sb.id, job).completed)
finally:
    client.delete(sb.id)
