"""Jobs that outlive a single HTTP request: started with nohup inside the sandbox, polled through an exit-code marker file."""

import time

from prime_b200.platform.sandboxes import APIClient, CreateSandboxRequest, SandboxClient

client = SandboxClient(APIClient())
sb = client.create(CreateSandboxRequest(name="bg-demo", docker_image="python:3.11-slim", timeout_minutes=30))
try:
    client.wait_for_creation(sb.id)
    job = client.start_background_job(sb.id, "for i in $(seq 1 20); do echo tick $i; sleep 1; done; echo done >&2; exit 3", working_dir="/tmp",
                                      env={"PYTHONUNBUFFERED": "1"})  # fmt: skip
    while True:
        st = client.get_background_job(sb.id, job)
        print(f"completed={st.completed} stdout_tail={(st.stdout or '').strip().splitlines()[-1:] }")
        if st.completed:
            print("exit code", st.exit_code, "| stderr:", (st.stderr or "").strip())
            break
        time.sleep(3)
    # or block in one call:
    r = client.run_background_job(sb.id, "sleep 5 && echo finished", timeout=120)
    print(r.exit_code, r.stdout.strip())
finally:
    client.delete(sb.id)
