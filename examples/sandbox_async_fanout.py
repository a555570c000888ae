"""High-volume pattern: N sandboxes × M commands each over ONE pooled async client.

    python examples/sandbox_async_fanout.py --sandboxes 20 --commands 200 --concurrency 256

Reports creation time, command throughput and the latency distribution. Every sandbox is labelled, and the label is
used for a bulk delete at the end (also on Ctrl-C).
"""

import argparse
import asyncio
import statistics
import time
import uuid

from prime_b200.platform.sandboxes import AsyncSandboxClient, CreateSandboxRequest


async def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sandboxes", type=int, default=5)
    ap.add_argument("--commands", type=int, default=50)
    ap.add_argument("--concurrency", type=int, default=64)
    a = ap.parse_args()
    label = f"fanout-{uuid.uuid4().hex[:8]}"
    async with AsyncSandboxClient() as client:
        try:
            t0 = time.perf_counter()
            created = await asyncio.gather(*[
                client.create(CreateSandboxRequest(name=f"{label}-{i}", docker_image="python:3.11-slim", timeout_minutes=20, labels=[label]))
                for i in range(a.sandboxes)])  # fmt: skip
            states = await client.bulk_wait_for_creation([s.id for s in created])
            ready = [sid for sid, st in states.items() if st == "RUNNING"]
            print(f"{len(ready)}/{a.sandboxes} sandboxes running after {time.perf_counter() - t0:.1f}s")

            gate, lat, failed = asyncio.Semaphore(a.concurrency), [], 0

            async def one(sid: str, j: int) -> None:
                nonlocal failed
                async with gate:
                    t = time.perf_counter()
                    try:
                        r = await client.execute_command(sid, f"echo {j} && python -c 'print({j}*{j})'", timeout=30)
                        assert r.exit_code == 0 and r.stdout.split()[-1] == str(j * j)
                        lat.append(time.perf_counter() - t)
                    except Exception:
                        failed += 1

            t1 = time.perf_counter()
            await asyncio.gather(*[one(sid, j) for sid in ready for j in range(a.commands)])
            dt = time.perf_counter() - t1
            lat.sort()
            q = lambda p: lat[min(len(lat) - 1, int(p * len(lat)))] * 1e3  # noqa: E731
            print(f"{len(lat)} commands in {dt:.1f}s = {len(lat) / dt:.0f}/s, {failed} failed; latency ms: p50 {q(.5):.0f}  p90 {q(.9):.0f}  p99 {q(.99):.0f}  "
                  f"mean {statistics.mean(lat) * 1e3:.0f}")  # fmt: skip
        finally:
            res = await client.bulk_delete(labels=[label])
            print(f"cleanup: {len(res.succeeded)} deleted, {len(res.failed)} failed")


if __name__ == "__main__":
    asyncio.run(main())
