"""Elastic DiLoCo membership: workers are launched independently, meet at outer-step boundaries, may die or join.

Each *worker* is its own ``torchrun`` world of F ranks (its FSDP group: NCCL / the fused P2P heap never span workers, so a
worker crash cannot wedge another worker's collectives). Workers find each other through ONE global ``TCPStore``
(``GLOBAL_ADDR:GLOBAL_PORT``; ``python -m prime_b200.parallel.elastic serve`` runs it as a standalone daemon so that no
training process is a single point of failure).

Protocol — all keys live in the global store; only fsdp-rank-0 ("leader") of each worker writes, every rank reads:

* ``hb/<wid>``            wall-clock heartbeat, refreshed by a daemon thread with its own store connection.
* ``join/<wid>``          a worker that is not (or no longer) a member asks to be admitted.
* ``arrive/<e>/<wid>``    a member reached outer boundary ``e``.
* ``decider/<e>``         lease ``"<wid>:<ts>"`` taken with ``compare_set``; a stale lease (> heartbeat timeout) is stolen.
* ``members/<e>``         the decision: ``{"workers": [...], "joiners": [...], "source": wid|None}`` — written once.
* ``admit/<wid>``         epoch in which a joiner becomes a member.

The decider for epoch ``e`` (whichever arrived member wins the lease) waits until every member of ``e-1`` whose heartbeat is
still fresh has arrived — that wait *is* the outer-step barrier — drops the ones whose heartbeat went stale, admits live
joiners, and publishes. Members then build one process group per fsdp-rank (``PrefixStore("pg/<e>/<r>")``, gloo on CPU,
NCCL on GPU) unless the membership is unchanged, and a joiner receives the *live checkpoint* — inner master, Adam moments,
θ₀ and outer momentum — by broadcast from a surviving worker over that same group (no disk, no extra server).

A member that was declared dead while merely slow finds itself missing from ``members/<e>`` and falls back to the joiner
path, so the state machine has exactly two states (member / joiner) and every transition goes through the same code.

The reference has no training-side analogue; its closest mechanisms are the tunnel supervisor's liveness loop and the
sandbox client's not-running classification (reference: packages/prime-tunnel/src/prime_tunnel/tunnel.py:246-302,
packages/prime-sandboxes/src/prime_sandboxes/sandbox.py:78-140). BASELINE.json's "elastic 4→3→4" config is the contract.
"""

from __future__ import annotations

import argparse
import datetime
import json
import os
import threading
import time
from dataclasses import dataclass, field
from typing import Any, Callable

import torch
import torch.distributed as dist


def connect(host: str, port: int, *, is_master: bool = False, timeout_s: float = 300.0) -> dist.TCPStore:
    return dist.TCPStore(host, port, None, is_master, datetime.timedelta(seconds=timeout_s), wait_for_workers=False)


@dataclass
class Membership:
    epoch: int
    workers: list[str]
    joiners: list[str]
    source: str | None
    index: int  # my worker's position in ``workers``
    pg: Any = None  # per-fsdp-rank process group over the workers (None when alone)
    changed: bool = True

    @property
    def size(self) -> int:
        return len(self.workers)


@dataclass
class ElasticConfig:
    heartbeat_interval_s: float = 2.0
    heartbeat_timeout_s: float = 20.0
    min_workers: int = 1  # cold start: wait for this many joiners (or ``start_grace_s``)
    start_grace_s: float = 10.0
    pg_timeout_s: float = 120.0
    poll_s: float = 0.02


class StoreView:
    """Read-only view of the rendezvous state (also what ``elastic status`` prints)."""

    def __init__(self, store: dist.TCPStore, cfg: ElasticConfig | None = None, clock: Callable[[], float] = time.time):
        self.store, self.cfg, self.clock = store, cfg or ElasticConfig(), clock

    def _has(self, key: str) -> bool:
        return bool(self.store.check([key]))

    def _get(self, key: str) -> str:
        return self.store.get(key).decode()

    def _roster(self) -> list[str]:
        n = self.store.add("roster/n", 0)
        return [self._get(f"roster/{i}") for i in range(n) if self._has(f"roster/{i}")]

    def fresh(self, wid: str) -> bool:
        if not self._has(f"hb/{wid}"):
            return False
        return self.clock() - float(self._get(f"hb/{wid}")) <= self.cfg.heartbeat_timeout_s

    def _members_of(self, epoch: int) -> dict[str, Any] | None:
        return json.loads(self._get(f"members/{epoch}")) if epoch > 0 and self._has(f"members/{epoch}") else None

    def _latest_epoch(self) -> int:
        return int(self._get("epoch/latest")) if self._has("epoch/latest") else 0

    def status(self) -> dict[str, Any]:
        latest = self._latest_epoch()
        return {"epoch": latest, "members": self._members_of(latest), "roster": {w: self.fresh(w) for w in self._roster()}}


def display_name(internal_id: str) -> str:
    return internal_id.split("#", 1)[0]


class ElasticContext(StoreView):
    def __init__(self, store: dist.TCPStore, worker_id: str, *, fsdp_rank: int = 0, backend: str = "gloo",
                 cfg: ElasticConfig | None = None, hb_store_factory: Callable[[], dist.TCPStore] | None = None,
                 clock: Callable[[], float] = time.time, incarnation: int | None = None):  # fmt: skip
        super().__init__(store, cfg, clock)
        # internal id = name#incarnation: a restarted worker never collides with keys of its previous life
        self.name = str(worker_id)
        if incarnation is None:
            incarnation = store.add(f"incarnation/{self.name}", 1) if fsdp_rank == 0 else store.add(f"incarnation/{self.name}", 0)
        self.wid, self.r = f"{self.name}#{incarnation:04d}", fsdp_rank
        self.leader = fsdp_rank == 0
        self.backend = backend
        self.membership: Membership | None = None
        self.epoch = 0  # last epoch this worker was a member of
        self._join_gen = 0  # every rank of a worker walks the same member/joiner transitions, so local counts agree
        self._hb_stop = threading.Event()
        self._hb_thread: threading.Thread | None = None
        self._hb_store_factory = hb_store_factory
        self.events: list[tuple[float, str]] = []  # (time, message) — surfaced in the training log
        if self.leader:
            self._register()
            self._beat(self.store)
            self._start_heartbeat()

    # ------------------------------------------------------------------ store helpers
    def _wait_for(self, key: str, *, while_waiting: Callable[[], None] | None = None, timeout_s: float | None = None) -> str:
        t0 = self.clock()
        while not self._has(key):
            if while_waiting is not None:
                while_waiting()
            if timeout_s is not None and self.clock() - t0 > timeout_s:
                raise TimeoutError(f"elastic: timed out waiting for {key}")
            time.sleep(self.cfg.poll_s)
        return self._get(key)

    def _log(self, msg: str) -> None:
        self.events.append((self.clock(), msg))

    # ------------------------------------------------------------------ roster + heartbeats
    def _register(self) -> None:
        if not self._has(f"registered/{self.wid}"):
            n = self.store.add("roster/n", 1)
            self.store.set(f"roster/{n - 1}", self.wid)
            self.store.set(f"registered/{self.wid}", "1")

    def _beat(self, store: dist.TCPStore) -> None:
        store.set(f"hb/{self.wid}", repr(self.clock()))

    def _start_heartbeat(self) -> None:
        def loop() -> None:
            store = self._hb_store_factory() if self._hb_store_factory else self.store
            while not self._hb_stop.wait(self.cfg.heartbeat_interval_s):
                try:
                    self._beat(store)
                except Exception:
                    return  # store gone: the job is shutting down

        self._hb_thread = threading.Thread(target=loop, name=f"elastic-hb-{self.wid}", daemon=True)
        self._hb_thread.start()

    def close(self, *, leave: bool = True) -> None:
        self._hb_stop.set()
        if self.leader and leave:
            try:
                self.store.set(f"hb/{self.wid}", "0")  # graceful exit: peers drop us at the next boundary without waiting
            except Exception:
                pass

    # ------------------------------------------------------------------ the decision
    def _acquire(self, epoch: int) -> bool:
        key, now = f"decider/{epoch}", self.clock()
        mine = f"{self.wid}:{now!r}"
        cur = self.store.compare_set(key, "", mine).decode()
        if cur == mine:
            return True
        owner, ts = cur.rsplit(":", 1)
        if owner == self.wid:
            return True
        if now - float(ts) > self.cfg.heartbeat_timeout_s:
            return self.store.compare_set(key, cur, mine).decode() == mine
        return False

    def _pending_joiners(self) -> list[str]:
        return [w for w in self._roster() if self._has(f"join/{w}") and self.fresh(w)]

    def _decide(self, epoch: int) -> None:
        """Runs on the lease holder only. Returns once ``members/<epoch>`` exists."""
        prev = (self._members_of(epoch - 1) or {}).get("workers", [])
        t0 = self.clock()
        while True:
            arrived = [w for w in prev if self._has(f"arrive/{epoch}/{w}")]
            waiting_on = [w for w in prev if w not in arrived and self.fresh(w)]
            joiners = [w for w in self._pending_joiners() if w not in arrived]
            if not waiting_on:
                if arrived:
                    break
                # cold start (or every previous member died): wait for a quorum of joiners, bounded by the grace period
                if joiners and (len(joiners) >= self.cfg.min_workers or self.clock() - t0 > self.cfg.start_grace_s):
                    break
            self.store.compare_set(f"decider/{epoch}", self._get(f"decider/{epoch}"), f"{self.wid}:{self.clock()!r}")  # renew lease
            time.sleep(self.cfg.poll_s)
        workers = sorted(set(arrived) | set(joiners))
        decision = {"workers": workers, "joiners": sorted(joiners), "source": sorted(arrived)[0] if arrived else None,
                    "dropped": sorted(set(prev) - set(arrived))}  # fmt: skip
        if self.store.compare_set(f"members/{epoch}", "", json.dumps(decision)).decode() != json.dumps(decision):
            return  # somebody else published first (lease takeover race): theirs stands
        self.store.set("epoch/latest", str(epoch))
        for j in joiners:
            admit = self._get(f"join/{j}")
            self.store.delete_key(f"join/{j}")
            self.store.set(admit, str(epoch))

    # ------------------------------------------------------------------ public: meet at the boundary
    def rendezvous(self) -> Membership:
        """Block until the membership of the next epoch is decided; returns it (with a process group when size > 1)."""
        decision: dict[str, Any] | None = None
        epoch = 0
        if self.membership is not None:  # member path
            epoch = self.epoch + 1
            if self.leader:
                self.store.set(f"arrive/{epoch}/{self.wid}", "1")

                def maybe_decide() -> None:
                    if self._acquire(epoch):
                        self._decide(epoch)

                self._wait_for(f"members/{epoch}", while_waiting=maybe_decide)
            else:
                self._wait_for(f"members/{epoch}")
            decision = self._members_of(epoch)
            if self.wid not in decision["workers"]:
                self._log(f"evicted at epoch {epoch} (declared dead while slow); rejoining")
                self.membership, decision = None, None
        if decision is None:  # joiner path (cold start, late join, or eviction)
            self._join_gen += 1
            admit = f"admit/{self.wid}/{self._join_gen}"
            if self.leader:
                self.store.set(f"join/{self.wid}", admit)

                def maybe_bootstrap() -> None:
                    latest = self._latest_epoch()
                    last = self._members_of(latest)
                    alive = [w for w in (last or {}).get("workers", []) if self.fresh(w)]
                    if not alive and self._acquire(latest + 1):  # nobody left to admit us: bootstrap the next epoch ourselves
                        self._decide(latest + 1)

                epoch = int(self._wait_for(admit, while_waiting=maybe_bootstrap))
            else:
                epoch = int(self._wait_for(admit))
            decision = self._wait_and_load(epoch)
        prev = self.membership
        workers = decision["workers"]
        m = Membership(epoch, workers, decision["joiners"], decision.get("source"), workers.index(self.wid))
        if prev is not None and prev.workers == workers and not decision["joiners"]:
            m.pg, m.changed = prev.pg, False
        else:
            m.pg = self._new_group(epoch, m.index, len(workers)) if len(workers) > 1 else None
            if prev is not None:
                self._log(f"epoch {epoch}: membership {prev.workers} → {workers} (dropped {decision.get('dropped', [])}, joined {decision['joiners']})")
        self.membership, self.epoch = m, epoch
        return m

    def _wait_and_load(self, epoch: int) -> dict[str, Any]:
        self._wait_for(f"members/{epoch}")
        return self._members_of(epoch)

    def _new_group(self, epoch: int, rank: int, size: int):
        if self.backend == "none":  # protocol-only mode (tests, dry runs): membership without a communicator
            return None
        pstore = dist.PrefixStore(f"pg/{epoch}/{self.r}", self.store)
        timeout = datetime.timedelta(seconds=self.cfg.pg_timeout_s)
        if self.backend == "nccl":
            opts = dist.ProcessGroupNCCL.Options()
            opts._timeout = timeout
            return dist.ProcessGroupNCCL(pstore, rank, size, opts)
        return dist.ProcessGroupGloo(pstore, rank, size, timeout)

    # ------------------------------------------------------------------ live checkpoint for joiners
    def sync_state(self, tensors: list[torch.Tensor], counters: dict[str, int]) -> dict[str, int]:
        """If this epoch admitted joiners and a survivor exists, broadcast ``tensors`` from the source worker (in place on
        the joiners) and return the source's ``counters`` (step numbers); otherwise return ``counters`` unchanged."""
        m = self.membership
        assert m is not None
        if not m.joiners or m.source is None or m.pg is None:
            return counters
        root = m.workers.index(m.source)
        key = f"state/{m.epoch}/{self.r}"
        if m.index == root:
            self.store.set(key, json.dumps(counters))
        opts = dist.BroadcastOptions()
        opts.rootRank, opts.rootTensor = root, 0
        for t in tensors:
            m.pg.broadcast([t], opts).wait()
        if self.wid in m.joiners:
            got = json.loads(self._wait_for(key, timeout_s=self.cfg.pg_timeout_s))
            self._log(f"epoch {m.epoch}: received live checkpoint from {m.source} ({sum(t.numel() * t.element_size() for t in tensors) >> 20} MiB)")
            return got
        return counters

    @classmethod
    def from_env(cls, *, fsdp_rank: int, backend: str, cfg: ElasticConfig | None = None) -> "ElasticContext":
        host = os.environ.get("GLOBAL_ADDR", "127.0.0.1")
        port = int(os.environ["GLOBAL_PORT"])
        wid = os.environ.get("GLOBAL_UNIQUE_ID") or os.environ.get("GLOBAL_RANK") or f"w{os.environ.get('MASTER_PORT', os.getpid())}"
        store = connect(host, port)
        inc = None
        if dist.is_initialized() and dist.get_world_size() > 1:  # all ranks of this worker must agree on the incarnation
            box = [store.add(f"incarnation/{wid}", 1) if fsdp_rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            inc = box[0]
        return cls(store, wid, fsdp_rank=fsdp_rank, backend=backend, cfg=cfg, hb_store_factory=lambda: connect(host, port), incarnation=inc)


@dataclass
class _Served:
    store: dist.TCPStore
    port: int
    stop: threading.Event = field(default_factory=threading.Event)


def serve(port: int, host: str = "0.0.0.0") -> _Served:
    """Host the global store in this process (the daemon entrypoint below, and tests)."""
    store = dist.TCPStore(host if host != "0.0.0.0" else "127.0.0.1", port, None, True, datetime.timedelta(seconds=300), wait_for_workers=False)
    return _Served(store, store.port)


def main(argv: list[str] | None = None) -> None:
    ap = argparse.ArgumentParser(prog="python -m prime_b200.parallel.elastic", description="Global rendezvous store for elastic DiLoCo workers")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("serve")
    s.add_argument("--port", type=int, default=int(os.environ.get("GLOBAL_PORT", "29400")))
    st = sub.add_parser("status")
    st.add_argument("--port", type=int, default=int(os.environ.get("GLOBAL_PORT", "29400")))
    st.add_argument("--addr", default=os.environ.get("GLOBAL_ADDR", "127.0.0.1"))
    a = ap.parse_args(argv)
    if a.cmd == "serve":
        served = serve(a.port)
        print(json.dumps({"serving": served.port}), flush=True)
        try:
            while True:
                time.sleep(3600)
        except KeyboardInterrupt:
            pass
    else:
        print(json.dumps(StoreView(connect(a.addr, a.port)).status()))


if __name__ == "__main__":
    main()
