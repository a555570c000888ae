"""Elastic membership protocol over a real TCPStore: cold-start quorum, drop on stale heartbeat, late join with a live
checkpoint source, eviction of a slow member, decider-lease takeover (SURVEY §5 failure detection; BASELINE elastic 4→3→4)."""

import json
import threading
import time

import pytest
import torch

from prime_b200.parallel import elastic as el

CFG = dict(heartbeat_interval_s=0.05, heartbeat_timeout_s=0.6, start_grace_s=0.5, poll_s=0.005, pg_timeout_s=20)


@pytest.fixture
def served():
    s = el.serve(0)
    yield s
    del s.store


def ctx(served, name, **kw):
    cfg = el.ElasticConfig(**{**CFG, **kw.pop("cfg", {})})
    mk = lambda: el.connect("127.0.0.1", served.port)  # noqa: E731
    return el.ElasticContext(mk(), name, backend=kw.pop("backend", "none"), cfg=cfg, hb_store_factory=mk, **kw)


def together(*fns, timeout=20):
    out, err = [None] * len(fns), []

    def run(i, f):
        try:
            out[i] = f()
        except BaseException as e:  # noqa: BLE001
            err.append(e)

    ts = [threading.Thread(target=run, args=(i, f)) for i, f in enumerate(fns)]
    [t.start() for t in ts]
    [t.join(timeout) for t in ts]
    assert not any(t.is_alive() for t in ts), "rendezvous hung"
    if err:
        raise err[0]
    return out


def names(m):
    return [el.display_name(w) for w in m.workers]


def test_cold_start_drop_and_join(served):
    a, b, c = (ctx(served, n, cfg={"min_workers": 3}) for n in "abc")
    ma, mb, mc = together(a.rendezvous, b.rendezvous, c.rendezvous)
    assert names(ma) == ["a", "b", "c"] and ma.epoch == mb.epoch == mc.epoch == 1 and ma.source is None
    assert [m.index for m in (ma, mb, mc)] == [0, 1, 2] and sorted(ma.joiners) == ma.workers
    # steady state: same members → nothing changes, no new communicator
    ma, mb, mc = together(a.rendezvous, b.rendezvous, c.rendezvous)
    assert ma.epoch == 2 and not ma.changed and not ma.joiners
    # c dies (graceful leave marks its heartbeat stale immediately)
    c.close()
    ma, mb = together(a.rendezvous, b.rendezvous)
    assert names(ma) == ["a", "b"] and ma.epoch == 3 and ma.changed and any("dropped" in msg for _, msg in a.events + b.events)
    # d joins late: it blocks until the members reach their next boundary, then receives state from the lowest survivor
    d = ctx(served, "d")
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("m", d.rendezvous()))
    t.start()
    time.sleep(0.2)
    assert t.is_alive()  # not admitted before a boundary
    ma, mb = together(a.rendezvous, b.rendezvous)
    t.join(10)
    md = box["m"]
    assert names(ma) == ["a", "b", "d"] and md.epoch == ma.epoch == 4 and [el.display_name(j) for j in md.joiners] == ["d"]
    assert el.display_name(md.source) == "a" and md.index == 2
    st = el.StoreView(el.connect("127.0.0.1", served.port), el.ElasticConfig(**CFG)).status()
    assert st["epoch"] == 4 and sum(st["roster"].values()) == 3
    for x in (a, b, d):
        x.close()


def test_crash_without_goodbye_is_detected_by_heartbeat_timeout(served):
    a, b = ctx(served, "a", cfg={"min_workers": 2}), ctx(served, "b", cfg={"min_workers": 2})
    together(a.rendezvous, b.rendezvous)
    b._hb_stop.set()  # simulated SIGKILL: heartbeats just stop
    t0 = time.time()
    m = a.rendezvous()
    assert names(m) == ["a"] and m.pg is None and time.time() - t0 >= CFG["heartbeat_timeout_s"] * 0.8
    a.close()


def test_slow_member_is_evicted_then_rejoins_with_same_name(served):
    a, b = ctx(served, "a", cfg={"min_workers": 2}), ctx(served, "b", cfg={"min_workers": 2})
    together(a.rendezvous, b.rendezvous)
    b._hb_stop.set()  # b stalls (e.g. a long GC pause): no heartbeats, no arrival
    m = a.rendezvous()
    assert names(m) == ["a"]
    b._hb_stop.clear()
    b._start_heartbeat()
    b._beat(b.store)  # b wakes up, believes it is still a member of epoch 1 → arrives at epoch 2, finds itself dropped
    mb, ma = together(b.rendezvous, lambda: (time.sleep(0.1), a.rendezvous())[1])
    assert names(ma) == ["a", "b"] and mb.epoch == ma.epoch == 3 and mb.joiners == [b.wid] and any("evicted" in msg for _, msg in b.events)
    # a restarted process with the same name gets a new incarnation, so stale keys of its previous life cannot admit it
    b.close()
    b2 = ctx(served, "b")
    assert b2.wid != b.wid and el.display_name(b2.wid) == "b"
    mb2, ma = together(b2.rendezvous, a.rendezvous)
    assert ma.workers == sorted([a.wid, b2.wid]) and mb2.joiners == [b2.wid]
    a.close(), b2.close()


def test_stale_decider_lease_is_taken_over(served):
    a = ctx(served, "a")
    a.rendezvous()
    # someone grabbed the lease for the next epoch long ago and died
    a.store.set("decider/2", f"ghost#0001:{time.time() - 100!r}")
    m = a.rendezvous()
    assert m.epoch == 2 and names(m) == ["a"]
    # a live lease held by someone else is respected
    a.store.set("decider/3", f"ghost#0001:{time.time()!r}")
    assert not a._acquire(3)
    a.close()


def test_live_checkpoint_broadcast_over_gloo(served):
    a, b = ctx(served, "a", backend="gloo", cfg={"min_workers": 2}), ctx(served, "b", backend="gloo", cfg={"min_workers": 2})
    ma, mb = together(a.rendezvous, b.rendezvous)
    assert ma.pg is not None and ma.source is None
    ta, tb = torch.arange(8, dtype=torch.float32), torch.zeros(8)
    # cold start: no source → nothing is overwritten
    ca, cb = together(lambda: a.sync_state([ta], {"step": 7}), lambda: b.sync_state([tb], {"step": 0}))
    assert cb == {"step": 0} and tb.sum() == 0
    c = ctx(served, "c", backend="gloo")
    tc = torch.zeros(8)
    mc, ma, mb = together(c.rendezvous, a.rendezvous, b.rendezvous)
    assert mc.source == a.wid
    cc, ca, cb = together(lambda: c.sync_state([tc], {"step": 0}), lambda: a.sync_state([ta], {"step": 7}), lambda: b.sync_state([tb], {"step": 7}))
    assert cc == {"step": 7} and torch.equal(tc, ta) and any("live checkpoint" in msg for _, msg in c.events)
    # the group really spans the three workers
    outs = together(*[lambda x=x, t=t: x.membership.pg.allreduce([t]).wait() for x, t in ((a, torch.ones(1)), (b, torch.ones(1)), (c, torch.ones(1)))])
    assert outs is not None
    for x in (a, b, c):
        x.close()


def test_status_cli(served, capsys):
    a = ctx(served, "solo")
    a.rendezvous()
    el.main(["status", "--port", str(served.port)])
    out = json.loads(capsys.readouterr().out)
    assert out["epoch"] == 1 and [el.display_name(w) for w in out["members"]["workers"]] == ["solo"]
    a.close()
