"""Rank-aware logger: one line per event, prefixed with worker/rank; INFO on the worker leader, WARNING elsewhere."""

from __future__ import annotations

import logging
import os
import sys

_FMT = "%(asctime)s [%(name)s] %(levelname)s %(message)s"


def get_logger(worker: str | int | None = None, rank: int | None = None, level: str | None = None) -> logging.Logger:
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    tag = worker if isinstance(worker, str) else f"w{worker}"
    name = f"pb200 {tag} r{rank}" if worker is not None else f"pb200 r{rank}"
    log = logging.getLogger(name)
    if not log.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter(_FMT, datefmt="%H:%M:%S"))
        log.addHandler(h)
        log.propagate = False
    lvl = level or os.environ.get("PRIME_B200_LOG_LEVEL") or ("INFO" if rank == 0 else "WARNING")
    log.setLevel(lvl.upper())
    return log
