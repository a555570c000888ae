"""Two-level device mesh on one NVSwitch box: ``num_workers`` DiLoCo workers × ``fsdp_size`` shards.

rank = worker_id * fsdp_size + fsdp_rank.  The *fsdp group* (ranks of one worker) carries the inner
loop's parameter all-gather / gradient reduce-scatter; the *diloco group* (same fsdp_rank across
workers) carries the outer pseudo-gradient all-reduce every H steps.  This is the DiLoCo engine's
ElasticDeviceMesh collapsed onto one box (BASELINE.json north star); the elastic, multi-launch
variant lives in :mod:`prime_b200.parallel.elastic`.
"""

from __future__ import annotations

import datetime
import os
from dataclasses import dataclass, field

import torch
import torch.distributed as dist


@dataclass
class WorldInfo:
    rank: int = 0
    local_rank: int = 0
    world_size: int = 1
    # multi-launch (elastic) coordinates: one launch per worker
    global_rank: int = 0
    global_world_size: int = 1
    global_unique_id: str = "0"
    global_addr: str = "127.0.0.1"
    global_port: int = 29600

    @property
    def device_index(self) -> int:
        """CUDA device of this rank. ``PRIME_B200_DEVICES="4,5"`` (set by ``launch.py`` for elastic workers) maps local ranks onto
        a slice of the box WITHOUT hiding the other GPUs: the fused elastic outer step maps the exchange buffers of other workers'
        GPUs through cudaIpc, which needs those devices visible to this process (CUDA_VISIBLE_DEVICES slicing would hide them)."""
        pool = os.environ.get("PRIME_B200_DEVICES", "")
        ids = [int(x) for x in pool.split(",") if x.strip()]
        return ids[self.local_rank] if self.local_rank < len(ids) else self.local_rank

    @classmethod
    def from_env(cls) -> "WorldInfo":
        e = os.environ
        return cls(
            rank=int(e.get("RANK", 0)),
            local_rank=int(e.get("LOCAL_RANK", 0)),
            world_size=int(e.get("WORLD_SIZE", 1)),
            global_rank=int(e.get("GLOBAL_RANK", 0)),
            global_world_size=int(e.get("GLOBAL_WORLD_SIZE", 1)),
            global_unique_id=e.get("GLOBAL_UNIQUE_ID", e.get("GLOBAL_RANK", "0")),
            global_addr=e.get("GLOBAL_ADDR", "127.0.0.1"),
            global_port=int(e.get("GLOBAL_PORT", 29600)),
        )


def init_distributed(backend: str = "auto", timeout_s: float = 600.0) -> WorldInfo:
    """Initialise the default process group from torchrun's env (127.0.0.1 rendezvous by default)."""
    info = WorldInfo.from_env()
    if backend == "auto":
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(info.device_index)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", info.device_index)
        dist.init_process_group(
            backend, rank=info.rank, world_size=info.world_size, timeout=datetime.timedelta(seconds=timeout_s), **kwargs
        )
    return info


@dataclass
class Mesh:
    world: WorldInfo
    num_workers: int
    fsdp_size: int
    device: torch.device
    fsdp_group: dist.ProcessGroup | None = None
    diloco_group: dist.ProcessGroup | None = None
    fsdp_ranks: list[int] = field(default_factory=list)  # global ranks of my worker
    diloco_ranks: list[int] = field(default_factory=list)  # global ranks holding my shard index

    @property
    def worker_id(self) -> int:
        return self.world.rank // self.fsdp_size

    @property
    def fsdp_rank(self) -> int:
        return self.world.rank % self.fsdp_size

    @property
    def is_master(self) -> bool:
        return self.world.rank == 0

    def describe(self) -> str:
        return f"dl{self.num_workers}xfsdp{self.fsdp_size}"


def resolve_shape(world_size: int, num_workers: int = 0, fsdp_size: int = 0) -> tuple[int, int]:
    if num_workers and fsdp_size:
        if num_workers * fsdp_size != world_size:
            raise ValueError(f"mesh {num_workers}x{fsdp_size} does not match world size {world_size}")
    elif fsdp_size:
        if world_size % fsdp_size:
            raise ValueError(f"world size {world_size} not divisible by fsdp_size {fsdp_size}")
        num_workers = world_size // fsdp_size
    elif num_workers:
        if world_size % num_workers:
            raise ValueError(f"world size {world_size} not divisible by num_workers {num_workers}")
        fsdp_size = world_size // num_workers
    else:
        num_workers, fsdp_size = 1, world_size
    return num_workers, fsdp_size


def build_mesh(world: WorldInfo, num_workers: int = 0, fsdp_size: int = 0, device: torch.device | None = None) -> Mesh:
    num_workers, fsdp_size = resolve_shape(world.world_size, num_workers, fsdp_size)
    if device is None:
        device = torch.device("cuda", world.device_index) if torch.cuda.is_available() else torch.device("cpu")
    mesh = Mesh(world=world, num_workers=num_workers, fsdp_size=fsdp_size, device=device)
    mesh.fsdp_ranks = [mesh.worker_id * fsdp_size + i for i in range(fsdp_size)]
    mesh.diloco_ranks = [w * fsdp_size + mesh.fsdp_rank for w in range(num_workers)]
    if dist.is_initialized() and world.world_size > 1:
        # every rank must create every group, in the same order
        for w in range(num_workers):
            ranks = [w * fsdp_size + i for i in range(fsdp_size)]
            g = dist.new_group(ranks) if fsdp_size > 1 else None
            if w == mesh.worker_id:
                mesh.fsdp_group = g
        for s in range(fsdp_size):
            ranks = [w * fsdp_size + s for w in range(num_workers)]
            g = dist.new_group(ranks) if num_workers > 1 else None
            if s == mesh.fsdp_rank:
                mesh.diloco_group = g
    return mesh
