"""Print what a checkpoint holds:  python examples/inspect_checkpoint.py /tmp/prime_b200/1b [--step 500]"""

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from prime_b200 import checkpoint as ck  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("root")
ap.add_argument("--step", type=int)
a = ap.parse_args()
path = ck.step_dir(Path(a.root), a.step) if a.step is not None else ck.resolve_resume("latest", a.root)
if path is None:
    sys.exit(f"no complete checkpoint under {a.root}")
meta = json.loads((path / "meta.json").read_text())
print(f"{path}: step {meta['step']}, written by {meta['world_size']} ranks, mesh {meta.get('mesh')}, {meta.get('n_params', 0) / 1e6:.1f}M params")
for shard in sorted(path.glob("rank_*.pbck")):
    tensors, extra = ck.read_shard(shard)
    size = sum(t.numel() * t.element_size() for t in tensors.values())
    print(f"  {shard.name}: {size / 2**20:.1f} MiB  fsdp_rank {extra['fsdp_rank']}/{extra['fsdp_size']}  trainer_step {extra['trainer_step']}")
    for name, t in tensors.items():
        print(f"      {name:12s} {str(t.dtype):14s} {tuple(t.shape)}  |x|₂={float(t.float().norm()):.4g}")
