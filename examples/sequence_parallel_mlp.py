"""Sequence-parallel SwiGLU MLP on the fused collective GEMMs (all-gather ⊕ GEMM, GEMM ⊕ reduce-scatter over NVLink peer memory).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/sequence_parallel_mlp.py

Every rank holds T/n tokens and FF/n hidden features; forward and backward move activations between GPUs only inside GEMM kernels.
"""

import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from prime_b200.parallel.mesh import init_distributed  # noqa: E402
from prime_b200.parallel.symm import SymmetricHeap, dist_exchange  # noqa: E402
from prime_b200.parallel.tensor_parallel import SequenceParallelMLP  # noqa: E402


def main() -> None:
    w = init_distributed("nccl")
    dev = torch.device("cuda", w.local_rank)
    n, r = w.world_size, w.rank
    heap = SymmetricHeap(1 << 30, r, n, dist_exchange(), dev)
    T, D, FF = 16384, 2048, 5632 // (128 * n) * (128 * n)  # Llama-1B-like, FF rounded so FF/n stays a multiple of 128
    mlp = SequenceParallelMLP(heap, list(range(n)), D, FF, T // n)
    torch.manual_seed(0)
    mlp.load_full_weights((torch.randn(2 * FF, D, device=dev) * 0.02).to(torch.bfloat16), (torch.randn(D, FF, device=dev) * 0.02).to(torch.bfloat16))
    x = torch.randn(T // n, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for step in range(3):
        mlp.zero_grad(set_to_none=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = mlp(x)
        y.backward(torch.ones_like(y))
        e1.record()
        torch.cuda.synchronize()
        if r == 0:
            print(f"step {step}: fwd+bwd {e0.elapsed_time(e1):.3f} ms, |y| {float(y.float().abs().mean()):.4f}, |dW13| {float(mlp.w13.grad.float().abs().mean()):.5f}")
    heap.check_errors()
    dist.barrier()
    heap.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
