#!/bin/bash
# Run ON the GPU box (via gpurun): ncu --set full captures of every hand-written kernel + launch breakdown of one step.
# gpurun copies back at most 64 MiB, so every report is exported to CSV here and oversized .ncu-rep files are dropped.
set -u
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none"
$NCU --import-source on -k regex:gemm_bf16_kernel -s 30 -c 2 -o gpurun_out/ncu/gemm -f python tools/gemm_bench.py > gpurun_out/ncu/gemm.log 2>&1
$NCU --import-source on -k regex:'flash_fwd|bwd_dkdv_kernel|bwd_dq_kernel|bwd_delta_kernel' -s 16 -c 4 -o gpurun_out/ncu/attn -f python tools/attn_bench.py > gpurun_out/ncu/attn.log 2>&1
$NCU -k regex:'rmsnorm|rope|swiglu|cross_entropy|colsum|grad_reduce|norm_publish|adamw_push|pseudograd|outer_nesterov|cast_push' -c 40 -o gpurun_out/ncu/ops -f \
    python tools/op_bench.py --once > gpurun_out/ncu/ops.log 2>&1
$NCU --import-source on -k regex:'gemm_mxfp8_kernel|quantize_mxfp8' -s 6 -c 4 -o gpurun_out/ncu/mxfp8 -f python tools/mxfp8_bench.py > gpurun_out/ncu/mxfp8.log 2>&1
# round 2: the kernels added since (captured separately: the first matches would otherwise use up the -c budget)
$NCU --import-source on -k regex:'cross_entropy_reg|embedding|grad_reduce' -c 12 -o gpurun_out/ncu/ops2 -f python tools/op_bench.py --once > gpurun_out/ncu/ops2.log 2>&1
$NCU --import-source on -k regex:'outer_nesterov|pseudograd' -c 2 -o gpurun_out/ncu/outer -f python tools/op_bench.py --once > gpurun_out/ncu/outer.log 2>&1
$NCU --import-source on -k regex:'adamw_push' -c 4 -o gpurun_out/ncu/adamw -f python tools/op_bench.py --once > gpurun_out/ncu/adamw.log 2>&1
# SwiGLU-backward epilogue GEMM (EPI = 3) out of a real step; parameter-gather GEMMs (IO = 3) from the single-rank tests
$NCU --import-source on --kernel-name-base demangled -k regex:'gemm_bf16_kernel<.*3, .*0>' -s 8 -c 2 -o gpurun_out/ncu/gemm_epi3 -f python bench.py --steps 1 --warmup 1 --no-b0 > gpurun_out/ncu/gemm_epi3.log 2>&1
$NCU --import-source on --kernel-name-base demangled -k regex:'gemm_bf16_kernel<.*3>' -c 4 -o gpurun_out/ncu/gemm_io3 -f python -m pytest tests/test_kernels_gpu.py -m gpu -q -k weight_gather_gemm_single_rank > gpurun_out/ncu/gemm_io3.log 2>&1
for r in gemm attn ops mxfp8 ops2 outer adamw gemm_epi3 gemm_io3; do
  ncu -i gpurun_out/ncu/$r.ncu-rep --page raw --csv > gpurun_out/ncu/$r.raw.csv 2>/dev/null
  # per-instruction stall samples of the tensor-core kernels (read with tools/ncu_source_hotspots.py)
  if [ "$r" = attn ] || [ "$r" = mxfp8 ]; then ncu -i gpurun_out/ncu/$r.ncu-rep --page source --csv > gpurun_out/ncu/$r.source.csv 2>/dev/null; fi
  sz=$(stat -c %s gpurun_out/ncu/$r.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/ncu/$r.ncu-rep; echo "dropped $r.ncu-rep ($sz bytes), kept CSV"; fi
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 5100 -c 1750 --csv --log-file gpurun_out/launches_native.csv \
    python bench.py --steps 1 --warmup 3 --no-b0 > gpurun_out/launches_native.log 2>&1
python tools/launch_breakdown.py gpurun_out/launches_native.csv > gpurun_out/step_composition.md
du -sh gpurun_out; ls -la gpurun_out/ncu
