mkdir -p gpurun_out/ncu
ncu --set full --clock-control none --import-source on -k regex:flash_fwd2_kernel -s 3 -c 1 -o gpurun_out/ncu/fwd2 -f python tools/attn_bench.py > gpurun_out/ncu/fwd2.log 2>&1
ncu -i gpurun_out/ncu/fwd2.ncu-rep --page source --csv > gpurun_out/ncu/fwd2.source.csv 2>/dev/null
ncu -i gpurun_out/ncu/fwd2.ncu-rep --page raw --csv > gpurun_out/ncu/fwd2.raw.csv 2>/dev/null
python tools/ncu_source_hotspots.py gpurun_out/ncu/fwd2.source.csv 25 | tee gpurun_out/fwd2_hotspots.txt | head -60
rm -f gpurun_out/ncu/fwd2.ncu-rep
