PB_MX_DEBUG=0 timeout 300 python tools/mxfp8_bench.py 2>gpurun_out/mx.err > gpurun_out/mxfp8_bench.json; tail -2 gpurun_out/mx.err; python -c "
import json,sys; d=json.load(open('gpurun_out/mxfp8_bench.json'))
for r in d['rows']: print(' ', r['shape'], r['mxfp8_tflops'], r['mxfp8_single_cta_tflops'], r['bf16_tflops'])
print(json.dumps(d['sustained_w13_shape']))"
