#!/bin/bash
# wgrad-on-its-own-stream sweep (run on the GPU box): ms/step of bench.py for several (CTAs, smem floor) settings
mkdir -p gpurun_out
run() {
  echo -n "$* : "
  env "$@" timeout 120 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  e2e %.3f  wgrad_tc %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['per_step_ms'].get('conv_wgrad_tc', 0)))"
}
run VIRCONV_WGRAD_STREAM=0
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=148 VIRCONV_WGRAD_SMEM_KB=0
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=64 VIRCONV_WGRAD_SMEM_KB=0
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=96 VIRCONV_WGRAD_SMEM_KB=0
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=64 VIRCONV_WGRAD_SMEM_KB=190
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=80 VIRCONV_WGRAD_SMEM_KB=190
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=96 VIRCONV_WGRAD_SMEM_KB=190
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=112 VIRCONV_WGRAD_SMEM_KB=190
run VIRCONV_WGRAD_STREAM=0
