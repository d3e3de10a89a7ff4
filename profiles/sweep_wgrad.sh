#!/bin/bash
# wgrad-on-its-own-stream sweep (run on the GPU box): ms/step of bench.py for several CTA counts
mkdir -p gpurun_out
run() {
  echo -n "$* : "
  env "$@" timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.3f  e2e %.3f  mallocs %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['cuda_mallocs_in_timed_region']))"
}
for rep in 1 2; do
for c in 64 96 112 128 148; do
run VIRCONV_WGRAD_STREAM=1 VIRCONV_WGRAD_CTAS=$c
done
done
run VIRCONV_WGRAD_STREAM=0
