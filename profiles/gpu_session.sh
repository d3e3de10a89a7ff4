#!/bin/bash
# One GPU session, cheapest / most basic checks first (round 2).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=${1:-a}
timeout 300 python profiles/microbench_conv2.py > gpurun_out/micro2_$T.txt 2>&1; echo "== micro conv rc=$?"; tail -3 gpurun_out/micro2_$T.txt
timeout 300 python profiles/microbench_wgrad2.py > gpurun_out/microw_$T.txt 2>&1; echo "== micro wgrad rc=$?"; tail -3 gpurun_out/microw_$T.txt
for V in ${VARIANTS:-}; do
  if [ -f virconv_b200/lib/libvirconv_sm100_$V.so ]; then
    VIRCONV_LIB=virconv_b200/lib/libvirconv_sm100_$V.so timeout 300 python profiles/microbench_conv2.py > gpurun_out/micro2_${T}_$V.txt 2>&1; echo "== micro conv $V rc=$?"; tail -2 gpurun_out/micro2_${T}_$V.txt
  fi
done
TORCH_SHOW_CPP_STACKTRACES=1 timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -q -x -k rotating > gpurun_out/pytest_rot_$T.log 2>&1; echo "== rotating rc=$?"; tail -3 gpurun_out/pytest_rot_$T.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_r2$T.log 2>&1; echo "== pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2$T.log | tail -20
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2$T.json 2> gpurun_out/bench_r2$T.err; echo "== bench graph rc=$?"; cat gpurun_out/bench_r2$T.json; tail -5 gpurun_out/bench_r2$T.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode eager > gpurun_out/bench_r2${T}_eager.json 2> gpurun_out/bench_r2${T}_eager.err; echo "== bench eager rc=$?"; cat gpurun_out/bench_r2${T}_eager.json; tail -3 gpurun_out/bench_r2${T}_eager.err
VIRCONV_LIB=virconv_b200/lib/libvirconv_sm100_trace.so timeout 200 python profiles/trace_tc2.py > gpurun_out/trace2_$T.txt 2>&1; echo "== trace rc=$?"; tail -30 gpurun_out/trace2_$T.txt
