#!/bin/bash
# One GPU session, cheapest / most basic checks first (round 2).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=${1:-a}
timeout 300 python profiles/microbench_conv2.py > gpurun_out/micro2_$T.txt 2>&1; echo "== micro conv rc=$?"; tail -3 gpurun_out/micro2_$T.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_r2$T.log 2>&1; echo "== pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2$T.log | tail -20
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2$T.json 2> gpurun_out/bench_r2$T.err; echo "== bench graph rc=$?"; cat gpurun_out/bench_r2$T.json | cut -c1-260; tail -2 gpurun_out/bench_r2$T.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode eager --no-grid41 > gpurun_out/bench_r2${T}_eager.json 2> gpurun_out/bench_r2${T}_eager.err; echo "== bench eager rc=$?"; cat gpurun_out/bench_r2${T}_eager.json | cut -c1-260; tail -2 gpurun_out/bench_r2${T}_eager.err
