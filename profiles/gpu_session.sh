#!/bin/bash
# Final validation session of round 2: tests, smoke, bench (graph / eager), CPU thread sweep, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=${1:-a}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/pytest_r2$T.log 2>&1; echo "== pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_r2$T.log | tail -20
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r2$T.json 2> gpurun_out/bench_r2$T.err; echo "== bench (default flags) rc=$?"; cat gpurun_out/bench_r2$T.json | cut -c1-260; tail -2 gpurun_out/bench_r2$T.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode eager --no-grid41 > gpurun_out/bench_r2${T}_eager.json 2> gpurun_out/bench_r2${T}_eager.err; echo "== bench eager rc=$?"; cat gpurun_out/bench_r2${T}_eager.json | cut -c1-260
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision fp32 --no-grid41 > gpurun_out/bench_r2${T}_fp32.json 2> gpurun_out/bench_r2${T}_fp32.err; echo "== bench fp32 rc=$?"; cat gpurun_out/bench_r2${T}_fp32.json | cut -c1-200
timeout 400 python bench.py --model T --steps 20 --warmup 5 > gpurun_out/bench_r2${T}_modelT.json 2> gpurun_out/bench_r2${T}_modelT.err; echo "== bench model T rc=$?"; cat gpurun_out/bench_r2${T}_modelT.json | cut -c1-260
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-grid41 --mode eager > gpurun_out/launches_r2$T.log 2>&1; echo "== ncu launch list rc=$?"; wc -l gpurun_out/launches_r2$T.csv
timeout 900 python bench.py --cpu-thread-sweep > gpurun_out/cpu_thread_sweep_r2.txt 2>&1; echo "== cpu sweep rc=$?"; tail -8 gpurun_out/cpu_thread_sweep_r2.txt
