#!/bin/bash
mkdir -p gpurun_out
for c in c1 c5-5 c5-7 c5-11 c5-13; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$c.log 2>&1
  echo "$c: $(tail -c 4000 gpurun_out/bench_$c.log | grep -o '"value": [0-9.]*' | head -1) $(tail -c 4000 gpurun_out/bench_$c.log | grep -o '"ms_per_move_by_kernel": {[^}]*}') $(grep -i error gpurun_out/bench_$c.log | tail -1 | cut -c1-200)"
done
timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c3.log 2>&1
echo "c3: $(tail -c 4000 gpurun_out/bench_c3.log | grep -o '"value": [0-9.]*' | head -1) $(tail -c 4000 gpurun_out/bench_c3.log | grep -o '"ms_per_move_by_kernel": {[^}]*}') $(grep -i error gpurun_out/bench_c3.log | tail -1 | cut -c1-200)"
timeout 300 python tools/descend_phases.py c2 > gpurun_out/phases.log 2>&1
