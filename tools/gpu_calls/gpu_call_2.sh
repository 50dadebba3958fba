#!/bin/bash
# round 2, call 2: refactored library (batches, colour, async senders, map formats): full GPU tests, bench with the configs
# table, A/B of pack mode and on-the-fly rectification, batch-size sweep of the small shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -30 gpurun_out/r2_t2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo "bench rc=$?"
B200S_PACK_DIRECT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b2_direct.json 2> gpurun_out/r2_b2_direct.err
B200S_BENCH_RECT_FLY=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b2_fly.json 2> gpurun_out/r2_b2_fly.err
B200S_MAP_ABS32=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b2_abs32.json 2> gpurun_out/r2_b2_abs32.err
for cfg in C1 C2 C3; do
  for b in 1 2 4 8 16; do
    B200S_BENCH_BATCH=$b B200S_VH_VERBOSE=0 timeout 300 python bench.py --config $cfg --steps 6 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_sweep_${cfg}_b${b}.json 2> gpurun_out/r2_sweep_${cfg}_b${b}.err
  done
done
for b in 2 4; do
  B200S_BENCH_BATCH=$b timeout 300 python bench.py --config C4 --steps 6 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_sweep_C4_b${b}.json 2> gpurun_out/r2_sweep_C4_b${b}.err
done
ls gpurun_out | head -50
