#!/bin/bash
# round 2, call 3: address-table inputs + fused float plane: GPU tests, bench with the configs table, batch sweeps,
# launch list and full ncu captures of the kernels around the matcher
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; echo "bench rc=$?"
for cfg in C1 C2; do
  for b in 4 8 16; do
    B200S_BENCH_BATCH=$b timeout 300 python bench.py --config $cfg --steps 6 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_sw3_${cfg}_b${b}.json 2> gpurun_out/r2_sw3_${cfg}_b${b}.err
  done
done
for b in 2 4 8; do
  B200S_BENCH_BATCH=$b timeout 300 python bench.py --config C3 --steps 6 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_sw3_C3_b${b}.json 2> gpurun_out/r2_sw3_C3_b${b}.err
done
for b in 2 4; do
  B200S_BENCH_BATCH=$b timeout 300 python bench.py --config C4 --steps 6 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_sw3_C4_b${b}.json 2> gpurun_out/r2_sw3_C4_b${b}.err
done
timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b3_c4r.json 2> gpurun_out/r2_b3_c4r.err
# launch list (cold, serialised) of one short C4 run, then full captures of the non-matcher kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rectify_xsobel|reproject_pack|fill_border" -s 8 -c 6 -o gpurun_out/r2_small_kernels -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_ncu_small.log 2>&1
ls -la gpurun_out | tail -30
