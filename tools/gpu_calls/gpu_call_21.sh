#!/bin/bash
# round 2, call 21: ncu evidence for the final kernels -- launch list of a C4 run, full captures of the small kernels and the strip kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --config C4 --steps 2 --warmup 1 --no-cpu --no-check --table ''"
timeout 200 bash -c "$B" > gpurun_out/r2_c21_plain.json 2> gpurun_out/r2_c21_plain.err; echo "plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c4_final.csv bash -c "$B" > gpurun_out/r2_c21_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'reproject_pack|rectify_xsobel_quad|reproject_lut' --launch-skip 10 -c 4 -f -o gpurun_out/r2_final_small bash -c "$B" > gpurun_out/r2_c21_ncu2.log 2>&1; echo "small kernels rc=$?"
DISP12=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bm_strip' --launch-skip 3 -c 1 -f -o gpurun_out/r2_final_strip python tools/time_bm.py C4 3 4 > gpurun_out/r2_c21_ncu3.log 2>&1; echo "strip rc=$?"
ls -la gpurun_out/r2_final_*.ncu-rep gpurun_out/r2_launches_c4_final.csv
