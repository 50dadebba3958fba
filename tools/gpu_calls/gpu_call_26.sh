#!/bin/bash
# round 2, call 26: ncu evidence for the kernels as shipped -- launch list of a C4 run, full captures of rectify (quad, two-row Sobel) and pack (table path, pipelined)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --config C4 --steps 2 --warmup 1 --no-cpu --no-check --table ''"
timeout 200 bash -c "$B" > gpurun_out/r2_c26_plain.json 2> gpurun_out/r2_c26_plain.err; echo "plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c4_shipped.csv bash -c "$B" > gpurun_out/r2_c26_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'pack_lut|rectify_xsobel_quad' --launch-skip 10 -c 2 -f -o gpurun_out/r2_shipped_small bash -c "$B" > gpurun_out/r2_c26_ncu2.log 2>&1; echo "small kernels rc=$?"
ls -la gpurun_out/r2_shipped_small.ncu-rep gpurun_out/r2_launches_c4_shipped.csv
