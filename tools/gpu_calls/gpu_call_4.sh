#!/bin/bash
# round 2, call 4: fused border strips, exact fast division in the pack kernel, shared-memory source window in the rectify
# kernel, misc-buffer fix: GPU tests, bench, C4r, matcher tuning data for C1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -5 gpurun_out/r2_t4.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; echo "bench rc=$?"
timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b4_c4r.json 2> gpurun_out/r2_b4_c4r.err
B200S_STRIPS=0 timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b4_c4r_nostrips.json 2> gpurun_out/r2_b4_c4r_nostrips.err
(for st in 2 3 4; do for ncb in 4 6 8; do for bands in 2 3 4 6; do
  B200S_STAGERS=$st B200S_VH_NCB=$ncb B200S_VH_BANDS=$bands B200S_VH_VERBOSE=1 timeout 120 python tools/time_bm.py C1 5 16 2>&1 | grep -E "plan|bm " | sort -u | sed 's/.*grid=/grid=/; s/.*: bm/bm/' | tr '\n' ' '; echo " [st=$st ncb=$ncb bands=$bands]"
done; done; done) > gpurun_out/r2_c1_sweep.log 2>&1
DISP12=0 timeout 120 python tools/time_bm.py C4 10 1 > gpurun_out/r2_c4_disp12.log 2>&1
B200S_STRIPS=0 DISP12=0 timeout 120 python tools/time_bm.py C4 10 1 >> gpurun_out/r2_c4_disp12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bm_vh_kernel" -s 2 -c 1 -o gpurun_out/r2_c1_vh -f \
    python tools/time_bm.py C1 2 16 > gpurun_out/r2_ncu_c1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rectify_xsobel|reproject_pack|bm_strip" -s 8 -c 6 -o gpurun_out/r2_small_kernels2 -f \
    python bench.py --config C4r --steps 1 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_ncu_small2.log 2>&1
ls -la gpurun_out | tail -12
