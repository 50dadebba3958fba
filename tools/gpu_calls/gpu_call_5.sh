#!/bin/bash
# round 2, call 5: do the small kernels hide under the matcher when their blocks can share its SMs?  (128-thread rectify /
# pack blocks, matcher at 72 registers) -- A/B of three builds of the library on C4 and C5, plus parity of the variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for sfx in "" _co _co72; do
  B200S_LIB_SUFFIX=$sfx timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b5_C4${sfx}.json 2> gpurun_out/r2_b5_C4${sfx}.err
  B200S_LIB_SUFFIX=$sfx timeout 300 python bench.py --config C5 --steps 4 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b5_C5${sfx}.json 2> gpurun_out/r2_b5_C5${sfx}.err
  B200S_LIB_SUFFIX=$sfx timeout 300 python bench.py --config C3 --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b5_C3${sfx}.json 2> gpurun_out/r2_b5_C3${sfx}.err
done
B200S_LIB_SUFFIX=_co72 timeout 900 python -m pytest tests -m gpu -x -q -k "baseline_configs or fused_chain or bench_configuration or vh_kernel or ragged" > gpurun_out/r2_t5_co72.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5_co72.log
tail -3 gpurun_out/r2_t5_co72.log
for f in gpurun_out/r2_b5_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=list(d["configs"])[0]; v=d["configs"][k]
    print(sys.argv[1], "fps", round(v["frames_per_s"]), "e2e", round(v["e2e_frames_per_s"]), "matcher us", round(v["matcher_us"],1), {a:round(x,1) for a,x in v["stage_us"].items()}, v["parity_checked"]["mismatches"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
