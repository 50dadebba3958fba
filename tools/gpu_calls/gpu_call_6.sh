#!/bin/bash
# round 2, call 6: standard-Q pack kernel (4 rows per warp), kept disparity border, C probe inside bench: GPU tests + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log
tail -5 gpurun_out/r2_t6.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b6.json 2> gpurun_out/r2_b6.err; echo "bench rc=$?"
timeout 300 python bench.py --config C4r --steps 6 --warmup 3 --no-cpu --table '' > gpurun_out/r2_b6_c4r.json 2> gpurun_out/r2_b6_c4r.err
timeout 120 python tools/d2h_probe.py --gpus 1 --seconds 1.0 --streams 1 > gpurun_out/r2_probe6.json 2> gpurun_out/r2_probe6.err
timeout 120 python tools/d2h_probe.py --gpus 1 --seconds 1.0 --streams 4 --buffers 8 >> gpurun_out/r2_probe6.json 2>> gpurun_out/r2_probe6.err
cat gpurun_out/r2_probe6.json
python - <<'PY'
import json
for f in ("gpurun_out/r2_b6.json","gpurun_out/r2_b6_c4r.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceil", round(d["e2e"]["copy_ceiling_frames_per_s"]), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "launches", d["gpu_launches"])
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
