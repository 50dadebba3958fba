#!/bin/bash
# round 2, call 10: halo exchange between the VH threads (nd 64 / 128), pack kernel with prefetched rows: tests, A/B, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t10.log
tail -4 gpurun_out/r2_t10.log
timeout 600 python tools/fuzz_parity.py 120 99 > gpurun_out/r2_fuzz10.log 2>&1; tail -2 gpurun_out/r2_fuzz10.log
(for cfg in "C1 16" "C2 16" "C3 8" "C1 1" "C3 1"; do set -- $cfg
   for hx in 1 0; do B200S_VH_HX=$hx B200S_VH_VERBOSE=1 timeout 120 python tools/time_bm.py $1 10 $2 2>&1 | grep -E "plan|bm " | sort -u | sed 's/.*grid=/grid=/; s/stages.*//' | tr '\n' ' '; echo " [hx=$hx]"; done
 done) > gpurun_out/r2_hx10.log 2>&1
cat gpurun_out/r2_hx10.log
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err; echo "bench rc=$? in ${SECONDS}s"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b10.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"], "roof", round(d["roofline"]["frac"],3))
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v.get("matcher_us",0),1), round(v.get("matcher_tevals_per_s",0),3), round(v.get("frac",0),3), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
