#!/bin/bash
# round 2, call 9: smoke(), full GPU tests, default bench run, reference arm, ncu launch list + full capture of the C4 chain
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke9.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke9.log; tail -2 gpurun_out/r2_smoke9.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t9.log
tail -4 gpurun_out/r2_t9.log
/usr/bin/time -v timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b9.json 2> gpurun_out/r2_b9.err; echo "bench rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/r2_b9.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_b9_ref.json 2> gpurun_out/r2_b9_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches9_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_ncu_list9.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bm_vh_kernel|rectify_xsobel|reproject_pack" -s 9 -c 3 -o gpurun_out/r2_c4_chain9 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-check --table '' > gpurun_out/r2_ncu_chain9.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2_b9.json","gpurun_out/r2_b9_ref.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
        if "configs" in d:
            print("   parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"], "roof", round(d["roofline"]["frac"],3))
            for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v.get("matcher_us",0),1), round(v.get("matcher_tevals_per_s",0),3), round(v.get("frac",0),3))
    except Exception as e: print(f, "ERR", e)
PY
