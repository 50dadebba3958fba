#!/bin/bash
# round 2, call 14: final state -- GPU tests, default bench line, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke14.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke14.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t14.log; tail -3 gpurun_out/r2_t14.log
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b14.json 2> gpurun_out/r2_b14.err; echo "bench rc=$? in ${SECONDS}s"
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_b14_ref.json 2> gpurun_out/r2_b14_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b14.json","gpurun_out/r2_b14_ref.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "config", json.dumps(d["config"])[:120])
        if "configs" in d:
            r=d["roofline"]
            print("   parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "launches", d["gpu_launches"], "cpu", (d["cpu_baseline"] or {}).get("value"), "roof", round(r["frac"],3), r["frac_vs_theoretical_37p2"], r["issue_slot_util"], r["kernel_ms"], r["share_of_step"])
            print("   other", {k:(round(v["kernel_ms"]*1e3,1), round(v.get("frac",0),3)) for k,v in d["roofline_other"].items() if isinstance(v,dict)})
            for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v.get("matcher_us",0),1), round(v.get("matcher_tevals_per_s",0),3), round(v.get("frac",0),3), {a:round(x,1) for a,x in v.get("stage_us",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
