#!/bin/bash
# round 2, third 8-GPU call: the driver's scaling sequence with the default bench line (all configs in the table) at 8 and 4 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for n in 8 4; do
  SECONDS=0
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 20 --warmup 5 \
      > gpurun_out/r2_b8c_n$n.json 2> gpurun_out/r2_b8c_n$n.err; echo "bench n=$n rc=$? in ${SECONDS}s"
done
python - <<'PY'
import json
for f in ("gpurun_out/r2_b8c_n8.json","gpurun_out/r2_b8c_n4.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d["n_gpus"], "fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_checked"]["frames"], d["parity_checked"]["mismatches"], "ceil", round(d["e2e"]["copy_ceiling_frames_per_s"]), round(d["e2e"]["copy_ceiling_equal_shares_frames_per_s"]))
        for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3), round(v["e2e"]["frac_of_equal_shares_ceiling"],3))
    except Exception as e: print(f, "ERR", e)
PY
