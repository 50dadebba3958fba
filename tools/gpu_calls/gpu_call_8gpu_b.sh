#!/bin/bash
# round 2, second 8-GPU call: why does the end-to-end path reach 93 GB/s at 8 GPUs when bare copies reach 152?
# probe with 1 / 4 copy streams and a larger working set, bench with 1 / 2 / 8 frame slots (= concurrent D2H streams per GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
P=gpurun_out/r2_probe8b.json
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --streams 1 > $P 2> gpurun_out/r2_probe8b.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --streams 4 --buffers 8 >> $P 2>> gpurun_out/r2_probe8b.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --streams 2 --buffers 16 >> $P 2>> gpurun_out/r2_probe8b.err
cat $P
for s in 2 1 8; do
  B200S_BENCH_SLOTS=$s timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$s bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu --no-check --table '' \
      > gpurun_out/r2_b8b_slots$s.json 2> gpurun_out/r2_b8b_slots$s.err; echo "bench slots=$s rc=$?"
done
for f in gpurun_out/r2_b8b_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n", d["n_gpus"], "slots", d["slots"], "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceil", round(d["e2e"]["copy_ceiling_frames_per_s"]), "frac", round(d["e2e"]["frac_of_copy_ceiling"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
