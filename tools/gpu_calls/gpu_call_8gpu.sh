#!/bin/bash
# round 2, 8-GPU call: what does the host side of this box absorb (bare D2H probe at 1/2/4/8 GPUs, allocation variants),
# and where does the end-to-end path sit against it (bench under torchrun at 8 and 2 GPUs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(nvidia-smi topo -m; echo; lscpu | grep -E "^CPU\(s\)|NUMA|Socket|Model name|Thread"; echo; grep MemTotal /sys/devices/system/node/node*/meminfo; free -g | head -2;
 for i in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi -i $i --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu$i $b numa=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) cpus=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"; done) > gpurun_out/r2_host8.txt 2>&1
P=gpurun_out/r2_probe8.json
timeout 300 python tools/d2h_probe.py --gpus 1,2,4,8 --seconds 1.0 > $P 2> gpurun_out/r2_probe8.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --mode wc >> $P 2>> gpurun_out/r2_probe8.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --mode registered >> $P 2>> gpurun_out/r2_probe8.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --numa interleave >> $P 2>> gpurun_out/r2_probe8.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --numa local >> $P 2>> gpurun_out/r2_probe8.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --with-h2d >> $P 2>> gpurun_out/r2_probe8.err
timeout 120 python tools/d2h_probe.py --gpus 8 --seconds 1.0 --chunk-mb 8 >> $P 2>> gpurun_out/r2_probe8.err
cat $P
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --table C5 \
      > gpurun_out/r2_b8_n$n.json 2> gpurun_out/r2_b8_n$n.err; echo "bench n=$n rc=$?"
done
B200S_BENCH_WC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-check --table '' \
      > gpurun_out/r2_b8_n8_wc.json 2> gpurun_out/r2_b8_n8_wc.err; echo "bench wc rc=$?"
for f in gpurun_out/r2_b8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n", d["n_gpus"], "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceil", round(d["e2e"]["copy_ceiling_frames_per_s"]), "frac", round(d["e2e"]["frac_of_copy_ceiling"],3), "parity", d["parity_checked"]["mismatches"])
    for k,v in d["configs"].items(): print("   ",k, round(v["frames_per_s"]), round(v["e2e_frames_per_s"]), round(v["e2e"]["frac_of_copy_ceiling"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
