#!/bin/bash
# run inside gpurun --gpus 8: scaling sweep N=1,2,4,8
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print($n, "value %.3f Gpts/s" % (d["value"]/1e9), "ms %.3f" % d["ms_per_step"], d["stage_ms"], "e2e %.3f" % (d["e2e"]["value"]/1e9), d.get("exchange"), d["clocks"]["reasons"])
except Exception as e:
    print($n, "FAILED", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
