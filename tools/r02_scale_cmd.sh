mkdir -p gpurun_out
for N in $NS; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
  echo "N=$N rc=$?"
  python - <<PY
import json
t=open("gpurun_out/r02_scale_n$N.json").read()
if '{"metric"' in t:
    d=json.loads(t[t.index('{"metric"'):].splitlines()[0])
    print(d["n_gpus"], round(d["value"]/1e9,3), "Gpts/s", round(d["ms_per_step"],4), "ms/step", d["stage_ms"], "e2e", round(d["e2e"]["ms_per_step"],3), d.get("identity_ok"), d.get("exchange","")[:50])
    print([round(sum(r)/len(r),4) for r in d.get("step_ms_per_rank",[])])
else:
    print("no line"); print(open("gpurun_out/r02_scale_n$N.err").read()[-1500:])
PY
done
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $PN --master-addr 127.0.0.1 --master-port 29533 tools/r02_stage_probe.py 2>&1 | grep rank
