#!/bin/bash
# GPU-box command: the full bench (all BASELINE configs) and a one-line summary per config
N=${1:-1}
OUT=gpurun_out/r02_bench_${N}gpu
if [ "$N" = "1" ]; then
  python bench.py --steps 20 --warmup 5 > $OUT.json 2> $OUT.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > $OUT.json 2> $OUT.err
fi
tail -c 1200 $OUT.err
python - $OUT.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(r, name):
    c = r["config"]
    print(name, "value %.3e" % r["value"], "ms/step %.3f" % r["ms_per_step"], "kernel_ms %.3f" % r["roofline"].get("kernel_ms", 0),
          "e2e %.3e" % (r["e2e"]["value"] if r.get("e2e") else 0), "fit", c.get("fit_ms"), "state", c.get("score_state_ms"),
          "bcast", c.get("broadcast_ms"), c.get("broadcast_cold_ms"), "bo_iter", c.get("bo_iter_ms"), c.get("bo_iter_from_scratch_ms"), c.get("bo_iter_appended"))
    print("   parity", r["parity"])
show(d, "cfg2")
for k, v in d["configs"].items():
    if v is None or "error" in v:
        print(k, v)
    else:
        show(v, k)
PY
