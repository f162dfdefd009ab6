#!/bin/bash
# GPU-box command: tests of the new paths, then the full bench (all BASELINE configs), then a summary line
python -m pytest tests/test_gpu_packed.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -30
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_b1.json 2> gpurun_out/r02_b1.err
tail -c 1500 gpurun_out/r02_b1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_b1.json"))
def show(r, name):
    print(name, "value %.3e" % r["value"], "ms/step %.3f" % r["ms_per_step"], "kernel_ms %.3f" % r["roofline"].get("kernel_ms", 0),
          "e2e %.3e" % (r["e2e"]["value"] if r.get("e2e") else 0), r["parity"], "fit", r["config"].get("fit_ms"),
          "state", r["config"].get("score_state_ms"), "bo_iter", r["config"].get("bo_iter_ms"))
show(d, "cfg2")
for k, v in d["configs"].items():
    if "error" in v:
        print(k, v)
    else:
        show(v, k)
PY
