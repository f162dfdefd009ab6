"""cfg5 on one GPU: row-block size of the symmetric schedule (the whole 40 GB u8 matrix is materialised every time)."""
import json, sys
sys.path.insert(0, '.')
import torch
from nasbowl_b200.engine import Engine
eng = Engine.get()
N = 200000
cells = eng.generate_cells(0, 2024, 0, N)
fit = eng.wl_fit(cells, 8, 3, False)
D = fit.k_end(3)
out = {}
for block in (1024, 2048, 3072, 4096, 8192):
    ts = []
    for rep in range(2):
        r = eng.gram_stress(fit, D, block)
        ts.append(r["ms"])
        ok = all(r["checks"].values())
        del r
        torch.cuda.empty_cache()
    out[str(block)] = {"ms": ts, "checks_ok": ok}
    print(block, ts, ok, file=sys.stderr)
print(json.dumps(out))
