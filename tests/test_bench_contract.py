"""bench.py's reference arm (CPU only) honours the JSON contract; the B200 arm needs a GPU and is
covered by the gpu-marked test below."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]


def _run(args, env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--ref-sample", "300"],
             env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "candidates/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    # the loop the reference's propose_location really runs (one eval per candidate) is timed beside the batched form
    pc = d["cpu_baseline"]["propose_location_path"]
    assert pc["unit"] == "candidates/s" and 0 < pc["value"] < d["value"]
    # the CPU arm names exactly the configuration the B200 arm runs at this N (the static part of `config`)
    import bench
    w = bench.WORKLOADS["cfg2"]
    want = bench.workload_config(w, w["pool"], w["pool"], d["config"]["h_star"])
    assert {k: d["config"][k] for k in want} == want
    assert "1000000-candidate EI pool per GPU (1000000 in total)" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", CUDA_VISIBLE_DEVICES=""))
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_json_line():
    d = _run(["--gpus", "1", "--steps", "5", "--warmup", "3", "--cpu-sample", "1000",
              "--configs", "cfg3,cfg5", "--gram-cells", "8192"])
    for k in REQUIRED + ["roofline", "clocks", "parity", "configs"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["value"] > 1e7 and d["gpu_launches"] >= 5
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    if r["bound"] == "issue":
        # the scoring kernel is bound by instruction issue: warp instructions per candidate come from the committed
        # ncu counters, the HBM figures of SURVEY 8(d) ride along
        assert r["unit"] == "Gwarp-inst/s" and r["warp_inst_per_candidate"] > 0 and r["traffic"] > 0
        assert r["hbm"]["unit"] == "GB/s" and r["hbm"]["bytes_per_unit"] == 36
    else:
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # NB pools travel as 8-byte wire records; the 16-byte form is measured beside them
    assert d["e2e"]["h2d_bytes_per_step"] == 1000000 * 8 and d["e2e"]["d2h_bytes_per_step"] >= 1000000 * 4
    assert d["e2e"]["full_records"]["h2d_bytes_per_step"] == 1000000 * 16
    assert d["parity"]["max_rel_err_ei_fp64"] < 1e-6 and d["parity"]["h_star_equal"] and d["parity"]["ok"]
    c3, c5 = d["configs"]["cfg3"], d["configs"]["cfg5"]
    assert c3["parity"]["ok"] and c3["scaling"] == "strong" and c3["roofline"]["bound"] in ("issue", "hbm")
    assert c5["parity"]["ok"] and c5["roofline"]["bound"] == "tensor"
