"""Developer tool: distil gpurun_out/triton_parity.json (written on a B200 by tools/triton_parity.py) into the committed
fixture tests/golden/triton_b200_errors.json that tests/test_triton_error_parity.py compares the CUDA kernels with.
    python tools/make_triton_fixture.py [gpurun_out/triton_parity.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "triton_parity.json")
KEEP = ["c2", "reftest_d128", "reftest_d128_causal", "reftest_d64", "reftest_d64_causal", "reftest_d64_bias1H"]

d = json.load(open(SRC))
out = {
    "provenance": "Errors of the REFERENCE Triton kernels (src/model/ops/flash_attention_v2_bias.py, staged by tools/stage_reference.sh) "
                  "against an fp64 evaluation of the same formula, measured on %s with torch %s / triton %s by tools/triton_parity.py "
                  "(inputs: make_inputs(case, dtype, seed=1234)); rel_f = ||x - ref||_F / ||ref||_F.  triton_def = the configuration the "
                  "reference ships for compute capability 10.0 (32x32 tiles, 1 stage, 4 warps); triton_a100 = its A100 tile table forced.  "
                  "Regenerate: tools/stage_reference.sh && python tools/triton_parity.py && python tools/make_triton_fixture.py"
                  % (d["device"], d["torch"], d.get("triton")),
    "cases": [],
}
for r in d["cases"]:
    if r["case"] not in KEEP:
        continue
    e = r["err_vs_fp64"]
    rec = {k: r[k] for k in ("case", "dtype", "shape", "bias", "causal", "sm_scale")}
    for impl in ("triton_def", "triton_a100", "eager_lowp", "new"):
        rec[impl] = {t: (round(v["rel_f"], 9) if v else None) for t, v in e[impl].items()} if e.get(impl) else None
    out["cases"].append(rec)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "triton_b200_errors.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
