"""Developer tool (GPU box): run a matrix of attention cases through the public op and print
per-output error metrics against the fp64 oracle.  Cases run in a child process; if a case kills
the CUDA context (trap / illegal address) the parent records it and restarts after it.

    python tools/gpu_check.py [--set smoke|full] [--out gpurun_out/check.jsonl]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name, B, H, M, N, D, dtype, bias kind (None|'1H'|'BH'|'11'|'B1'), causal, sm_scale, layout, bwd
def case_sets():
    smoke = [
        ("d64_nobias", 1, 1, 128, 128, 64, "bf16", None, False, 0.125, "bhsd", True),
        ("d64_bias1H", 2, 2, 128, 128, 64, "bf16", "1H", False, 1.0, "bhsd", True),
        ("d64_2tiles", 1, 2, 256, 256, 64, "bf16", "1H", False, 1.0, "bshd", True),
        ("d64_causal", 2, 2, 256, 256, 64, "bf16", "1H", True, 1.0, "bshd", True),
        ("d64_ragged", 2, 3, 200, 328, 64, "bf16", "1H", False, 1.0, "bshd", True),
        ("d64_fp16", 2, 2, 256, 256, 64, "fp16", "1H", False, 1.0, "bshd", True),
        ("d128", 1, 2, 256, 256, 128, "bf16", "1H", False, 1.0, "bshd", True),
        ("d32", 1, 2, 256, 256, 32, "bf16", "1H", False, 1.0, "bshd", True),
        ("d16", 1, 2, 256, 256, 16, "bf16", "1H", False, 1.0, "bshd", True),
        ("d64_mode2", 2, 2, 130, 131, 64, "bf16", "1H", True, 1.0, "bshd", True),
        ("d64_biasBH", 2, 2, 256, 256, 64, "bf16", "BH", True, 1.0, "bshd", True),
        ("d64_bias11", 2, 2, 256, 256, 64, "bf16", "11", False, 1.0, "bshd", True),
        ("d64_mgtn", 1, 2, 384, 200, 64, "bf16", "1H", True, 1.0, "bshd", True),
        ("d64_s1024", 2, 4, 1024, 1024, 64, "bf16", "1H", False, 1.0, "bshd", True),
    ]
    full = smoke + [
        ("ref_test_d128", 2, 4, 512, 612, 128, "fp16", "BH", True, 1.0, "bhsd", True),
        ("ref_test_d64", 2, 4, 1024, 1045, 64, "bf16", "11", False, 1.0, "bhsd", True),
        ("d128_causal", 2, 2, 512, 512, 128, "bf16", "1H", True, 1.0, "bshd", True),
        ("d32_causal_rag", 2, 2, 300, 333, 32, "fp16", "1H", True, 0.5, "bshd", True),
        ("d16_causal_rag", 2, 2, 300, 333, 16, "bf16", "B1", True, 0.5, "bshd", True),
        ("cross_nobias", 2, 4, 512, 384, 64, "bf16", None, False, 1.0, "bshd", True),
    ]
    return {"smoke": smoke, "full": full}


def run_case(spec):
    import torch
    from oracle import attn_bias_ref as orc
    from flasht5_b200 import flash_attention_v2_bias
    name, B, H, M, N, D, dt, bk, causal, scale, layout, bwd = spec
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    g = torch.Generator().manual_seed(1234)

    def mk(b, h, s, d):
        if layout == "bshd":
            return torch.randn(b, s, h, d, generator=g).to(dtype).permute(0, 2, 1, 3)
        return torch.randn(b, h, s, d, generator=g).to(dtype)
    q, k, v, do = mk(B, H, M, D), mk(B, H, N, D), mk(B, H, N, D), mk(B, H, M, D)
    bias = None
    if bk is not None:
        shape = {"1H": (1, H, M, N), "BH": (B, H, M, N), "11": (1, 1, M, N), "B1": (B, 1, M, N)}[bk]
        bias = torch.randn(*shape, generator=g).to(dtype)
    if causal and M > N:   # rows with no visible key: dO there must not matter
        pass
    # oracle (fp64 on CPU)
    o_ref, L_ref, dq_ref, dk_ref, dv_ref, db_ref = orc.attn_fwd_bwd(q.float(), k.float(), v.float(),
                                                                      None if bias is None else bias.float(),
                                                                      do.float(), causal, scale)
    dev = torch.device("cuda:0")
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    bd = bias.to(dev).requires_grad_(True) if bias is not None else None
    res = {"name": name, "spec": spec[1:]}
    t0 = time.time()
    o = flash_attention_v2_bias(qd, kd, vd, bd, causal, scale)
    torch.cuda.synchronize()
    res["o"] = orc.error_metrics(o, o_ref)
    # LSE through the raw op
    _, L = torch.ops.b200t5.attn_bias_fwd(qd.detach(), kd.detach(), vd.detach(), None if bd is None else bd.detach(), causal, float(scale))
    torch.cuda.synchronize()
    res["lse"] = orc.error_metrics(L, L_ref)
    if bwd:
        ins = [qd, kd, vd] + ([bd] if bd is not None else [])
        grads = torch.autograd.grad(o, ins, do.to(dev))
        torch.cuda.synchronize()
        res["dq"] = orc.error_metrics(grads[0], dq_ref)
        res["dk"] = orc.error_metrics(grads[1], dk_ref)
        res["dv"] = orc.error_metrics(grads[2], dv_ref)
        if bd is not None:
            res["dbias"] = orc.error_metrics(grads[3], db_ref)
    res["sec"] = round(time.time() - t0, 3)
    # eager low-precision error (the reference's tolerance yardstick) for o only
    o_low = orc.attn_eager_lowp(q, k, v, bias, causal, scale)
    res["o_eager"] = orc.error_metrics(o_low, o_ref)
    return res


def child(specs_json, out_path):
    specs = json.loads(specs_json)
    import torch  # noqa: F401
    for spec in specs:
        print("BEGIN", spec[0], flush=True)
        try:
            r = run_case(tuple(spec))
            r["status"] = "ok"
        except Exception as e:   # noqa: BLE001
            r = {"name": spec[0], "spec": spec[1:], "status": "error", "error": repr(e)[:400]}
            with open(out_path, "a") as f:
                f.write(json.dumps(r) + "\n")
            print("END", spec[0], "error", r["error"], flush=True)
            if "CUDA" in r["error"] or "cuda" in r["error"]:
                sys.exit(3)          # context is probably dead
            continue
        with open(out_path, "a") as f:
            f.write(json.dumps(r) + "\n")
        print("END", spec[0], "ok", flush=True)


def fmt(x):
    return "%9.2e/%8.2e" % tuple(x) if x else " " * 18


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="smoke")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "check.jsonl"))
    ap.add_argument("--child", default=None)
    ap.add_argument("--only", default=None)
    ap.add_argument("--case-timeout", type=int, default=240)
    a = ap.parse_args()
    if a.child is not None:
        child(a.child, a.out)
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").close()
    specs = [list(s) for s in case_sets()[a.set]]
    if a.only:
        specs = [s for s in specs if s[0] in a.only.split(",")]
    remaining = specs
    while remaining:
        p = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", json.dumps(remaining), "--out", a.out],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        try:
            out, _ = p.communicate(timeout=a.case_timeout * max(1, len(remaining)))
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        begun = [ln.split()[1] for ln in out.splitlines() if ln.startswith("BEGIN")]
        ended = [ln.split()[1] for ln in out.splitlines() if ln.startswith("END")]
        tail = "\n".join(out.splitlines()[-15:])
        if p.returncode == 0:
            break
        crashed = [n for n in begun if n not in ended]
        names = [s[0] for s in remaining]
        if crashed:
            with open(a.out, "a") as f:
                f.write(json.dumps({"name": crashed[0], "status": "crash", "error": tail[-1500:]}) + "\n")
            idx = names.index(crashed[0]) + 1
        elif ended:
            idx = names.index(ended[-1]) + 1
        else:
            print("child failed before any case:\n" + tail)
            break
        remaining = remaining[idx:]
    print("%-16s %-6s %-18s %-18s %-18s %-18s %-18s %-18s %-18s" % ("case", "status", "o max/relF", "lse", "dq", "dk", "dv", "dbias", "o_eager"))
    for ln in open(a.out):
        r = json.loads(ln)
        if r["status"] != "ok":
            print("%-16s %-6s %s" % (r["name"], r["status"], r.get("error", "")[-600:]))
            continue
        print("%-16s %-6s %s %s %s %s %s %s %s" % (r["name"], r["status"], fmt(r.get("o")), fmt(r.get("lse")), fmt(r.get("dq")),
                                              fmt(r.get("dk")), fmt(r.get("dv")), fmt(r.get("dbias")), fmt(r.get("o_eager"))))


if __name__ == "__main__":
    main()
