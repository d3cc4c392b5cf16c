"""Scoring kernel K8 (fused in-batch debiased CE forward): algorithmic HBM GB/s and TFLOP/s at the cfg-3 local shape
and at the G=8 all-gathered shape.  Algorithmic bytes = s*(R+C)*D + 12*C + 12*R  (operands, ids/log-pop
as the bit-matrix inputs, log_mask, row outputs); the [R,C] logits are never written."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from idvs.morec_b200 import lib
from idvs.morec_b200.synth import synth_batch

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
out = []
B, L, D = 64, 25, 512
for G in (1, 8):
    bs = [synth_batch(B, L, 50000, 0, seed=100 + g, modal=False, n_users_pop=2000) for g in range(G)]
    ids_all = torch.cat([b["ids"].reshape(-1) for b in bs]).cuda()
    ids_loc = bs[0]["ids"].cuda()
    lm = bs[0]["log_mask"].reshape(-1).cuda()
    R, C = B * L, ids_all.numel()
    member, pad = lib.inbatch_mask(ids_loc, ids_all, B, L)
    logp = torch.rand(C, device="cuda").log()
    for dt, x3, name, s in ((torch.float32, True, "fp32 (3xTF32)", 4), (torch.float32, False, "tf32", 4), (torch.bfloat16, False, "bf16", 2)):
        P = (torch.randn(R, D, device="cuda") * 0.3).to(dt)
        E = (torch.randn(C, D, device="cuda") * 0.3).to(dt)
        with lib.fp32_mode(x3):
            for _ in range(5):
                lib.inbatch_ce_fwd(P, E, member, pad, logp, lm, B, L)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50
            e0.record()
            for _ in range(n):
                lib.inbatch_ce_fwd(P, E, member, pad, logp, lm, B, L)
            e1.record()
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        byts = s * (R + C) * D + 12 * C + 12 * R
        flops = 2.0 * R * C * D
        rec = dict(shape=f"R={R} C={C} D={D} (G={G})", mode=name, us=round(us, 2), alg_bytes=byts,
                   gbs=round(byts / us / 1e3, 1), frac_hbm=round(byts / us / 1e3 / peaks["hbm_gbs"], 4),
                   tflops=round(flops / us / 1e6, 1))
        out.append(rec)
        print(rec)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scoring_kernel_r01.json"), "w"), indent=1)
