"""Measured deviation of every compute mode from the CPU oracle at the real configurations (gpurun).

    python tools/parity_report.py [--out gpurun_out/parity_modes.json] [--cases bert_base_b4,...] [--modes fp32,tf32,bf16]

For each case x mode: |loss - oracle|, max |score_embs - oracle| on non-pad slots, the relative L2 error of the whole
gradient vector, the worst per-tensor relative L2 error / max-abs-relative error and the gradient cosine."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import real_cases as RC  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_modes.json"))
    ap.add_argument("--cases", default="bert_base_b4,bert_tiny_t128_b16,swin_t_b2")
    ap.add_argument("--modes", default="fp32,tf32,bf16,fp16")
    a = ap.parse_args()
    rep = {}
    for name in a.cases.split(","):
        c = RC.CASES[name]
        d = RC.build_inputs(c)
        if c["kind"] == "text":
            from idvs.morec_b200.model import Model
        else:
            from idvs.morec_b200.model_vision import Model
        model = RC.build_model(c, Model, d["pop_prob"])
        out, gref = RC.run_oracle(c, model, d)
        model = model.cuda().eval()
        nonpad = d["ids"].reshape(-1) != 0
        rep[name] = {}
        for mode in a.modes.split(","):
            try:
                model.set_compute_dtype(mode)
            except AssertionError:
                continue
            cap = {}
            orig = model._encode_items
            model._encode_items = lambda i, x, h=None: cap.setdefault("E", orig(i, x, h))
            model.zero_grad()
            loss = model(d["ids"].reshape(-1).cuda(), d["items"].cuda(), d["log_mask"].cuda(), 0)
            loss.backward()
            model._encode_items = orig
            E = cap["E"].detach().float().cpu()
            num = da = db = 0.0
            worst_l2, worst_max = ("", 0.0), ("", 0.0)
            per = []
            for k, gr in gref.items():
                if "pooler" in k or RC.is_null_gradient(k):
                    continue
                g = dict(model.named_parameters())[k].grad.detach().float().cpu().double()
                gr = gr.double()
                diff = g - gr
                num += float((g * gr).sum()); da += float((g * g).sum()); db += float((gr * gr).sum())
                l2 = float(diff.norm()) / (float(gr.norm()) + 1e-30)
                mx = float(diff.abs().max()) / (float(gr.abs().max()) + 1e-30)
                per.append((round(l2, 5), k))
                if l2 > worst_l2[1]:
                    worst_l2 = (k, l2)
                if mx > worst_max[1]:
                    worst_max = (k, mx)
            tot = sum(float(((dict(model.named_parameters())[k].grad.detach().float().cpu().double() - gr.double()) ** 2).sum())
                      for k, gr in gref.items() if "pooler" not in k and not RC.is_null_gradient(k))
            rep[name][mode] = dict(loss=float(loss), loss_oracle=float(out.loss), loss_err=abs(float(loss) - float(out.loss)),
                                   emb_max_err=float((E[nonpad] - out.score_embs.detach()[nonpad]).abs().max()),
                                   emb_absmax=float(out.score_embs.detach().abs().max()),
                                   grad_rel_l2=(tot ** 0.5) / (db ** 0.5 + 1e-30), grad_cos=num / ((da * db) ** 0.5 + 1e-30),
                                   worst_tensor_rel_l2=worst_l2, worst_tensor_rel_max=worst_max,
                                   worst8=sorted(per, reverse=True)[:8])
            print(name, mode, json.dumps(rep[name][mode]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
