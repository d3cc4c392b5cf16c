"""Bisect the bf16-mode gradient deviation on the text_tiny golden: compare intermediate tensors / grads per mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_parity_gpu as T
from idvs.morec_b200 import ops

g = torch.load(os.path.join(ROOT, "tests", "golden", "text_tiny.pt"), map_location="cpu", weights_only=False)
res = {}
for mode in ("fp32", "bf16"):
    model = T.build_model(g)
    model.set_compute_dtype(mode)
    cap = {}
    orig_enc = model._encode_items
    def enc(i, x, h=None, orig_enc=orig_enc, cap=cap):
        e = orig_enc(i, x, h)
        cap["E"] = e.detach().float().cpu()
        e.register_hook(lambda gr: cap.__setitem__("dE", gr.detach().float().cpu()))
        return e
    model._encode_items = enc
    orig_ue = model.user_encoder.forward
    def ue(x, lm, lr, orig_ue=orig_ue, cap=cap):
        x.register_hook(lambda gr: cap.__setitem__("dX", gr.detach().float().cpu()))
        o = orig_ue(x, lm, lr)
        cap["P"] = o.detach().float().cpu()
        o.register_hook(lambda gr: cap.__setitem__("dP", gr.detach().float().cpu()))
        return o
    model.user_encoder.forward = ue
    model.zero_grad()
    loss = model(g["ids"].reshape(-1).cuda(), g["items"].cuda(), g["log_mask"].cuda(), 0)
    loss.backward()
    cap["loss"] = float(loss)
    cap["g_fc"] = model.bert_encoder.text_encoders.title.fc.weight.grad.float().cpu()
    res[mode] = cap
a, b = res["fp32"], res["bf16"]
print("loss", a["loss"], b["loss"])
for k in ("E", "P", "dP", "dE", "dX", "g_fc"):
    x, y = a[k].double(), b[k].double()
    print(f"{k:5s} rel err {float((x-y).norm()/(x.norm()+1e-30)):.4f}  |ref| {float(x.norm()):.4e} shape {tuple(x.shape)}")
