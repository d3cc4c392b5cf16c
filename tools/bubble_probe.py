"""Where the GPU idles at the start of a step: host time from the plan's device->host sync (GPU idle from here) to the
launch of the first BERT layer, and to the end of issue; plus CUDA-event time of whole steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
from idvs.morec_b200 import lib

step, host, resident, _ = bench.setup_training(dict(bench.CFG), "bf16", 12)
big = max(range(len(host)), key=lambda i: int((host[i][1].reshape(-1, 60)[:, 30:] != 0).sum()))
step(*resident[big])
marks = {}
orig_d2h, orig_layer, orig_embed = lib.d2h_end, lib.bert_layer_fwd, lib.bert_embed_fwd
def d2h(h):
    t0 = time.perf_counter(); r = orig_d2h(h); t1 = time.perf_counter()
    marks.setdefault("sync_wait", []).append(t1 - t0); marks["t_sync"] = t1
    marks["first"] = True
    return r
def embed(*a, **k):
    if marks.get("first"):
        marks.setdefault("sync_to_embed", []).append(time.perf_counter() - marks["t_sync"])
    return orig_embed(*a, **k)
lib.bert_embed_fwd = embed
def layer(a):
    if marks.get("first"):
        marks.setdefault("sync_to_layer0", []).append(time.perf_counter() - marks["t_sync"]); marks["first"] = False
    return orig_layer(a)
lib.d2h_end, lib.bert_layer_fwd = d2h, layer
import idvs.morec_b200.model.model as mm, idvs.morec_b200.ops as ops
for i in range(4):
    step(*resident[i])
torch.cuda.synchronize()
marks.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
for i in range(4, 12):
    step(*resident[i])
t1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
print(f"gpu {e0.elapsed_time(e1)/8:.2f} ms/step, host issue {1e3*(t1-t0)/8:.2f} ms/step")
for k in ("sync_wait", "sync_to_embed", "sync_to_layer0"):
    v = marks[k][1:]
    print(k, " ".join(f"{1e3*x:.2f}" for x in v), "ms")
