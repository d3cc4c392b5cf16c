"""Per-tensor deviation of the fast modes from the reference goldens (GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_parity_gpu as T

for name in ("text_tiny",):
    g = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), map_location="cpu", weights_only=False)
    for mode in ("bf16",):
        model = T.build_model(g)
        model.set_compute_dtype(mode)
        loss, E, P, grads = T.run_cuda(model, g)
        print(f"== {name} {mode}: loss {loss:.6f} ref {float(g['loss']):.6f}")
        nonpad = g["ids"].reshape(-1) != 0
        print("   score_embs max err", float((E[nonpad] - g["score_embs"][nonpad]).abs().max()))
        worst = []
        for k, gref in g["grads"].items():
            if "pooler" in k: continue
            a, b = grads[k].double(), gref.double()
            rel = float((a - b).norm() / (b.norm() + 1e-30))
            worst.append((rel, k, float(b.norm())))
        worst.sort(reverse=True)
        for r, k, n in worst[:45]:
            print(f"   {r:.4f}  |g|={n:.3e}  {k}")
