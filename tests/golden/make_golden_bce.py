"""Golden fixture for the BCE head: the UNMODIFIED reference bce_text/main-end2end/model on CPU fp32 (authoring
container only), BERT-tiny item tower, seeded weights (rebuilt by the tests), sampled negatives stored in the fixture."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import real_cases as RC  # noqa: E402

CASE = dict(kind="text", seed=51, B=6, L=9, N=200, T=14, D=64, heads=2, blocks=2, bert=RC.BERT_TINY, bert_name="bert_tiny",
            word_dim=128, bert_heads=2)


def bce_inputs(c):
    """ids [B, L+1, 2] (positive, sampled negative not in the user's sequence; 0 on pad slots and on the last slot),
    token rows [B*(L+1)*2, 2T], log_mask [B, L]"""
    from idvs.morec_b200.synth import synth_batch
    d = synth_batch(c["B"], c["L"], c["N"], c["T"], c["seed"], modal=True, n_users_pop=100)
    g = np.random.default_rng(c["seed"] + 3)
    ids = d["ids"].numpy()
    neg = np.zeros_like(ids)
    for b in range(ids.shape[0]):
        seq = set(ids[b].tolist())
        for t in range(ids.shape[1] - 1):
            if ids[b, t] != 0:
                x = int(g.integers(1, c["N"] + 1))
                while x in seq:
                    x = int(g.integers(1, c["N"] + 1))
                neg[b, t] = x
    pair = torch.from_numpy(np.stack([ids, neg], axis=-1))                      # [B, L+1, 2]
    items = d["item_content"][pair.reshape(-1)]
    return dict(pair=pair, items=items, log_mask=d["log_mask"])


def build(c, ModelCls):
    net = RC.build_encoder(c)
    torch.manual_seed(c["seed"] + 1)
    return RC.portable_reinit(ModelCls(RC.make_args(c), c["N"], True, net).eval(), c["seed"] + 2)


if __name__ == "__main__":
    d0 = "/root/reference/bce_text/main-end2end/model"
    spec = importlib.util.spec_from_file_location("_ref_bce_model", os.path.join(d0, "__init__.py"), submodule_search_locations=[d0])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_bce_model"] = mod
    spec.loader.exec_module(mod)
    c = CASE
    d = bce_inputs(c)
    model = build(c, mod.Model)
    loss = model(d["items"], d["log_mask"], "cpu")
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    out = dict(meta=dict(case=c, torch=torch.__version__, reference_commit="ce372cf"), loss=loss.detach().clone(),
               grads=RC.summarize_grads(grads), weight_checksums=RC.checksums(model.state_dict()))
    path = os.path.join(HERE, "real_bce_tiny.pt")
    torch.save(out, path)
    print(f"bce_tiny: loss={float(loss):.6f} n_grads={len(grads)} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")
