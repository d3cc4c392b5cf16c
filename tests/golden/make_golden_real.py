"""Golden fixtures for the REAL encoder configurations (BERT-base 12 layers, BERT-tiny with T=128 at B=16, Swin-T and
Swin-B at 224x224), produced by the UNMODIFIED reference `model` packages on CPU fp32.

Run ONLY in the authoring container (needs /root/reference):

    python tests/golden/make_golden_real.py [case ...]

Weights are rebuilt from the seed by the tests (tests/golden/real_cases.py), so the fixture holds the inputs' seed,
the reference's loss, score_embs, prec_vec, a per-tensor summary of every parameter gradient (L2 norm, max-abs and a
strided 256-element sample) and per-tensor weight checksums -- a few hundred KB per case instead of 0.5 GB.
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import real_cases as RC  # noqa: E402

REF = "/root/reference"


def ref_model_cls(pkg):
    name = "_ref_" + pkg + "_model"
    if name in sys.modules:
        return sys.modules[name].Model
    d = os.path.join(REF, pkg, "model")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod.Model


def run_case(name):
    c = RC.CASES[name]
    Model = ref_model_cls("inbatch_sasrec_e2e_text" if c["kind"] == "text" else "inbatch_sasrec_e2e_vision")
    d = RC.build_inputs(c)
    model = RC.build_model(c, Model, d["pop_prob"])
    cap = {}
    tower = model.bert_encoder if c["kind"] == "text" else model.cv_encoder
    h1 = tower.register_forward_hook(lambda m, i, o: cap.__setitem__("score_embs", o.detach().clone()))
    h2 = model.user_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("prec_vec", o.detach().clone()))
    loss = model(d["ids"].reshape(-1), d["items"], d["log_mask"], "cpu")
    loss.backward()
    h1.remove(); h2.remove()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    out = dict(meta=dict(name=name, case={k: v for k, v in c.items()}, torch=torch.__version__,
                         transformers=__import__("transformers").__version__, reference_commit="ce372cf"),
               loss=loss.detach().clone(), score_embs=cap["score_embs"], prec_vec=cap["prec_vec"],
               grads=RC.summarize_grads(grads), weight_checksums=RC.checksums(model.state_dict()),
               input_checksums={k: float(v.double().abs().sum()) for k, v in d.items()})
    path = os.path.join(HERE, f"real_{name}.pt")
    torch.save(out, path)
    print(f"{name}: loss={float(loss):.6f} n_grads={len(grads)} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    assert os.path.isdir(REF), "needs /root/reference (authoring container)"
    for n in (sys.argv[1:] or list(RC.CASES)):
        run_case(n)
