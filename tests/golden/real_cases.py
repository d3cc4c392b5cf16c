"""Seeded construction of the REAL-configuration parity cases (BASELINE.json configs 2-5 at sizes a CPU finishes in
seconds).  Shared by tests/golden/make_golden_real.py (which runs the UNMODIFIED reference on them in the authoring
container) and by the tests (which rebuild the very same inputs and weights from the seed on any box).

Weights are NOT stored in the fixture (BERT-base is 440 MB).  torch's CPU initialisers are not bit-portable across
hosts (vectorised normal_ paths differ with the CPU), so after construction EVERY floating-point tensor of the model
is overwritten from numpy's PCG64 generator (`portable_reinit`: bit-identical on every platform) with
initialiser-like statistics -- and non-trivial LayerNorm gains / biases, which exercises more of the backward than
the stock ones / zeros.  The fixture keeps a per-tensor checksum of the reference's state dict so a silent divergence
fails loudly.
Encoder configurations are the reference's own `pretrained_models/<name>/config.json` (values restated here because
/root/reference does not exist on the GPU box).
"""
import types

import numpy as np
import torch

# pretrained_models/bert_base_uncased/config.json
BERT_BASE = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, layer_norm_eps=1e-12, initializer_range=0.02, pad_token_id=0)
# BERT-tiny as inbatch_sasrec_e2e_text/run.py:55-57 expects it (H=128, 2 layers, 2 heads, I=512)
BERT_TINY = dict(vocab_size=30522, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, layer_norm_eps=1e-12, initializer_range=0.02, pad_token_id=0)
# pretrained_models/swin_tiny/config.json and swin_base/config.json
SWIN_TINY = dict(image_size=224, patch_size=4, num_channels=3, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                 window_size=7, mlp_ratio=4.0, qkv_bias=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                 drop_path_rate=0.1, hidden_act="gelu", layer_norm_eps=1e-5, initializer_range=0.02,
                 use_absolute_embeddings=False)
SWIN_BASE = dict(SWIN_TINY, embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32])

CASES = {
    # cfg-3 architecture (headline): BERT-base 12 layers, T=30, L=25, D=512
    "bert_base_b4": dict(kind="text", seed=41, B=4, L=25, N=500, T=30, D=512, heads=2, blocks=2, bert=BERT_BASE,
                         bert_name="bert_base_uncased", word_dim=768, bert_heads=12),
    # cfg-2: BERT-tiny with T = 128 word pieces, B = 16
    "bert_tiny_t128_b16": dict(kind="text", seed=42, B=16, L=25, N=2000, T=128, D=64, heads=2, blocks=2, bert=BERT_TINY,
                               bert_name="bert_tiny", word_dim=128, bert_heads=2),
    # cfg-4 architecture: the real Swin-T (4 stages, 224x224, last stage 7x7 = one window, no shift)
    "swin_t_b2": dict(kind="vision", seed=43, B=2, L=3, N=40, D=512, heads=2, blocks=2, swin=SWIN_TINY, cv_name="swin_tiny"),
    # cfg-5 architecture: Swin-B
    "swin_b_b2": dict(kind="vision", seed=44, B=2, L=2, N=40, D=512, heads=2, blocks=2, swin=SWIN_BASE, cv_name="swin_base"),
}


def make_args(c):
    a = types.SimpleNamespace(max_seq_len=c["L"], embedding_dim=c["D"], num_attention_heads=c["heads"], drop_rate=0.1,
                              transformer_block=c["blocks"])
    if c["kind"] == "text":
        a.num_words_title = c["T"]; a.num_words_abstract = 50; a.num_words_body = 50; a.news_attributes = ["title"]
        a.bert_model_load = c["bert_name"]; a.word_embedding_dim = c["word_dim"]
    else:
        a.CV_model_load = c["cv_name"]
    return a


def build_inputs(c):
    """ids [B, L+1], items ([C, 2T] tokens | [C, 3, 224, 224] images), log_mask [B, L], pop_prob [N+1]"""
    from idvs.morec_b200.synth import synth_batch
    if c["kind"] == "text":
        d = synth_batch(c["B"], c["L"], c["N"], c["T"], c["seed"], modal=True, n_users_pop=300)
        return dict(ids=d["ids"], items=d["items"], log_mask=d["log_mask"], pop_prob=d["pop_prob"])
    d = synth_batch(c["B"], c["L"], c["N"], 0, c["seed"], modal=False, n_users_pop=50, mind_shape=False)   # HM-shape
    ids = d["ids"]
    if c["B"] > 1:
        ids[0, 0] = 0                                   # one pad slot (zero image) and one duplicate item in the batch
        ids[1, 1] = ids[0, 2]
    g = np.random.default_rng(c["seed"] + 7)                      # numpy: bit-portable across hosts
    content = torch.from_numpy(g.standard_normal((c["N"] + 1, 3, 224, 224), dtype=np.float32))
    content[0] = 0
    from idvs.morec_b200.synth import log_mask_from_ids
    return dict(ids=ids, items=content[ids.reshape(-1)], log_mask=log_mask_from_ids(ids), pop_prob=d["pop_prob"])


def build_encoder(c):
    """the HF encoder exactly as the reference's run.py prepares it (T/run.py:51-75, V/run.py:47-60), seeded"""
    torch.manual_seed(c["seed"])
    np.random.seed(c["seed"])
    if c["kind"] == "text":
        from transformers import BertConfig, BertModel
        net = BertModel(BertConfig(**c["bert"]))
        pooler = {"bert_base_uncased": (197, 198), "bert_tiny": (37, 38)}[c["bert_name"]]
        for i, (n, p) in enumerate(net.named_parameters()):
            if i in pooler:
                p.requires_grad = False
        return net
    from transformers import SwinConfig, SwinForImageClassification
    net = SwinForImageClassification(SwinConfig(**c["swin"]))
    net.classifier = torch.nn.Linear(net.classifier.in_features, c["D"])
    torch.nn.init.xavier_normal_(net.classifier.weight.data)
    torch.nn.init.constant_(net.classifier.bias.data, 0)
    return net


def portable_reinit(model, seed):
    """overwrite every floating-point parameter / buffer from numpy's PCG64 stream, in sorted key order:
    matrices ~ N(0, min(0.05, sqrt(2 / (rows + cols)))), LayerNorm gains ~ 1 + 0.1 N(0, 1), other vectors ~ 0.02 N(0, 1)"""
    rng = np.random.default_rng(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for name in sorted(sd):
            t = sd[name]
            if not t.is_floating_point():
                continue
            z = rng.standard_normal(t.numel(), dtype=np.float32).reshape(tuple(t.shape))
            low = name.lower()
            if t.dim() >= 2:
                std = np.float32(min(0.05, (2.0 / (t.shape[0] + t.shape[-1])) ** 0.5))
                v = z * std
            elif ("layernorm" in low or "layer_norm" in low) and low.endswith("weight"):
                v = np.float32(1.0) + np.float32(0.1) * z
            else:
                v = np.float32(0.02) * z
            t.copy_(torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)))
    return model


def build_model(c, ModelCls, pop_prob):
    """ModelCls(args, N, True, encoder, pop_prob), then portable_reinit -> bit-identical weights for the reference's
    class and for idvs.morec_b200's on any host (both expose the same state-dict keys)"""
    net = build_encoder(c)
    torch.manual_seed(c["seed"] + 1)
    model = ModelCls(make_args(c), c["N"], True, net, pop_prob.numpy()).eval()
    return portable_reinit(model, c["seed"] + 2)


def checksums(state_dict):
    return {k: float(v.double().abs().sum()) for k, v in state_dict.items() if v.is_floating_point()}


def checksums_match(cs, ref, rtol=1e-9):
    """(summation order of the checksum itself may differ between hosts: compare to 1e-9 relative)"""
    bad = [k for k, v in ref.items() if k not in cs or abs(cs[k] - v) > rtol * max(abs(v), 1e-30)]
    return bad


def grad_sample_index(numel, n=256):
    n = min(n, numel)
    return (torch.arange(n, dtype=torch.int64) * (numel - 1)) // max(n - 1, 1)


def summarize_grads(grads):
    """per-tensor L2 norm, max-abs and a strided sample of <= 256 elements (the fixture stays small)"""
    out = {}
    for k, g in grads.items():
        f = g.detach().double().reshape(-1)
        out[k] = dict(norm=float(f.norm()), absmax=float(f.abs().max()), sample=f[grad_sample_index(f.numel())].float().clone())
    return out


def load_golden(name):
    import os
    return torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"real_{name}.pt"), map_location="cpu",
                      weights_only=False)


def run_oracle(c, model, d):
    """fwd + bwd of the CPU oracle (oracle/morec_oracle.py) on the state dict of `model` (an idvs.morec_b200 Model built
    by build_model on the CPU).  Returns (StepOut, {parameter name: gradient}).  Vision: the oracle's item tower is
    the installed HF SwinForImageClassification itself (oracle.vision_item_encoder), run on a deep copy."""
    import copy
    from oracle import morec_oracle as O
    sd = model.state_dict()
    p = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    if c["kind"] == "text":
        out = O.model_forward(p, d["ids"], d["items"], d["log_mask"], d["pop_prob"], use_modal=True,
                              n_heads_user=c["heads"], n_heads_bert=c["bert_heads"])
        out.loss.backward()
        grads = {k: v.grad for k, v in p.items() if v.is_floating_point() and v.grad is not None}
        return out, grads
    net = copy.deepcopy(model.cv_encoder.image_net).eval()
    E = O.vision_item_encoder(net, d["items"])
    out = O.model_forward_from_embs(p, E, d["ids"], d["log_mask"], d["pop_prob"], c["heads"])
    out.loss.backward()
    grads = {k: v.grad for k, v in p.items() if v.is_floating_point() and v.grad is not None}
    for k, v in net.named_parameters():
        if v.grad is not None:
            grads["cv_encoder.image_net." + k] = v.grad
    return out, grads


def is_null_gradient(name):
    """The key-projection bias of softmax attention has an EXACTLY zero gradient: q.(k + b) = q.k + q.b adds the same
    constant to every key of a query, which softmax ignores.  The reference's value for it is rounding noise (~1e-9 of
    its neighbours), so relative error is meaningless; it is only required to stay negligible."""
    return name.endswith("attention.self.key.bias")


def compare_to_golden(g, loss, score_embs, grads, nonpad, *, loss_tol, emb_tol, grad_tol):
    """loss / item embeddings / every parameter gradient (strided sample + L2 norm) against the reference fixture;
    returns the list of violations (empty = pass).  grad_tol is relative to the reference tensor's max-abs."""
    bad = []
    if abs(float(loss) - float(g["loss"])) > loss_tol:
        bad.append(("loss", float(loss), float(g["loss"])))
    e = float((score_embs[nonpad].float() - g["score_embs"][nonpad]).abs().max())
    if e > emb_tol:
        bad.append(("score_embs", e, emb_tol))
    gmax = max(r["absmax"] for r in g["grads"].values())
    nmax = max(r["norm"] for r in g["grads"].values())
    # floors: a tensor whose whole gradient is ~1e-5 of the largest one (e.g. q / k of Swin's last single-window
    # block) is rounding noise of its neighbours; it only has to stay that small
    for k, ref in g["grads"].items():
        if "pooler" in k:
            continue
        if k not in grads:
            bad.append((k, "missing gradient", None))
            continue
        f = grads[k].detach().double().reshape(-1).cpu()
        if is_null_gradient(k):
            lim = max(1e-2 * g["grads"][k.replace("key.bias", "query.bias")]["absmax"], 1e-5 * gmax)
            if float(f.abs().max()) > lim:
                bad.append((k, "null gradient too large", float(f.abs().max()), lim))
            continue
        smp = f[grad_sample_index(f.numel())].float()
        scale = ref["absmax"] + 1e-12
        err = float((smp - ref["sample"]).abs().max())
        if err > grad_tol * scale + 1e-5 * gmax + 1e-7:
            bad.append((k, "sample", err / scale))
        if abs(float(f.norm()) - ref["norm"]) > grad_tol * ref["norm"] + 1e-4 * nmax + 1e-7:
            bad.append((k, "norm", float(f.norm()), ref["norm"]))
    return bad
