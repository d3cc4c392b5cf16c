"""Pin the CPU oracle against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py) and against the installed HF BertModel."""
import numpy as np
import pytest
import torch

from oracle import morec_oracle as O

CASES = ["id_small_collide", "id_cfg1_shape", "text_tiny", "text_tiny_collide"]


def run_oracle(g, dtype=torch.float32, grad=False):
    m = g["meta"]
    p = {k: v.to(dtype).clone().requires_grad_(grad) if v.is_floating_point() else v for k, v in g["state_dict"].items()}
    out = O.model_forward(p, g["ids"], g["items"], g["log_mask"], g["pop_prob"], use_modal=m["modal"],
                          n_heads_user=m["heads"], n_heads_bert=(m["bert_cfg"] or {}).get("num_attention_heads", 0))
    return p, out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_forward(goldens, name):
    g = goldens[name]
    _, out = run_oracle(g)
    assert abs(float(out.loss) - float(g["loss"])) <= 1e-5
    nonpad = (g["ids"].reshape(-1) != 0)
    assert torch.allclose(out.score_embs[nonpad], g["score_embs"][nonpad], atol=2e-5, rtol=1e-5)
    pv = g["prec_vec"].reshape(out.prec_vec.shape)
    valid = O.valid_rows(g["log_mask"])
    assert torch.allclose(out.prec_vec[valid], pv[valid], atol=5e-5, rtol=1e-4)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_grads(goldens, name):
    g = goldens[name]
    p, out = run_oracle(g, grad=True)
    out.loss.backward()
    for k, gref in g["grads"].items():
        if k.startswith("bert_encoder") and "pooler" in k:
            continue
        got = p[k].grad
        assert got is not None, k
        scale = float(gref.abs().max()) + 1e-12
        if k == "id_embedding.weight":
            # reference nn.Embedding(padding_idx=0) zeroes the grad of row 0; pad slots get exactly 0 anyway
            assert float(got[0].abs().max()) == 0.0
        assert float((got - gref).abs().max()) <= 2e-4 * scale + 1e-7, k


@pytest.mark.parametrize("name", CASES)
def test_mask_closed_form_equals_loops(goldens, name):
    g = goldens[name]
    if g["ids"].numel() > 400:
        ids = g["ids"][:5, -7:].clone()
    else:
        ids = g["ids"]
    assert torch.equal(O.reject_mask_closed_form(ids), O.reject_mask_loops(ids))


def test_mask_edge_cases():
    # fully padded user except the mandatory last item, duplicates inside a user, same item across users, B=1
    ids = torch.tensor([[0, 0, 0, 5], [5, 5, 7, 5], [0, 7, 5, 9]])
    a, b = O.reject_mask_closed_form(ids), O.reject_mask_loops(ids)
    assert torch.equal(a, b)
    one = torch.tensor([[0, 3, 3, 4]])
    assert torch.equal(O.reject_mask_closed_form(one), O.reject_mask_loops(one))
    # target column of a valid row is never masked
    B, Lp1 = ids.shape
    tgt = O.ce_labels(B, Lp1 - 1)
    v = O.valid_rows(O.log_mask_from_ids(ids))
    assert not a[torch.arange(a.shape[0]), tgt][v].any()


def test_labels_and_logmask():
    assert O.ce_labels(2, 3).tolist() == [1, 2, 3, 5, 6, 7]
    ids = torch.tensor([[0, 0, 4, 9], [1, 2, 3, 4]])
    assert O.log_mask_from_ids(ids).tolist() == [[0, 0, 1], [1, 1, 1]]


def test_bert_restatement_matches_hf():
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    cfg = BertConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                     vocab_size=100, max_position_embeddings=32)
    m = BertModel(cfg).eval()
    ids = torch.randint(1, 100, (5, 9))
    am = torch.ones(5, 9, dtype=torch.long)
    am[1, 5:] = 0
    am[3, 2:] = 0
    ids = ids * am
    with torch.no_grad():
        ref = m(input_ids=ids, attention_mask=am)[0]
        got = O.bert_forward(dict(m.state_dict()), ids, am, 4)
    # only compare tokens that can matter (valid tokens); pad-token rows are never consumed
    assert torch.allclose(got[am.bool()], ref[am.bool()], atol=2e-5, rtol=1e-5)


def test_synth_batch_properties():
    d = O.synth_batch(16, 25, 5000, 30, 3, modal=True)
    ids = d["ids"]
    assert ids.shape == (16, 26) and (ids[:, -1] != 0).all()
    nz = (ids != 0).sum(1)
    assert int(nz.min()) >= 3
    # left padded: zeros then non-zeros
    for r in ids:
        k = int((r != 0).nonzero()[0])
        assert (r[:k] == 0).all() and (r[k:] != 0).all()
    assert (d["pop_prob"][ids.reshape(-1)] > 0).all() and float(d["pop_prob"][0]) == 1.0
    assert abs(float(d["pop_prob"][1:].sum()) - 1.0) < 1e-9
    it = d["items"]
    assert it.shape == (16 * 26, 60)
    assert (it[ids.reshape(-1) == 0] == 0).all()
    assert (it[ids.reshape(-1) != 0, 0] == 101).all()


def _vision_oracle(g, grad=False):
    from transformers import SwinConfig, SwinForImageClassification
    m = g["meta"]
    net = SwinForImageClassification(SwinConfig(**m["swin_cfg"]))
    net.classifier = torch.nn.Linear(net.classifier.in_features, m["D"])
    sd = g["state_dict"]
    net.load_state_dict({k[len("cv_encoder.image_net."):]: v for k, v in sd.items() if k.startswith("cv_encoder.image_net.")})
    net.eval()
    p = {k: v.clone().requires_grad_(grad) for k, v in sd.items() if k.startswith("user_encoder.")}
    if not grad:
        with torch.no_grad():
            E = O.vision_item_encoder(net, g["images"])
    else:
        E = O.vision_item_encoder(net, g["images"])
    return net, p, O.model_forward_from_embs(p, E, g["ids"], g["log_mask"], g["pop_prob"], m["heads"])


def test_vision_oracle_matches_reference():
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vision_tiny.pt"), map_location="cpu", weights_only=False)
    net, p, out = _vision_oracle(g, grad=True)
    assert abs(float(out.loss) - float(g["loss"])) <= 1e-5
    assert torch.allclose(out.score_embs, g["score_embs"], atol=2e-5)
    out.loss.backward()
    for k, gref in g["grads"].items():
        if k.startswith("cv_encoder.image_net."):
            got = dict(net.named_parameters())[k[len("cv_encoder.image_net."):]].grad
        else:
            got = p[k].grad
        assert got is not None, k
        assert float((got - gref).abs().max()) <= 2e-4 * (float(gref.abs().max()) + 1e-12) + 1e-7, k


def test_mask_closed_form_equals_loops_random():
    """property test: the closed form of SURVEY.md Appendix A == the element-by-element re-derivation of
    model.py:51-63 on random left-padded batches with heavy id collisions (small catalogues), B = 1 included"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 5), st.integers(2, 7), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
    def check(B, L, N, seed):
        g = torch.Generator().manual_seed(seed)
        ids = torch.randint(1, N + 1, (B, L + 1), generator=g)
        n_real = torch.randint(2, L + 2, (B,), generator=g)              # at least the target and one input
        for b in range(B):
            ids[b, : L + 1 - int(n_real[b])] = 0                         # left padding (dataset.py:24-36)
        a, lo = O.reject_mask_closed_form(ids), O.reject_mask_loops(ids)
        assert torch.equal(a, lo)
        tgt = O.ce_labels(B, L)
        v = O.valid_rows(O.log_mask_from_ids(ids))
        assert not a[torch.arange(a.shape[0]), tgt][v].any()             # a valid row never masks its own target

    check()


@pytest.mark.parametrize("name", ["bert_base_b4", "bert_tiny_t128_b16", "swin_t_b2", "swin_b_b2"])
def test_oracle_matches_reference_real_configs(name):
    """REAL encoder configurations (BERT-base 12 layers; BERT-tiny T=128 at B=16; Swin-T / Swin-B at 224x224): the
    seeded construction reproduces the reference's weights exactly, and the oracle reproduces the unmodified
    reference's loss, item embeddings and every parameter gradient (tests/golden/make_golden_real.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import real_cases as RC
    c = RC.CASES[name]
    g = RC.load_golden(name)
    d = RC.build_inputs(c)
    for k, v in g["input_checksums"].items():
        assert abs(float(d[k].double().abs().sum()) - v) <= 1e-9 * abs(v), k
    if c["kind"] == "text":
        from idvs.morec_b200.model import Model
    else:
        from idvs.morec_b200.model_vision import Model
    model = RC.build_model(c, Model, d["pop_prob"])
    bad = RC.checksums_match(RC.checksums(model.state_dict()), g["weight_checksums"])
    assert not bad, f"portable weights diverged from the reference's at {bad[:5]}"
    out, grads = RC.run_oracle(c, model, d)
    nonpad = d["ids"].reshape(-1) != 0
    if c["kind"] == "vision":
        nonpad = torch.ones_like(nonpad)            # the reference encodes the zero image of a pad slot too
    bad = RC.compare_to_golden(g, out.loss, out.score_embs.detach(), grads, nonpad, loss_tol=2e-5, emb_tol=5e-5, grad_tol=3e-3)    # fp32 op-order noise through 12 / 24 layers: measured <= 2.9e-3
    assert not bad, bad[:10]


def test_bce_oracle_matches_reference():
    """BCE head (bce_text/main-end2end/model/model.py:30-51): the oracle reproduces the unmodified reference's loss
    and every parameter gradient on a seeded BERT-tiny case (tests/golden/make_golden_bce.py)"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_bce as GB
    import real_cases as RC
    from idvs.morec_b200.model_bce import Model
    c = GB.CASE
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "real_bce_tiny.pt"), map_location="cpu", weights_only=False)
    d = GB.bce_inputs(c)
    model = GB.build(c, Model)
    assert not RC.checksums_match(RC.checksums(model.state_dict()), g["weight_checksums"])
    p = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    loss = O.bce_model_forward(p, d["items"], d["log_mask"], use_modal=True, max_seq_len=c["L"], n_heads_user=c["heads"],
                               n_heads_bert=c["bert_heads"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 2e-5
    for k, ref in g["grads"].items():
        if "pooler" in k or RC.is_null_gradient(k):
            continue
        f = p[k].grad.double().reshape(-1)
        smp = f[RC.grad_sample_index(f.numel())].float()
        assert float((smp - ref["sample"]).abs().max()) <= 5e-4 * ref["absmax"] + 1e-7, k
