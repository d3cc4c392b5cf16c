"""CPU oracle for the MoRec in-batch training step  --  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 (optionally fp64) PyTorch-on-CPU restatement of the reference's hot path.
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product (``idvs.morec_b200``) never
does; it fails loudly when the CUDA extension is missing.

Parity pin: the reference ships no tests, golden vectors or fixtures of its own, so
the pin is *generated*: ``tests/golden/make_golden.py`` imports the UNMODIFIED reference ``model``
package from /root/reference in the authoring container, runs it on CPU fp32 with fixed seeds and
commits inputs, weights and outputs as small fixtures under ``tests/golden/``.  ``tests/test_oracle.py``
checks every function below against those fixtures (bit-exact for masks/labels/valid rows, <=1e-5 for
loss/logits/embeddings, grads <=1e-5 relative).

What is restated, and where it lives in the reference (paths relative to /root/reference):

* input layout / left padding / log_mask ........ inbatch_sasrec_e2e_text/data_utils/dataset.py:24-36
* popularity table p_i and log p[ids] ............ data_utils/preprocess.py:50,60-61,71-76; model/model.py:14,32-33
* item tower dispatch (ID embedding | text CLS) .. model/model.py:34-37; model/encoders.py:63-70,107-117
* BERT encoder (third-party HF ``transformers`` BertModel, reference pins ==4.20.1, README.md:44; the
  installed 5.5.0 is the de-facto oracle).  Restated here from the published algorithm (post-LN BERT:
  word+pos+type embeddings -> LN(1e-12); per layer QKV(+bias) -> softmax(QK^T/sqrt(d_h) + key mask) -> V
  -> dense + residual + LN -> dense + erf-GELU -> dense + residual + LN) and pinned against the
  installed ``BertModel`` itself in tests/test_oracle.py.  Call site: model/encoders.py:68.
* SASRec user encoder ............................ model/encoders.py:23-28 (mask), model/modules.py:14-17,
  27-31, 52-63, 73-75, 89-96
* labels, scoring, debias, masks, CE ............. model/model.py:45-67

Nothing here is copied from the reference: the mask is the closed form of SURVEY.md Appendix A, the
loops version (``reject_mask_loops``) re-derives the same predicate element by element for small cases.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

NEG_MASK = -1e4      # masked logit value (finite, NOT -inf): model/model.py:52,63
ATT_NEG = -1e9       # SASRec additive attention mask: model/encoders.py:27


# --------------------------------------------------------------------------------------------------
# integer / predicate part (bit-exact)
# --------------------------------------------------------------------------------------------------

def ce_labels(B: int, L: int) -> torch.Tensor:
    """target(r) for r = b*L + (j-1), j = 1..L  ==  b*(L+1) + j   (model/model.py:45-48)."""
    b = torch.arange(B).view(B, 1)
    j = torch.arange(1, L + 1).view(1, L)
    return (b * (L + 1) + j).reshape(-1).to(torch.long)


def log_mask_from_ids(ids: torch.Tensor) -> torch.Tensor:
    """dataset.py:24-36: log_mask[b,t] = 1 <=> slot t (t < L) of user b holds a real item."""
    return (ids[:, :-1] != 0).to(torch.float32)


def reject_mask_closed_form(ids: torch.Tensor) -> torch.Tensor:
    """masked(r,c) = (id_c == 0) or (id_c in ids(user(r)) and c != target(r)).

    ids: int64 [B, L+1] -> bool [B*L, B*(L+1)].  Closed form of model/model.py:51-63.
    """
    B, Lp1 = ids.shape
    L = Lp1 - 1
    flat = ids.reshape(-1)                                   # [C]
    C = flat.numel()
    # member[b, c] = id_c in {ids[b, :]}
    member = (flat.view(1, 1, C) == ids.view(B, Lp1, 1)).any(dim=1)      # [B, C]
    member = member.view(B, 1, C).expand(B, L, C).reshape(B * L, C).clone()
    tgt = ce_labels(B, L)
    member[torch.arange(B * L), tgt] = False
    pad_col = (flat == 0).view(1, C)
    return member | pad_col


def reject_mask_loops(ids: torch.Tensor) -> torch.Tensor:
    """Element-by-element derivation of the same predicate (small cases only).

    Follows the *order of operations* of model/model.py:51-63: pad columns first, then for every user
    the membership mask with the row's own target re-enabled.  Note the subtlety the closed form must
    reproduce: a pad column stays masked even when it is... never a target (targets of valid rows are
    real items), and un-masking the target only undoes the membership mask, not the pad mask -- but a
    target column of a *valid* row is never a pad column, and invalid rows are dropped.
    """
    B, Lp1 = ids.shape
    L = Lp1 - 1
    C = B * Lp1
    flat = [int(v) for v in ids.reshape(-1)]
    out = torch.zeros(B * L, C, dtype=torch.bool)
    for b in range(B):
        own = set(int(v) for v in ids[b])
        for j in range(L):
            r = b * L + j
            tgt = b * Lp1 + j + 1
            for c in range(C):
                m = flat[c] == 0
                if flat[c] in own and c != tgt:
                    m = True
                out[r, c] = m
    return out


def valid_rows(log_mask: torch.Tensor) -> torch.Tensor:
    """rows kept by the CE: model/model.py:65."""
    return (log_mask.reshape(-1) != 0)


# --------------------------------------------------------------------------------------------------
# SASRec user encoder (model/encoders.py:23-28, model/modules.py)
# --------------------------------------------------------------------------------------------------

def sasrec_att_mask(log_mask: torch.Tensor) -> torch.Tensor:
    """[B,1,L,L] additive mask: 0 where (k <= q and log_mask[b,k] != 0) else -1e9."""
    B, L = log_mask.shape
    key_valid = (log_mask != 0).view(B, 1, 1, L).expand(B, 1, L, L)
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool)).view(1, 1, L, L)
    ok = key_valid & causal
    return torch.where(ok, torch.zeros((), dtype=log_mask.dtype), torch.full((), ATT_NEG, dtype=log_mask.dtype))


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def sasrec_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, log_mask: torch.Tensor,
                   n_heads: int, prefix: str = "user_encoder.transformer_encoder.") -> torch.Tensor:
    """x [B,L,D] -> [B,L,D]; eval-mode (no dropout).  Weight names = reference state_dict keys."""
    B, L, D = x.shape
    dk = D // n_heads
    mask = sasrec_att_mask(log_mask.to(x.dtype))
    pos = p[prefix + "position_embedding.weight"][:L]
    h = layer_norm(x + pos.view(1, L, D), p[prefix + "layer_norm.weight"], p[prefix + "layer_norm.bias"], 1e-6)
    blk = 0
    while (prefix + f"transformer_blocks.{blk}.multi_head_attention.w_Q.weight") in p:
        q_ = prefix + f"transformer_blocks.{blk}."
        a = q_ + "multi_head_attention."
        q = (h @ p[a + "w_Q.weight"].t()).view(B, L, n_heads, dk).transpose(1, 2)
        k = (h @ p[a + "w_K.weight"].t()).view(B, L, n_heads, dk).transpose(1, 2)
        v = (h @ p[a + "w_V.weight"].t()).view(B, L, n_heads, dk).transpose(1, 2)
        s = q @ k.transpose(-2, -1) / (dk ** 0.5) + mask
        pr = torch.softmax(s, dim=-1)
        o = (pr @ v).transpose(1, 2).reshape(B, L, D)
        o = o @ p[a + "fc.weight"].t()
        h1 = layer_norm(h + o, p[a + "layer_norm.weight"], p[a + "layer_norm.bias"], 1e-6)
        f = q_ + "feed_forward."
        u = torch.relu(h1 @ p[f + "w_1.weight"].t() + p[f + "w_1.bias"])
        u = u @ p[f + "w_2.weight"].t() + p[f + "w_2.bias"]
        h = layer_norm(h1 + u, p[f + "layer_norm.weight"], p[f + "layer_norm.bias"], 1e-6)
        blk += 1
    return h


# --------------------------------------------------------------------------------------------------
# BERT text tower (HF BertModel restated; model/encoders.py:63-70)
# --------------------------------------------------------------------------------------------------

def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def bert_forward(p: Dict[str, torch.Tensor], input_ids: torch.Tensor, att_mask: torch.Tensor,
                 n_heads: int, prefix: str = "") -> torch.Tensor:
    """HF BertModel(input_ids, attention_mask)[0], eval mode.  [n,T] -> [n,T,H].

    Key mask: additive (1-mask)*finfo.min, i.e. masked keys get zero probability unless a sequence has
    NO valid key (pad item), in which case the softmax degenerates to uniform -- such rows never reach
    the loss, and the CUDA path returns zeros for them instead.
    """
    n, T = input_ids.shape
    e = prefix + "embeddings."
    H = p[e + "word_embeddings.weight"].shape[1]
    dh = H // n_heads
    dt = p[e + "word_embeddings.weight"].dtype
    x = p[e + "word_embeddings.weight"][input_ids] + p[e + "token_type_embeddings.weight"][0].view(1, 1, H) \
        + p[e + "position_embeddings.weight"][:T].view(1, T, H)
    x = layer_norm(x, p[e + "LayerNorm.weight"], p[e + "LayerNorm.bias"], 1e-12)
    add = (1.0 - att_mask.to(dt)).view(n, 1, 1, T) * torch.finfo(dt).min
    l = 0
    while (prefix + f"encoder.layer.{l}.attention.self.query.weight") in p:
        q_ = prefix + f"encoder.layer.{l}."
        a = q_ + "attention.self."
        q = (x @ p[a + "query.weight"].t() + p[a + "query.bias"]).view(n, T, n_heads, dh).transpose(1, 2)
        k = (x @ p[a + "key.weight"].t() + p[a + "key.bias"]).view(n, T, n_heads, dh).transpose(1, 2)
        v = (x @ p[a + "value.weight"].t() + p[a + "value.bias"]).view(n, T, n_heads, dh).transpose(1, 2)
        s = q @ k.transpose(-2, -1) / math.sqrt(dh) + add
        pr = torch.softmax(s, dim=-1)
        o = (pr @ v).transpose(1, 2).reshape(n, T, H)
        o = o @ p[q_ + "attention.output.dense.weight"].t() + p[q_ + "attention.output.dense.bias"]
        x1 = layer_norm(x + o, p[q_ + "attention.output.LayerNorm.weight"],
                        p[q_ + "attention.output.LayerNorm.bias"], 1e-12)
        u = gelu_erf(x1 @ p[q_ + "intermediate.dense.weight"].t() + p[q_ + "intermediate.dense.bias"])
        u = u @ p[q_ + "output.dense.weight"].t() + p[q_ + "output.dense.bias"]
        x = layer_norm(x1 + u, p[q_ + "output.LayerNorm.weight"], p[q_ + "output.LayerNorm.bias"], 1e-12)
        l += 1
    return x


def text_item_encoder(p: Dict[str, torch.Tensor], items: torch.Tensor, n_heads: int,
                      prefix: str = "bert_encoder.text_encoders.title.") -> torch.Tensor:
    """items [n, 2T] (ids || attention-mask) -> GELU(fc(BERT(...)[:,0]))  [n, D]  (model/encoders.py:63-70)."""
    T = items.shape[1] // 2
    ids, am = items[:, :T], items[:, T:]
    h = bert_forward(p, ids, am, n_heads, prefix + "bert_model.")
    cls = h[:, 0] @ p[prefix + "fc.weight"].t() + p[prefix + "fc.bias"]
    return gelu_erf(cls)


# --------------------------------------------------------------------------------------------------
# scoring + debiased in-batch CE (model/model.py:32-33, 45-67)
# --------------------------------------------------------------------------------------------------

def inbatch_logits(prec_vec: torch.Tensor, score_embs: torch.Tensor, ids: torch.Tensor,
                   log_pop: torch.Tensor) -> torch.Tensor:
    """masked, debiased logits [R, C]."""
    S = prec_vec @ score_embs.t() - log_pop.view(1, -1)
    m = reject_mask_closed_form(ids)
    return torch.where(m, torch.full((), NEG_MASK, dtype=S.dtype), S)


def inbatch_ce(prec_vec, score_embs, ids, log_pop, log_mask):
    """returns (loss, logits[R,C], row_lse[R], n_valid)."""
    B, Lp1 = ids.shape
    L = Lp1 - 1
    S = inbatch_logits(prec_vec, score_embs, ids, log_pop)
    tgt = ce_labels(B, L)
    v = valid_rows(log_mask)
    lse = torch.logsumexp(S, dim=1)
    row = lse - S[torch.arange(B * L), tgt]
    n_valid = int(v.sum())
    loss = (row * v.to(row.dtype)).sum() / n_valid
    return loss, S, lse, n_valid


@dataclass
class StepOut:
    loss: torch.Tensor
    logits: torch.Tensor
    score_embs: torch.Tensor
    prec_vec: torch.Tensor
    n_valid: int


def model_forward_from_embs(p: Dict[str, torch.Tensor], E: torch.Tensor, ids: torch.Tensor, log_mask: torch.Tensor,
                            pop_prob: torch.Tensor, n_heads_user: int) -> StepOut:
    """Model.forward after the item tower (model/model.py:39-69): E [C, D] are the slot embeddings."""
    B, Lp1 = ids.shape
    L = Lp1 - 1
    log_pop = torch.log(pop_prob.to(torch.float32)[ids.reshape(-1)]).to(E.dtype)   # FloatTensor in the reference
    D = E.shape[1]
    X = E.view(B, Lp1, D)[:, :-1, :]
    Hh = sasrec_forward(p, X, log_mask, n_heads_user)
    P = Hh.reshape(B * L, D)
    loss, S, lse, n_valid = inbatch_ce(P, E, ids, log_pop, log_mask)
    return StepOut(loss, S, E, P, n_valid)


def model_forward(p: Dict[str, torch.Tensor], ids: torch.Tensor, items: torch.Tensor,
                  log_mask: torch.Tensor, pop_prob: torch.Tensor, *, use_modal: bool,
                  n_heads_user: int, n_heads_bert: int = 0) -> StepOut:
    """Full Model.forward of the TEXT package (model/model.py:31-69) in eval mode.

    ids [B, L+1] int64; items [C, 2T] int64 (modal) or [C] int64 (ID); log_mask [B, L]; pop_prob [N+1].
    """
    if use_modal:
        E = text_item_encoder(p, items, n_heads_bert)
    else:
        E = p["id_embedding.weight"][items.reshape(-1)]
    return model_forward_from_embs(p, E, ids, log_mask, pop_prob, n_heads_user)


def vision_item_encoder(image_net, images: torch.Tensor) -> torch.Tensor:
    """Vit_Encoder.forward (inbatch_sasrec_e2e_vision/model/encoders.py:30-31): GELU(image_net(pixel_values)[0]).

    `image_net` is the third-party HF SwinForImageClassification itself (the reference calls exactly this module,
    inbatch_sasrec_e2e_vision/run.py:49); it is installed on the GPU box too, so the oracle calls it directly on CPU
    instead of restating ~400 lines of modeling_swin.py.  Pinned by tests/golden/vision_tiny.pt (reference output)."""
    return gelu_erf(image_net(images)[0])


# --------------------------------------------------------------------------------------------------
# synthetic data of SURVEY.md §8(d): lives in the product package (bench.py uses it too); re-exported here
# --------------------------------------------------------------------------------------------------
from idvs.morec_b200.synth import synth_batch  # noqa: E402,F401


def bce_model_forward(p: Dict[str, torch.Tensor], sample_items: torch.Tensor, log_mask: torch.Tensor, *, use_modal: bool,
                      max_seq_len: int, n_heads_user: int, n_heads_bert: int = 0):
    """Model.forward of the BCE packages (bce_text/main-end2end/model/model.py:30-51) in eval mode: the item tower over
    every (positive, sampled negative) slot, SASRec over the positives, and two BCE-with-logits means over the valid
    rows.  sample_items: [B*(L+1)*2, 2T] token rows (modal) or int ids [B, L+1, 2]."""
    E = text_item_encoder(p, sample_items, n_heads_bert) if use_modal else p["id_embedding.weight"][sample_items.reshape(-1)]
    D = E.shape[1]
    E = E.view(-1, max_seq_len + 1, 2, D)
    pos, neg = E[:, :, 0], E[:, :, 1]
    prec = sasrec_forward(p, pos[:, :-1], log_mask, n_heads_user)
    ps = (prec * pos[:, 1:]).sum(-1)
    ns = (prec * neg[:, :-1]).sum(-1)
    idx = log_mask != 0
    return F.softplus(-ps[idx]).mean() + F.softplus(ns[idx]).mean()
