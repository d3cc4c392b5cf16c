"""`Model` of the in-batch SASRec + modality-encoder recommender on the morec_b200 CUDA kernels.

Drop-in for inbatch_sasrec_e2e_text/model/model.py:7-69: same constructor, same forward signature and return value
(0-dim loss tensor that supports .backward() through GradScaler), same sub-module / parameter names.
"""
import numpy as np
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from .. import lib, ops
from .encoders import Bert_Encoder, User_Encoder


class _IdEmbedding(nn.Embedding):
    """nn.Embedding(item_num + 1, D, padding_idx=0) whose forward is a morec gather (reference: model.py:27,37)."""

    def forward(self, ids):
        # row 0 is looked up like any other row (it is NOT zero after the reference's xavier re-init, model.py:28);
        # pad slots provably receive an exactly-zero gradient, which matches padding_idx=0 semantics.
        shape = ids.shape
        idx = ids.reshape(-1).to(torch.int32).contiguous()
        out = ops.GatherRowsFn.apply(self.weight, idx, getattr(self, "out_dtype", self.weight.dtype))
        return out.view(*shape, self.weight.shape[1])


class Model(torch.nn.Module):
    def __init__(self, args, item_num, use_modal, bert_model, pop_prob_list):
        super().__init__()
        self.args = args
        self.use_modal = use_modal
        self.max_seq_len = args.max_seq_len
        self.pop_prob_list = torch.FloatTensor(pop_prob_list)            # plain attribute, as in model.py:14
        self._log_pop = None
        self.user_encoder = User_Encoder(item_num=item_num, max_seq_len=args.max_seq_len, item_dim=args.embedding_dim,
                                         num_attention_heads=args.num_attention_heads, dropout=args.drop_rate,
                                         n_layers=args.transformer_block)
        if self.use_modal:
            self.bert_encoder = Bert_Encoder(args=args, bert_model=bert_model)
        else:
            self.id_embedding = _IdEmbedding(item_num + 1, args.embedding_dim, padding_idx=0)
            xavier_normal_(self.id_embedding.weight.data)
        # 'auto': encode each distinct non-pad item once when that is exact (no dropout active), otherwise encode
        # every non-pad slot.  'slots' forces the latter.  Pad slots are never encoded (their embedding is 0 and
        # provably receives a zero gradient: SURVEY.md §3.2).
        self.item_dedup = "auto"
        self.compute_dtype = getattr(args, "compute_dtype", "fp32")
        self.set_compute_dtype(self.compute_dtype)

    def set_compute_dtype(self, name):
        assert name in ("fp32", "tf32", "bf16")
        self.compute_dtype = name
        self.user_encoder.compute_dtype = name
        if self.use_modal:
            self.bert_encoder.text_encoders['title'].compute_dtype = name
        else:
            self.id_embedding.out_dtype = torch.bfloat16 if name == "bf16" else torch.float32

    # -------------------------------------------------------------------------------------------
    def _encode_items(self, ids_flat, sample_items):
        if not self.use_modal:
            return self.id_embedding(sample_items.reshape(-1))
        te = self.bert_encoder
        cfg = te.text_encoders['title'].bert_model.config
        dropout_on = self.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)
        if self.item_dedup == "auto" and not dropout_on:
            uniq, inv = torch.unique(ids_flat, return_inverse=True)
            # first slot holding each distinct id
            C = ids_flat.numel()
            first = torch.full((uniq.numel(),), C, device=ids_flat.device, dtype=torch.long)
            first.scatter_reduce_(0, inv, torch.arange(C, device=ids_flat.device), reduce="amin")
            E_u = te(sample_items[first])
            adt = E_u.dtype
            return ops.GatherRowsFn.apply(E_u, inv.to(torch.int32).contiguous(), adt)
        return te(sample_items)

    def forward(self, sample_items_id, sample_items, log_mask, local_rank):
        dev = sample_items_id.device
        if self._log_pop is None or self._log_pop.device != dev:
            self.pop_prob_list = self.pop_prob_list.to(dev)
            self._log_pop = torch.log(self.pop_prob_list)
        ids_flat = sample_items_id.reshape(-1)
        L = self.max_seq_len
        B = log_mask.size(0)
        D = self.args.embedding_dim
        log_pop_c = self._log_pop[ids_flat].contiguous()                  # model.py:32-33
        score_embs = self._encode_items(ids_flat, sample_items)          # [C, D]   model.py:34-37
        # input_embs[:, :-1]  (model.py:39-41)
        key = (B, L, str(dev))
        if getattr(self, "_in_rows_key", None) != key:
            self._in_rows = (torch.arange(B, device=dev, dtype=torch.int32).view(B, 1) * (L + 1)
                             + torch.arange(L, device=dev, dtype=torch.int32).view(1, L)).reshape(-1).contiguous()
            self._in_rows_key = key
        in_rows = self._in_rows
        X = ops.GatherRowsFn.apply(score_embs, in_rows, score_embs.dtype)
        prec_vec = self.user_encoder(X.view(B, L, D), log_mask, local_rank).reshape(B * L, D)
        # in-batch debiased CE (model.py:45-67)
        lm = log_mask.to(torch.float32).contiguous()
        member, pad = lib.inbatch_mask(sample_items_id.reshape(B, L + 1).contiguous(), ids_flat.contiguous(), B, L)
        loss, _ = ops.InbatchCEFn.apply(dict(x3=self.compute_dtype == "fp32"), prec_vec, score_embs, member, pad, log_pop_c, lm.reshape(-1), B, L, 0, None)
        return loss
