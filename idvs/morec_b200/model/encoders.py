"""User / item encoders, mirroring inbatch_sasrec_e2e_text/model/encoders.py on the morec_b200 CUDA kernels."""
import itertools

import numpy as np
import torch
import torch.nn as nn
from torch.nn.init import xavier_normal_, constant_

from .. import lib, ops
from .modules import TransformerEncoder

_call_counter = itertools.count(1)


def _drop_ctx(training, p_hidden, p_attn):
    if not training or (p_hidden <= 0 and p_attn <= 0):
        return ops.DropCtx()
    return ops.DropCtx(p_hidden=float(p_hidden), p_attn=float(p_attn), seed=int(torch.initial_seed()),
                       base=next(_call_counter) << 44)


COMPUTE_DTYPES = {"fp32": torch.float32, "tf32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def _adt(module):
    """activation storage dtype of a compute mode: 'fp32' (3xTF32 parity) and 'tf32' store fp32, 'bf16' stores bf16,
    'fp16' stores IEEE half (the arithmetic of the reference's own torch.cuda.amp.autocast path, run.py:242)"""
    return COMPUTE_DTYPES[getattr(module, "compute_dtype", "fp32")]


def _x3(module):
    return getattr(module, "compute_dtype", "fp32") == "fp32"


class User_Encoder(torch.nn.Module):               # reference: encoders.py:7-28
    def __init__(self, item_num, max_seq_len, item_dim, num_attention_heads, dropout, n_layers):
        super().__init__()
        self.transformer_encoder = TransformerEncoder(n_vocab=item_num, n_position=max_seq_len, d_model=item_dim,
                                                      n_heads=num_attention_heads, dropout=dropout, n_layers=n_layers)
        self.apply(self._init_weights)
        self.compute_dtype = "fp32"
        self._qkv_groups = None
        self._shadows = ops.ShadowSet()

    def _fused_qkv(self):
        if self._qkv_groups is None:
            self._qkv_groups = [ops.FusedParamGroup([b.multi_head_attention.w_Q.weight, b.multi_head_attention.w_K.weight,
                                                     b.multi_head_attention.w_V.weight])
                                for b in self.transformer_encoder.transformer_blocks]
        return [g.buffer() for g in self._qkv_groups]

    def _compute_weights(self, wqkv, adt):
        """16-bit copies of the GEMM weights (4 per block), kept by a ShadowSet (one cast launch, or none at all when
        the attached FusedAdamW writes them)"""
        srcs, owners = [], []
        for buf, b in zip(wqkv, self.transformer_encoder.transformer_blocks):
            m, f = b.multi_head_attention, b.feed_forward
            D = m.w_Q.weight.shape[0]
            srcs += [buf.detach(), m.fc.weight.detach(), f.w_1.weight.detach(), f.w_2.weight.detach()]
            owners += [[(m.w_Q.weight, 0, D), (m.w_K.weight, D, 2 * D), (m.w_V.weight, 2 * D, 3 * D)],
                       [(m.fc.weight, 0, m.fc.weight.shape[0])], [(f.w_1.weight, 0, f.w_1.weight.shape[0])],
                       [(f.w_2.weight, 0, f.w_2.weight.shape[0])]]
        if not all(t.is_contiguous() and t.data_ptr() % 16 == 0 for t in srcs):
            return None
        return self._shadows.get(srcs, owners, adt)

    def _init_weights(self, module):                # reference: encoders.py:14-21
        if isinstance(module, nn.Embedding):
            xavier_normal_(module.weight.data)
        elif isinstance(module, nn.Linear):
            xavier_normal_(module.weight.data)
            if module.bias is not None:
                constant_(module.bias.data, 0)

    def forward(self, input_embs, log_mask, local_rank=None):
        """input_embs [B, L, D], log_mask [B, L] -> [B, L, D]   (reference: encoders.py:23-28; also the eval entry
        data_utils/metrics.py:95)."""
        te = self.transformer_encoder
        B, L, D = input_embs.shape
        adt = _adt(self)
        drop = _drop_ctx(self.training, te.dropout_p, te.dropout_p)
        wqkv = self._fused_qkv()
        meta = dict(n_blocks=len(te.transformer_blocks), n_heads=te.n_heads, L=L, adt=adt, drop=drop, x3=_x3(self),
                    wqkv=wqkv, cw=self._compute_weights(wqkv, adt) if adt != torch.float32 else None)
        X = input_embs.reshape(B * L, D)
        if X.dtype != adt:
            X = X.to(adt)
        lm = log_mask.to(torch.float32).contiguous()
        out = ops.SasrecFn.apply(meta, X, lm, *te.flat_params())
        return out.view(B, L, D)


class Text_Encoder(torch.nn.Module):                # reference: encoders.py:53-70
    """GELU(fc(BertModel(ids, mask)[0][:, 0])).  `bert_model` is the caller's HF BertModel: it is kept as a
    sub-module (parameter objects, order and names unchanged: run.py:73-75,155,165 address them), but its forward is
    never called -- the encoder runs on packed tokens in ops.BertTowerFn."""

    def __init__(self, bert_model, item_embedding_dim, word_embedding_dim):
        super().__init__()
        self.bert_model = bert_model
        self.fc = nn.Linear(word_embedding_dim, item_embedding_dim)
        self.compute_dtype = "fp32"
        self._qkv_groups = None

    def _fused_qkv(self):
        if self._qkv_groups is None:
            self._qkv_groups = [(ops.FusedParamGroup([l.attention.self.query.weight, l.attention.self.key.weight,
                                                      l.attention.self.value.weight]),
                                 ops.FusedParamGroup([l.attention.self.query.bias, l.attention.self.key.bias,
                                                      l.attention.self.value.bias]))
                                for l in self.bert_model.encoder.layer]
        return [w.buffer() for w, _ in self._qkv_groups], [b.buffer() for _, b in self._qkv_groups]

    def _flat_params(self):
        bm = self.bert_model
        e = bm.embeddings
        ps = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight,
              e.LayerNorm.weight, e.LayerNorm.bias]
        for lyr in bm.encoder.layer:
            a, so = lyr.attention.self, lyr.attention.output
            ps += [a.query.weight, a.query.bias, a.key.weight, a.key.bias, a.value.weight, a.value.bias,
                   so.dense.weight, so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias,
                   lyr.intermediate.dense.weight, lyr.intermediate.dense.bias,
                   lyr.output.dense.weight, lyr.output.dense.bias, lyr.output.LayerNorm.weight, lyr.output.LayerNorm.bias]
        ps += [self.fc.weight, self.fc.bias]
        return ps

    def prepare(self):
        """batch-independent per-forward state (flat parameter list, fused QKV buffers, compute-dtype weights); issued
        before the packing plan's host sync so it overlaps the tail of the previous step"""
        if not hasattr(self, "_prep_cache"):
            self._prep_cache = {}
        c = self._prep_cache
        bm = self.bert_model
        key = (id(bm), len(bm.encoder.layer), id(bm.embeddings.word_embeddings.weight), id(self.fc.weight),
               id(bm.encoder.layer[-1].output.dense.weight))
        if c.get("flat_key") != key:                    # the ~200 Parameter OBJECTS only change under module surgery
            c["flat"], c["flat_key"] = self._flat_params(), key
        flat = c["flat"]
        wqkv, bqkv = self._fused_qkv()                  # re-validated every call (.to() re-allocates parameters)
        cw = ops.prepare_tower_weights(wqkv, flat, _adt(self), self._prep_cache)
        lay = c.get("layout")
        if lay is None or lay.key != ops.TowerLayout.make_key(cw, bqkv) or c.get("layout_flat") is not flat:
            lay = c["layout"] = ops.TowerLayout(flat, cw, bqkv)      # pointer tables + gradient-arena layout of the sequencer
            c["layout_flat"] = flat
        return dict(flat=flat, wqkv=wqkv, bqkv=bqkv, cw=cw, layout=lay)

    def forward(self, text, lens_host=None, prep=None):
        """text [n, 2T] int (ids || attention mask) -> [n, D].  Items whose mask is all zero (pad item) return 0.
        lens_host: optional host copy (numpy int [n]) of the real-token count per row when the caller already fetched
        it; prep: result of prepare() when the caller already issued it."""
        n, two_t = text.shape
        T = two_t // 2
        cfg = self.bert_model.config
        adt = _adt(self)
        D = self.fc.weight.shape[0]
        dev = text.device
        if text.dtype != torch.int64 or text.stride(1) != 1:
            text = text.to(torch.int64).contiguous()
        # ---- packing plan: ONE device->host round trip (n token counts), prefix sum in numpy, token placement on device
        if lens_host is None:
            h = lib.d2h_begin([lib.mask_row_lens(text, T)])         # size-determining sync ...
            if prep is None:
                prep = self.prepare()                               # ... with the weight casts issued behind the copy
            lens = lib.d2h_end(h)[0]
        else:
            lens = lens_host
            if prep is None:
                prep = self.prepare()
        enc_rows = np.flatnonzero(lens > 0).astype(np.int32)
        n_enc = int(enc_rows.size)
        if n_enc == 0:
            return torch.zeros(n, D, device=dev, dtype=adt)
        plan_np = np.empty(2 * n_enc + 1, dtype=np.int32)
        plan_np[:n_enc] = enc_rows
        plan_np[n_enc] = 0
        np.cumsum(lens[enc_rows], out=plan_np[n_enc + 1:])
        n_tok = int(plan_np[-1])
        plan = lib.h2d(plan_np, dev)
        cu = plan[n_enc:]
        tok_ids, tok_pos = lib.pack_tokens(text, T, plan[:n_enc], cu, n_tok)
        cls_rows = cu[:-1]
        drop = _drop_ctx(self.training, cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob)
        meta = dict(n_layers=cfg.num_hidden_layers, n_heads=cfg.num_attention_heads, eps=cfg.layer_norm_eps, max_len=T,
                    adt=adt, drop=drop, x3=_x3(self))
        meta["wqkv"], meta["bqkv"], meta["cw"], meta["layout"] = prep["wqkv"], prep["bqkv"], prep["cw"], prep.get("layout")
        meta["grad_sync"] = getattr(self, "overlap_grad_sync", False)
        E = ops.BertTowerFn.apply(meta, tok_ids, tok_pos, cu, cls_rows, *prep["flat"])
        if n_enc == n:
            return E
        s2e = np.full(n, -1, dtype=np.int32)
        s2e[enc_rows] = np.arange(n_enc, dtype=np.int32)
        return ops.GatherRowsFn.apply(E, lib.h2d(s2e, dev), adt)


class Bert_Encoder(torch.nn.Module):                # reference: encoders.py:73-117
    def __init__(self, args, bert_model):
        super().__init__()
        self.args = args
        self.attributes2length = {'title': args.num_words_title * 2, 'abstract': args.num_words_abstract * 2,
                                  'body': args.num_words_body * 2}
        for key in list(self.attributes2length.keys()):
            if key not in args.news_attributes:
                self.attributes2length[key] = 0
        keys = list(self.attributes2length.keys())
        self.attributes2start = {key: sum(self.attributes2length[k] for k in keys[:keys.index(key)]) for key in keys}
        assert len(args.news_attributes) > 0
        if 'opt' in args.bert_model_load:
            raise NotImplementedError("OPT mean-pooling text tower is outside the hot path (SURVEY.md §2.1 row 1)")
        self.text_encoders = nn.ModuleDict({'title': Text_Encoder(bert_model, args.embedding_dim, args.word_embedding_dim)})
        self.newsname = [name for name in set(args.news_attributes) & {'title', 'abstract', 'body'}]

    def forward(self, news, lens_host=None, prep=None):
        vecs = [self.text_encoders['title'](torch.narrow(news, 1, self.attributes2start[name], self.attributes2length[name]),
                                            lens_host if len(self.newsname) == 1 else None, prep)
                for name in self.newsname]
        if len(vecs) == 1:
            return vecs[0]
        return torch.mean(torch.stack(vecs, dim=1), dim=1)
