"""Parameter containers of the SASRec transformer, mirroring inbatch_sasrec_e2e_text/model/modules.py.

Module / parameter names, construction order and initialisation are identical to the reference (so state dicts and
seeded construction are interchangeable), but no module here computes anything in its own forward: the math runs in
`ops.SasrecFn` on the morec_b200 CUDA kernels.
"""
import torch
import torch.nn as nn


class PositionwiseFeedForward(nn.Module):          # reference: modules.py:5-17
    def __init__(self, d_model, d_inner, dropout):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_inner)
        self.w_2 = nn.Linear(d_inner, d_model)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.dropout_p = dropout


class MultiHeadedAttention(nn.Module):             # reference: modules.py:34-63
    def __init__(self, n_heads, d_model, dropout):
        super().__init__()
        assert d_model % n_heads == 0
        self.d_model = d_model
        self.d_k = d_model // n_heads
        self.n_heads = n_heads
        self.d_v = self.d_k
        self.w_Q = nn.Linear(d_model, n_heads * self.d_k, bias=False)
        self.w_K = nn.Linear(d_model, n_heads * self.d_k, bias=False)
        self.w_V = nn.Linear(d_model, n_heads * self.d_v, bias=False)
        self.fc = nn.Linear(n_heads * self.d_v, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.dropout_p = dropout


class TransformerBlock(nn.Module):                 # reference: modules.py:66-75
    def __init__(self, d_model, n_heads, d_inner, dropout):
        super().__init__()
        self.multi_head_attention = MultiHeadedAttention(n_heads=n_heads, d_model=d_model, dropout=dropout)
        self.feed_forward = PositionwiseFeedForward(d_model=d_model, d_inner=d_inner, dropout=dropout)


class TransformerEncoder(nn.Module):               # reference: modules.py:78-96
    def __init__(self, n_vocab, n_position, d_model, n_heads, dropout, n_layers):
        super().__init__()
        self.position_embedding = nn.Embedding(n_position, d_model)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.transformer_blocks = nn.ModuleList(
            [TransformerBlock(d_model=d_model, n_heads=n_heads, d_inner=d_model * 4, dropout=dropout)
             for _ in range(n_layers)])
        self.n_heads = n_heads
        self.n_position = n_position
        self.dropout_p = dropout

    def flat_params(self):
        """order expected by ops.SasrecFn"""
        ps = [self.position_embedding.weight, self.layer_norm.weight, self.layer_norm.bias]
        for blk in self.transformer_blocks:
            a, f = blk.multi_head_attention, blk.feed_forward
            ps += [a.w_Q.weight, a.w_K.weight, a.w_V.weight, a.fc.weight, a.layer_norm.weight, a.layer_norm.bias,
                   f.w_1.weight, f.w_1.bias, f.w_2.weight, f.w_2.bias, f.layer_norm.weight, f.layer_norm.bias]
        return ps
