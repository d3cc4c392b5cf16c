"""`Model` of the BCE packages (bce_text/main-end2end/model/model.py:7-51) on the morec_b200 kernels: the same item
encoder and SASRec user encoder as the in-batch packages with a sampled-negative BCE head -- row-wise dot products,
no [R, C] score matrix.  Same constructor / forward signature and sub-module names."""
import torch
from torch.nn.init import xavier_normal_

from .. import lib
from ..model.encoders import COMPUTE_DTYPES, Bert_Encoder, User_Encoder
from ..model.model import _IdEmbedding


class BceFn(torch.autograd.Function):
    """loss = BCEWithLogits(pos, 1) + BCEWithLogits(neg, 0) over the valid rows (model.py:44-51)"""

    @staticmethod
    def forward(ctx, P, Epos, Eneg, log_mask):
        P, Epos, Eneg = P.detach().contiguous(), Epos.detach().contiguous(), Eneg.detach().contiguous()
        pos, neg, sum_cnt = lib.bce_fwd(P, Epos, Eneg, log_mask)
        ctx.save_for_backward(P, Epos, Eneg, log_mask, pos, neg, sum_cnt)
        return (sum_cnt[0] / sum_cnt[1]).reshape(())

    @staticmethod
    def backward(ctx, dloss):
        P, Epos, Eneg, log_mask, pos, neg, sum_cnt = ctx.saved_tensors
        g = dloss.detach().reshape(1).to(torch.float32).contiguous()
        dP, dEp, dEn = lib.bce_bwd(P, Epos, Eneg, log_mask, pos, neg, g, sum_cnt)
        return dP, dEp, dEn, None


class Model(torch.nn.Module):
    def __init__(self, args, item_num, use_modal, bert_model):
        super().__init__()
        self.args = args
        self.use_modal = use_modal
        self.max_seq_len = args.max_seq_len + 1
        self.user_encoder = User_Encoder(item_num=item_num, max_seq_len=args.max_seq_len, item_dim=args.embedding_dim,
                                         num_attention_heads=args.num_attention_heads, dropout=args.drop_rate,
                                         n_layers=args.transformer_block)
        if self.use_modal:
            self.bert_encoder = Bert_Encoder(args=args, bert_model=bert_model)
        else:
            self.id_embedding = _IdEmbedding(item_num + 1, args.embedding_dim, padding_idx=0)
            xavier_normal_(self.id_embedding.weight.data)
        self.compute_dtype = "fp32"
        self.set_compute_dtype(getattr(args, "compute_dtype", "fp32"))

    def set_compute_dtype(self, name):
        assert name in COMPUTE_DTYPES, name
        self.compute_dtype = name
        self.user_encoder.compute_dtype = name
        if self.use_modal:
            self.bert_encoder.text_encoders['title'].compute_dtype = name
        else:
            self.id_embedding.out_dtype = COMPUTE_DTYPES[name]

    def forward(self, sample_items, log_mask, local_rank):
        """sample_items: [B*(L+1)*2, 2T] token rows (modal) or [B, L+1, 2] ids -- slot (b, t, 0) = the user's item,
        (b, t, 1) = its sampled negative (bce_text/main-end2end/data_utils/dataset.py:20-50)"""
        D = self.args.embedding_dim
        E = self.bert_encoder(sample_items) if self.use_modal else self.id_embedding(sample_items.reshape(-1))
        E = E.view(-1, self.max_seq_len, 2, D)
        B, L = E.shape[0], self.max_seq_len - 1
        pos, neg = E[:, :, 0], E[:, :, 1]
        prec = self.user_encoder(pos[:, :-1].contiguous(), log_mask, local_rank).reshape(B * L, D)
        lm = log_mask.to(torch.float32).reshape(-1).contiguous()
        return BceFn.apply(prec, pos[:, 1:].reshape(B * L, D), neg[:, :-1].reshape(B * L, D), lm)
