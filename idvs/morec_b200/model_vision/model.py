"""`Model` of the vision package (inbatch_sasrec_e2e_vision/model/model.py:7-73) on the morec_b200 kernels: same
constructor / forward / sub-module names (`cv_encoder.image_net`, `user_encoder`, `id_embedding`)."""
import numpy as np
import torch
from torch import nn
from torch.nn.init import xavier_normal_

from .. import lib, ops
from ..model.model import Model as _TextModel, _IdEmbedding
from ..model.encoders import User_Encoder
from .encoders import Vit_Encoder


class Model(_TextModel):
    def __init__(self, args, item_num, use_modal, image_net, pop_prob_list):
        torch.nn.Module.__init__(self)
        self.args = args
        self.use_modal = use_modal
        self.max_seq_len = args.max_seq_len
        self.pop_prob_list = torch.FloatTensor(pop_prob_list)
        self._log_pop = None
        self.user_encoder = User_Encoder(item_num=item_num, max_seq_len=args.max_seq_len, item_dim=args.embedding_dim,
                                         num_attention_heads=args.num_attention_heads, dropout=args.drop_rate,
                                         n_layers=args.transformer_block)
        if self.use_modal:
            name = getattr(args, "CV_model_load", "swin")
            if 'swin' not in name and 'beit' not in name:
                raise NotImplementedError("only the Swin tower is on the hot path (SURVEY.md §2.1 row 2)")
            self.cv_encoder = Vit_Encoder(image_net=image_net)
        else:
            self.id_embedding = _IdEmbedding(item_num + 1, args.embedding_dim, padding_idx=0)
            xavier_normal_(self.id_embedding.weight.data)
        self.item_dedup = "auto"
        self.parallel_mode = getattr(args, "parallel_mode", "local")
        self.compute_dtype = getattr(args, "compute_dtype", "fp32")
        self.set_compute_dtype(self.compute_dtype)

    # the gradient-sync flags of the text Model (its ops._GradSync only covers the BERT tower); images are never
    # exchanged between ranks, so the vision Model always runs the reference's DDP semantics
    _overlap_grad_sync = False
    _grad_sync_suspended = False

    def forward(self, sample_items_id, sample_items, log_mask, local_rank, host_ids=None):
        if self.use_modal and self.parallel_mode == "global":
            self.parallel_mode = "local"
        return super().forward(sample_items_id, sample_items, log_mask, local_rank, host_ids)

    def set_compute_dtype(self, name):
        from ..model.encoders import COMPUTE_DTYPES
        assert name in COMPUTE_DTYPES, name
        self.compute_dtype = name
        self.user_encoder.compute_dtype = name
        if self.use_modal:
            self.cv_encoder.compute_dtype = name
        else:
            self.id_embedding.out_dtype = COMPUTE_DTYPES[name]

    def _encode_items(self, ids_flat, sample_items, host_ids=None):
        if not self.use_modal:
            return self.id_embedding(sample_items.reshape(-1))
        # pad slots hold an all-zero image (inbatch_sasrec_e2e_vision/data_utils/dataset.py:86) and never matter:
        # encode each distinct non-pad item once (drop-path is per image, so this is exact in distribution per item)
        dev = ids_flat.device
        if host_ids is not None:                     # the caller had the ids on the host: no device->host wait
            ids_np = (host_ids.detach().cpu().numpy() if torch.is_tensor(host_ids) else np.asarray(host_ids)).reshape(-1)
        else:
            ids_np = lib.d2h_many([ids_flat])[0]
        nz = np.nonzero(ids_np)[0]
        # 'auto' (the text Model's rule): duplicates are encoded once only when that is exact -- under stochastic depth
        # (training, drop_path_rate > 0) the reference draws an independent mask per slot, so every slot is encoded;
        # 'always' shares one mask between the duplicates of a step (what bench.py measures), 'slots' never dedups
        cfgv = getattr(getattr(self.cv_encoder, "image_net", None), "config", None)
        stochastic = self.training and float(getattr(cfgv, "drop_path_rate", 0.0) or 0.0) > 0.0
        if self.item_dedup == "slots" or (self.item_dedup == "auto" and stochastic):
            rows = nz
            s2u = np.full(ids_np.size, -1, dtype=np.int32)
            s2u[nz] = np.arange(nz.size, dtype=np.int32)
        else:
            _, first, inv = np.unique(ids_np[nz], return_index=True, return_inverse=True)
            rows = nz[first]
            s2u = np.full(ids_np.size, -1, dtype=np.int32)
            s2u[nz] = inv.astype(np.int32)
        E_u = self.cv_encoder(sample_items[lib.h2d(rows, dev)])
        return ops.GatherRowsFn.apply(E_u, lib.h2d(s2u, dev), E_u.dtype)
