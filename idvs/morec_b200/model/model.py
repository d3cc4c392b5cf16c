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
    # The scoring / CE block ALWAYS runs on fp32 operands with the 3xTF32 tensor-core path, in every compute mode: its
    # backward dP = dS.E is a difference of nearly equal sums (sum_c softmax_c E_c - E_target, and item embeddings
    # share a large common component after GELU), so rounding dS or E to bf16 / single-pass TF32 costs ~20 % of the
    # gradient (measured: tools/diag_bf16.py) while the block is < 0.1 % of the step's FLOPs.
    _CE_META = dict(x3=True)

    def _ce_inputs(self, prec_vec, score_embs):
        """arithmetic of the scoring + CE kernel per precision mode: `fp32` -> 3xTF32 on fp32 operands (parity mode);
        `tf32` -> one kind::tf32 pass; `fp16` / `bf16` -> kind::f16 on the 16-bit activations as they are (what the
        reference's autocast does with `torch.matmul(prec_vec, score_embs.t())`, model.py:49 -- here with fp32
        accumulation and fp32 logits / softmax).  Column count grows with the number of GPUs in `global` mode, so at
        8 GPUs the 3xTF32 form cost ~1 ms per step."""
        mode = getattr(self, "compute_dtype", "fp32")
        if mode in ("fp16", "bf16") and prec_vec.dtype == score_embs.dtype and prec_vec.dtype != torch.float32:
            return dict(x3=False), prec_vec, score_embs
        return dict(x3=(mode != "tf32")), prec_vec.float(), score_embs.float()

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
        # every non-pad slot.  'slots' forces the latter, 'always' the former (under dropout duplicates of an item then
        # share one dropout mask instead of drawing independent ones -- the north star's "unique items" semantics).  Pad slots are never encoded (their embedding is 0 and
        # provably receives a zero gradient: SURVEY.md §3.2).
        self.item_dedup = "auto"
        # multi-GPU semantics (idvs/morec_b200/parallel.py): "local" = reference-exact DDP (rank-local negatives),
        # "global" = the G ranks act as one process with batch G*B (items encoded once, one all-gather of embeddings)
        self.parallel_mode = getattr(args, "parallel_mode", "local")
        self.compute_dtype = getattr(args, "compute_dtype", "fp32")
        self.set_compute_dtype(self.compute_dtype)
        # Multi-GPU: the text tower averages its own gradients, layer by layer, overlapped with its backward
        # (ops._GradSync); DistributedDataParallel (run.py:148) is told to leave those parameters alone (see the
        # `_ddp_params_and_buffers_to_ignore` property).  Everything else (SASRec, ID embedding) stays with DDP.
        self._overlap_grad_sync = False          # switched on explicitly: enable_overlap_grad_sync()
        self._grad_sync_suspended = False        # parallel.MorecDDP.no_sync()

    def enable_overlap_grad_sync(self, process_group=None):
        """Call BEFORE wrapping the model in DistributedDataParallel (parallel.wrap_ddp does): the text tower then
        averages its own gradients layer by layer, overlapped with its backward (ops._GradSync), and DDP is told to
        leave those parameters alone through the `_ddp_params_and_buffers_to_ignore` attribute its constructor reads.
        Because DDP no longer broadcasts them at construction, the tower's parameters and buffers are broadcast from
        rank 0 here.  Without this call the tower's gradients simply go through DDP's reducer like everything else."""
        import torch.distributed as dist
        if not (self.use_modal and hasattr(self, "bert_encoder")):
            return self
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            with torch.no_grad():
                for t in list(self.bert_encoder.parameters()) + list(self.bert_encoder.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        self._ddp_params_and_buffers_to_ignore = (
            [n for n, _ in self.named_parameters() if n.startswith("bert_encoder.")]
            + [n for n, _ in self.named_buffers() if n.startswith("bert_encoder.")])
        self._overlap_grad_sync = True
        return self

    def attach_optimizer(self, optimizer):
        """Let a FusedAdamW write the 16-bit compute copies of the GEMM weights in its update kernel (bf16 / fp16
        modes), so no separate cast pass runs per step.  Optional: without it the copies are refreshed by one
        multi-tensor cast launch per forward.  Any other writer of the parameters (load_state_dict, a second
        optimizer) is detected and falls back to the cast."""
        from ..optim import FusedAdamW
        opt = optimizer if isinstance(optimizer, FusedAdamW) else None
        self.user_encoder._shadows.attach(opt)
        if self.use_modal and hasattr(self, "bert_encoder"):
            te = self.bert_encoder.text_encoders['title']
            if not hasattr(te, "_prep_cache"):
                te._prep_cache = {}
            te._prep_cache.setdefault("shadows", ops.ShadowSet()).attach(opt)

    def set_item_content(self, item_content):
        """Tell the model the catalogue's token table (`item_content` [N+1, 2T]: T ids || T attention mask, the array
        run.py:93-98 builds).  Only the per-item real-token count is kept (host numpy [N+1]): it is a static property
        of the catalogue, so when the caller also passes the batch's ids as a host array (forward(..., host_ids=))
        the packed-token layout of a step is planned entirely on the host and the step contains NO device->host wait."""
        c = item_content.detach().cpu().numpy() if torch.is_tensor(item_content) else np.asarray(item_content)
        T = self.args.num_words_title
        self._item_lens = (c[:, T:2 * T] != 0).sum(axis=1).astype(np.int32)
        return self

    def set_compute_dtype(self, name):
        from .encoders import COMPUTE_DTYPES
        assert name in COMPUTE_DTYPES, name
        self.compute_dtype = name
        self.user_encoder.compute_dtype = name
        if self.use_modal:
            self.bert_encoder.text_encoders['title'].compute_dtype = name
        else:
            self.id_embedding.out_dtype = COMPUTE_DTYPES[name]

    # -------------------------------------------------------------------------------------------
    def _host_plan_inputs(self, ids_flat, host_ids, n_slots):
        """(ids, per-slot token counts) as host arrays WITHOUT touching the device, or None"""
        lens = getattr(self, "_item_lens", None)
        if host_ids is None or lens is None:
            return None
        ids_np = host_ids.detach().cpu().numpy() if torch.is_tensor(host_ids) else np.asarray(host_ids)
        ids_np = ids_np.reshape(-1).astype(np.int64, copy=False)
        if ids_np.size != n_slots:
            raise ValueError(f"host_ids has {ids_np.size} entries, the batch has {n_slots} slots")
        return ids_np, lens[ids_np]

    def _encode_items(self, ids_flat, sample_items, host_ids=None):
        if not self.use_modal:
            return self.id_embedding(sample_items.reshape(-1))
        te = self.bert_encoder
        te.text_encoders['title'].overlap_grad_sync = self._overlap_grad_sync and not self._grad_sync_suspended
        cfg = te.text_encoders['title'].bert_model.config
        dropout_on = self.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)
        if self.item_dedup == "always" or (self.item_dedup == "auto" and not dropout_on):
            # distinct non-pad ids -> first slot holding each.  One device->host fetch (ids + per-slot token counts, one
            # host wait) feeds both this index arithmetic and the token-packing plan of the text tower.
            T = self.args.num_words_title
            single = len(te.newsname) == 1 and te.attributes2start[te.newsname[0]] == 0
            if single and (sample_items.dtype != torch.int64 or sample_items.stride(1) != 1):
                sample_items = sample_items.to(torch.int64).contiguous()
            hp = self._host_plan_inputs(ids_flat, host_ids, ids_flat.numel()) if single else None
            if hp is not None:                       # host-side plan: no device->host wait in this step
                ids_np, lens_np = hp
                prep = te.text_encoders['title'].prepare()
            else:
                h = lib.d2h_begin([ids_flat, lib.mask_row_lens(sample_items, T)] if single else [ids_flat])
                prep = te.text_encoders['title'].prepare() if single else None   # weight casts behind the copy, before the wait
                got = lib.d2h_end(h)
                ids_np, lens_np = (got[0], got[1]) if single else (got[0], None)
            nz = np.nonzero(ids_np)[0]
            from ..parallel import unique_first
            _, first, inv = unique_first(ids_np[nz])
            dev = ids_flat.device
            rows = nz[first]
            E_u = te(sample_items[lib.h2d(rows, dev)], lens_np[rows] if lens_np is not None else None, prep)
            s2u = np.full(ids_np.size, -1, dtype=np.int32)
            s2u[nz] = inv.astype(np.int32)
            return ops.GatherRowsFn.apply(E_u, lib.h2d(s2u, dev), E_u.dtype)
        hp = self._host_plan_inputs(ids_flat, host_ids, ids_flat.numel()) if len(te.newsname) == 1 else None
        return te(sample_items, hp[1] if hp is not None else None)

    def forward(self, sample_items_id, sample_items, log_mask, local_rank, host_ids=None):
        """The reference's signature (model.py:31) plus one OPTIONAL argument: `host_ids`, the batch's item ids as a
        host array / CPU tensor (the DataLoader had them on the host anyway).  With it -- and set_item_content() --
        the step is planned without any device->host synchronisation."""
        if self.parallel_mode == "global" and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            return self._forward_global(sample_items_id, sample_items, log_mask, local_rank, host_ids)
        dev = sample_items_id.device
        if self._log_pop is None or self._log_pop.device != dev:
            self.pop_prob_list = self.pop_prob_list.to(dev)
            self._log_pop = torch.log(self.pop_prob_list)
        ids_flat = sample_items_id.reshape(-1)
        L = self.max_seq_len
        B = log_mask.size(0)
        D = self.args.embedding_dim
        log_pop_c = self._log_pop[ids_flat].contiguous()                  # model.py:32-33
        score_embs = self._encode_items(ids_flat, sample_items, host_ids)   # [C, D]   model.py:34-37
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
        ce_meta, P_ce, E_ce = self._ce_inputs(prec_vec, score_embs)
        loss, _ = ops.InbatchCEFn.apply(ce_meta, P_ce, E_ce, member, pad, log_pop_c, lm.reshape(-1), B, L, 0, None)
        return loss

    # -------------------------------------------------------------------------------------------
    def _host_group(self):
        """CPU (gloo) process group for the host-side exchange of the batch's item ids in `global` mode: created once,
        collectively (every rank reaches its first host-planned global step together)"""
        import torch.distributed as dist
        if getattr(self, "_gloo_group", None) is None:
            self._gloo_group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
        return self._gloo_group

    def _forward_global(self, sample_items_id, sample_items, log_mask, local_rank, host_ids=None):
        """`global` multi-GPU mode: see idvs/morec_b200/parallel.py.  Equals the reference at batch G*B in one
        process: every item of the global batch is encoded once, one all-gather of embeddings.
        With `host_ids` (+ set_item_content) the global plan is made from a HOST all-gather of the ids (gloo, 13 KB
        per rank) and the catalogue's static token counts: no device->host wait, the host runs ahead of the GPU."""
        import torch.distributed as dist
        from .. import parallel as par
        dev = sample_items_id.device
        G, rank = dist.get_world_size(), dist.get_rank()
        if self._log_pop is None or self._log_pop.device != dev:
            self.pop_prob_list = self.pop_prob_list.to(dev)
            self._log_pop = torch.log(self.pop_prob_list)
        ids_flat = sample_items_id.reshape(-1).contiguous()
        C = ids_flat.numel()
        L = self.max_seq_len
        B = log_mask.size(0)
        D = self.args.embedding_dim
        from .encoders import COMPUTE_DTYPES
        adt = COMPUTE_DTYPES[self.compute_dtype]
        ids_all = par.all_gather_small(ids_flat).reshape(-1)                      # [G*C]
        if self.use_modal:
            items_all = par.all_gather_small(sample_items.contiguous()).reshape(G * C, -1)
            te = self.bert_encoder
            te.text_encoders['title'].overlap_grad_sync = self._overlap_grad_sync and not self._grad_sync_suspended
            T = self.args.num_words_title
            single = len(te.newsname) == 1 and te.attributes2start[te.newsname[0]] == 0
            if single and items_all.dtype != torch.int64:
                items_all = items_all.to(torch.int64)
            hp = self._host_plan_inputs(ids_flat, host_ids, C) if single else None
            if hp is not None:
                # host-side plan: ids travel between the hosts, token counts come from the catalogue
                mine_t = torch.from_numpy(np.ascontiguousarray(hp[0]))
                parts = [torch.empty_like(mine_t) for _ in range(G)]
                dist.all_gather(parts, mine_t, group=self._host_group())
                ids_all_np = torch.cat(parts).numpy()
                lens_all = self._item_lens[ids_all_np]
                prep = te.text_encoders['title'].prepare()
            else:
                # ONE host wait: ids + per-slot token counts; the weight casts are issued behind the copies
                h = lib.d2h_begin([ids_all, lib.mask_row_lens(items_all, T)] if single else [ids_all])
                prep = te.text_encoders['title'].prepare() if single else None
                got = lib.d2h_end(h)
                ids_all_np, lens_all = got[0], (got[1] if single else None)
            plan = par.plan_global_batch(ids_all_np, G, rank)                     # host index arithmetic
            n_mine = int(plan.my_first_slots.size)
            enc_slots = plan.my_first_slots
            if n_mine == 0:
                # this rank owns no distinct item of the global batch (n_unique < G, e.g. the short last batch of an
                # epoch): it still runs the tower on one throw-away item whose embedding is gathered nowhere, so
                # that its backward -- and the per-layer gradient all-reduces the other ranks are issuing -- takes
                # place, with exactly zero gradient
                real = ids_all_np != 0
                if lens_all is not None:
                    real &= lens_all > 0
                enc_slots = np.flatnonzero(real)[:1].astype(plan.my_first_slots.dtype)
            if enc_slots.size > 0:
                my_items = items_all[lib.h2d(enc_slots, dev)]
                E_mine = te(my_items, lens_all[enc_slots] if (single and lens_all is not None) else None, prep)
                pad_idx = np.full(plan.u_max, -1, dtype=np.int32)
                pad_idx[:n_mine] = np.arange(n_mine, dtype=np.int32)
                E_pad = ops.GatherRowsFn.apply(E_mine, lib.h2d(pad_idx, dev), adt)   # [u_max, D]
            else:       # the whole global batch is padding: no rank runs its tower, the collectives still pair up
                E_pad = torch.zeros(plan.u_max, D, device=dev, dtype=adt)
            E_table = par.AllGatherRowsFn.apply(E_pad, dist.group.WORLD)         # ONE all-gather; bwd = reduce-scatter
            score_embs = ops.GatherRowsFn.apply(E_table, lib.h2d(plan.slot_to_row.astype(np.int32), dev), adt)
        else:
            score_embs = self.id_embedding(ids_all)                                # every rank holds the full table
        in_rows = (torch.arange(B, device=dev, dtype=torch.int32).view(B, 1) * (L + 1)
                   + torch.arange(L, device=dev, dtype=torch.int32).view(1, L)).reshape(-1) + rank * C
        X = ops.GatherRowsFn.apply(score_embs, in_rows.contiguous(), score_embs.dtype)
        prec_vec = self.user_encoder(X.view(B, L, D), log_mask, local_rank).reshape(B * L, D)
        lm = log_mask.to(torch.float32).contiguous()
        n_valid = (lm != 0).sum().to(torch.float32).reshape(1)
        dist.all_reduce(n_valid)                                                   # global valid-row count
        member, pad = lib.inbatch_mask(sample_items_id.reshape(B, L + 1).contiguous(), ids_all.contiguous(), B, L)
        log_pop_c = self._log_pop[ids_all].contiguous()
        ce_meta, P_ce, E_ce = self._ce_inputs(prec_vec, score_embs)
        loss, _ = ops.InbatchCEFn.apply(ce_meta, P_ce, E_ce, member, pad, log_pop_c, lm.reshape(-1), B, L, rank * C,
                                        n_valid)
        # x G: DistributedDataParallel averages gradients over the G ranks; the objective is the SUM of the per-rank
        # partial losses (each already divided by the global valid-row count).  The returned value is therefore G x this
        # rank's share of the global mean CE -- for logging use global_mean_loss().
        self._last_partial_loss = loss.detach()
        return loss * float(G)

    def global_mean_loss(self):
        """mean CE over the valid rows of the whole global batch of the last `global`-mode step (one scalar all-reduce;
        the number a log line should show -- forward() returns G x the rank's partial sum so that DDP's gradient
        averaging reproduces the single-process gradient)"""
        import torch.distributed as dist
        v = self._last_partial_loss.clone()
        dist.all_reduce(v)
        return v
