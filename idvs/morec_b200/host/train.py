"""run.py of the drop-in packages: `train`, `run_eval`, `setup_seed`, `main` with the reference's control flow
(inbatch_sasrec_e2e_text/run.py:27-352, inbatch_sasrec_e2e_vision/run.py) -- same command line, same log lines, same
checkpoint directory / file layout, same early stopping -- around the morec_b200 training step.

Differences that matter for speed (all switchable back, see host/params.py):
  * batches are assembled on the device from device-resident sequences and item content (DeviceBatcher) instead of
    DataLoader workers + a [B, L+1, 2T] int64 host->device copy per step;
  * the optimizer is FusedAdamW under torch.amp.GradScaler's fused protocol: unscale, overflow skip and the 16-bit
    weight copies happen inside ONE kernel and no `.item()` host wait is left in the step;
  * `torch.isnan(loss)` is checked when the loss is logged (every `steps_for_log` steps), not every step (run.py:249
    synchronises the host with the device once per step);
  * evaluation uses the sharded catalogue encoder and the fused rank kernel (host/metrics.py).
The compute dtype defaults to fp16 -- the arithmetic of the reference's `torch.cuda.amp.autocast()` loop (run.py:242).
"""
import os
import random
import re
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

from . import preprocess as PP
from .dataset import BuildTrainDataset, DeviceBatcher
from .metrics import eval_model, get_item_embeddings
from .utils import (get_checkpoint, get_time, para_and_log, report_time_eval, report_time_train, save_model, setuplogger)

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

PRETRAINED_DIR = '../../pretrained_models/'
POOLER = {'tiny': ([37, 38], 128), 'mini': ([69, 70], 256), 'small': ([69, 70], 512), 'medium': ([133, 134], 512),
          'base': ([197, 198], 768), 'large': ([389, 390], 1024)}                    # run.py:55-72


class _Ctx:
    """what the reference keeps in module globals of run.py (Log_file, Log_screen, model_dir, start_time, args)"""
    Log_file = None
    Log_screen = None
    model_dir = None
    start_time = None


def setup_seed(seed):
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


# ---------------------------------------------------------------------------------------------------------------
# encoders (run.py:29-75 text, V/run.py:27-60 vision): from_pretrained when the weights exist, random init from the
# bundled config.json otherwise (the repository ships configs and vocabularies only)
# ---------------------------------------------------------------------------------------------------------------
def load_text_encoder(args, Log_file):
    from transformers import BertConfig, BertModel, BertTokenizer
    path = PRETRAINED_DIR + args.bert_model_load
    if 'roberta' in args.bert_model_load or 'opt' in args.bert_model_load:
        raise NotImplementedError("only the BERT family is on the morec_b200 hot path (SURVEY.md §2.1)")
    Log_file.info('load bert model...')
    tokenizer = BertTokenizer.from_pretrained(path) if os.path.exists(os.path.join(path, "vocab.txt")) else None
    if os.path.isdir(path):
        config = BertConfig.from_pretrained(path, output_hidden_states=True)
    else:       # no local directory: the published shapes of the BERT family (hidden, layers, heads, intermediate)
        arch = {'tiny': (128, 2, 2, 512), 'mini': (256, 4, 4, 1024), 'small': (512, 4, 8, 2048),
                'medium': (512, 8, 8, 2048), 'base': (768, 12, 12, 3072), 'large': (1024, 24, 16, 4096)}
        h, nl, nh, it = next((v for k, v in arch.items() if k in args.bert_model_load), arch['base'])
        config = BertConfig(hidden_size=h, num_hidden_layers=nl, num_attention_heads=nh, intermediate_size=it,
                            output_hidden_states=True)
    has_weights = any(os.path.exists(os.path.join(path, n)) for n in ("pytorch_model.bin", "model.safetensors"))
    if has_weights:
        bert_model = BertModel.from_pretrained(path, config=config)
    else:
        Log_file.info(f"no weights under {path}: random initialisation from its config.json")
        bert_model = BertModel(config)
    pooler_para = []
    for key, (idx, dim) in POOLER.items():
        if key in args.bert_model_load:
            pooler_para, args.word_embedding_dim = idx, dim
    args.word_embedding_dim = config.hidden_size
    for index, (name, param) in enumerate(bert_model.named_parameters()):
        if index < args.freeze_paras_before or index in pooler_para or name.startswith("pooler."):
            param.requires_grad = False
    return bert_model, tokenizer


def load_vision_encoder(args, Log_file):
    from torch import nn
    from torch.nn.init import constant_, xavier_normal_
    from transformers import SwinConfig, SwinForImageClassification
    if 'swin' not in args.CV_model_load:
        raise NotImplementedError("only the Swin family is on the morec_b200 hot path (SURVEY.md §2.1)")
    path = PRETRAINED_DIR + args.CV_model_load
    has_weights = any(os.path.exists(os.path.join(path, n)) for n in ("pytorch_model.bin", "model.safetensors"))
    if has_weights:
        cv_model = SwinForImageClassification.from_pretrained(path)
    else:
        Log_file.info(f"no weights under {path}: random initialisation from its config.json")
        cv_model = SwinForImageClassification(SwinConfig.from_pretrained(path) if os.path.isdir(path) else SwinConfig())
    cv_model.classifier = nn.Linear(cv_model.classifier.in_features, args.embedding_dim)
    xavier_normal_(cv_model.classifier.weight.data)
    constant_(cv_model.classifier.bias.data, 0)
    for index, (name, param) in enumerate(cv_model.named_parameters()):
        if index < args.freeze_paras_before:
            param.requires_grad = False
    return cv_model


# ---------------------------------------------------------------------------------------------------------------
def _load_text_data(args, use_modal, tokenizer, Log_file):
    """-> (item_num, item_content, users_train, users_valid, users_history_for_valid, pop_prob_list)"""
    if 'None' not in args.synthetic:
        n_users, n_items = (int(x) for x in args.synthetic.split(','))
        from ..synth import synth_batch
        (item_num, _, users_train, users_valid, users_test, hist_valid, hist_test, _, _, pop) = \
            PP.synthetic_dataset(n_users, n_items, args.max_seq_len)
        content = synth_batch(1, 3, n_items, args.num_words_title, 4242, modal=True)["item_content"].numpy() \
            if use_modal else np.arange(item_num + 1)
        return item_num, content, users_train, users_valid, hist_valid, pop
    news_path = os.path.join(args.root_data_dir, args.dataset, args.news)
    beh_path = os.path.join(args.root_data_dir, args.dataset, args.behaviors)
    if use_modal:
        Log_file.info('read news...')
        b_dic, b_n2i, b_i2n = PP.read_news_bert(news_path, args, tokenizer)
    else:
        b_dic, b_n2i, b_i2n = PP.read_news(news_path)
    Log_file.info('read behaviors...')
    (item_num, item_id_to_dic, users_train, users_valid, users_test, hist_valid, hist_test, item_name_to_id, pop) = \
        PP.read_behaviors(beh_path, b_dic, b_n2i, b_i2n, args.max_seq_len, args.min_seq_len, Log_file)
    if use_modal:
        Log_file.info('combine news information...')
        parts = PP.get_doc_input_bert(item_id_to_dic, args)
        content = np.concatenate([x for x in parts if x is not None], axis=1)
    else:
        content = np.arange(item_num + 1)
    return item_num, content, users_train, users_valid, hist_valid, pop


def _param_groups(module, args, kind):
    """the reference's two AdamW groups: encoder (fine_tune_lr / fine_tune_l2_weight) and the rest (T/run.py:150-162;
    vision puts the classifier head into the recsys group, V/run.py:121-135)"""
    enc, rec = [], []
    for name, p in module.named_parameters():
        if not p.requires_grad:
            continue
        if kind == "text":
            (enc if 'bert_model' in name else rec).append(p)
        else:
            (enc if ('image_net' in name and 'fc' not in name and 'classifier' not in name) else rec).append(p)
    return [{'params': enc, 'lr': args.fine_tune_lr, 'weight_decay': args.fine_tune_l2_weight},
            {'params': rec, 'lr': args.lr, 'weight_decay': args.l2_weight}]


def train(args, use_modal, local_rank, kind="text", ctx=_Ctx):
    from ..optim import FusedAdamW
    from ..parallel import wrap_ddp
    Log_file, Log_screen = ctx.Log_file, ctx.Log_screen
    dev = torch.device("cuda", local_rank)
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1

    encoder, tokenizer = None, None
    if use_modal:
        if kind == "text":
            encoder, tokenizer = load_text_encoder(args, Log_file)
        else:
            encoder = load_vision_encoder(args, Log_file)
    if kind == "text":
        from ..model import Model
        item_num, item_content, users_train, users_valid, hist_valid, pop_prob_list = \
            _load_text_data(args, use_modal, tokenizer, Log_file)
    else:
        from ..model_vision import Model
        from .vision_data import load_vision_data
        item_num, item_content, users_train, users_valid, hist_valid, pop_prob_list = \
            load_vision_data(args, use_modal, Log_file)

    Log_file.info('build dataset...')
    use_device_batches = bool(args.device_batches) and (kind == "text" or not use_modal or torch.is_tensor(item_content))
    if use_device_batches:
        batcher = DeviceBatcher(users_train, item_content, args.max_seq_len, use_modal and kind == "text", dev)
        if kind != "text" and use_modal:
            batcher.images = item_content.to(dev)            # [N+1, 3, R, R] uint8/float store (synthetic vision runs)
    else:
        from torch.utils.data import DataLoader
        train_dataset = BuildTrainDataset(u2seq=users_train, item_content=item_content, item_num=item_num,
                                          max_seq_len=args.max_seq_len, use_modal=use_modal)
        Log_file.info('build DDP sampler...')
        train_sampler = torch.utils.data.distributed.DistributedSampler(train_dataset, num_replicas=world, rank=rank)
        Log_file.info('build dataloader...')
        train_dl = DataLoader(train_dataset, batch_size=args.batch_size, num_workers=args.num_workers, pin_memory=True,
                              sampler=train_sampler)

    Log_file.info('build model...')
    args.parallel_mode = args.parallel_mode if world > 1 else "local"
    model = Model(args, item_num, use_modal, encoder, pop_prob_list).to(dev)
    model.set_compute_dtype(args.compute_dtype)
    model.parallel_mode = args.parallel_mode
    if use_modal and kind == "text":
        model.set_item_content(item_content)       # static per-item token counts: the step needs no device->host wait

    checkpoint, ckpt_path, start_epoch, is_early_stop = None, None, 0, True
    if 'None' not in args.load_ckpt_name:
        Log_file.info('load ckpt if not None...')
        ckpt_path = get_checkpoint(ctx.model_dir, args.load_ckpt_name)
        checkpoint = torch.load(ckpt_path, map_location=torch.device('cpu'), weights_only=False)
        Log_file.info('load checkpoint...')
        model.load_state_dict(checkpoint['model_state_dict'])
        Log_file.info(f"Model loaded from {ckpt_path}")
        start_epoch = int(re.split(r'[._-]', args.load_ckpt_name)[1])
        torch.set_rng_state(checkpoint['rng_state'])
        torch.cuda.set_rng_state(checkpoint['cuda_rng_state'])
        is_early_stop = False

    Log_file.info('model.cuda()...')
    if world > 1:
        model = wrap_ddp(model, local_rank, overlap_grad_sync=(args.optimizer == "fused"))
        module = model.module
    else:
        module = model

        class _Wrap(torch.nn.Module):                       # single process: keep the `model.module` access pattern
            def __init__(self, m):
                super().__init__()
                self.module = m

            def forward(self, *a, **k):
                return self.module(*a, **k)
        model = _Wrap(module)

    if use_modal:
        groups = _param_groups(module, args, kind)
        optimizer = FusedAdamW(groups) if args.optimizer == "fused" else torch.optim.AdamW(groups)
        n_enc = len(list((module.bert_encoder.text_encoders.title.bert_model if kind == "text"
                          else module.cv_encoder.image_net).parameters()))
        Log_file.info("***** {} parameters in {}, {} parameters in model *****".format(
            n_enc, "bert" if kind == "text" else "images", len(list(module.parameters()))))
        for g in optimizer.state_dict()['param_groups']:
            Log_file.info("***** {} parameters have learning rate {}, weight_decay {} *****".format(
                len(g['params']), g['lr'], g['weight_decay']))
        req = [n for n, p in module.named_parameters() if p.requires_grad]
        frz = [n for n, p in module.named_parameters() if not p.requires_grad]
        Log_file.info("***** freeze parameters before {} in {} *****".format(args.freeze_paras_before,
                                                                          "bert" if kind == "text" else "cv"))
        Log_file.info("***** model: {} parameters require grad, {} parameters freeze *****".format(len(req), len(frz)))
    else:
        ps = [p for p in module.parameters() if p.requires_grad]
        optimizer = (FusedAdamW(ps, lr=args.lr, weight_decay=args.l2_weight) if args.optimizer == "fused"
                     else torch.optim.AdamW(ps, lr=args.lr, weight_decay=args.l2_weight))
    if args.optimizer == "fused":
        module.attach_optimizer(optimizer)
    if checkpoint is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
        Log_file.info(f"optimizer loaded from {ckpt_path}")

    Log_file.info('\n')
    Log_file.info('Training...')
    next_set_start_time = time.time()
    max_epoch, early_stop_epoch = 0, args.epoch
    max_eval_value, early_stop_count = 0, 0
    steps_for_log, _ = para_and_log(model, len(users_train), args.batch_size, Log_file, logging_num=args.logging_num,
                                    testing_num=args.testing_num)
    scaler = torch.amp.GradScaler("cuda", enabled=(args.compute_dtype == "fp16"))
    if checkpoint is not None and "scaler_state" in checkpoint:
        if scaler.is_enabled() and checkpoint["scaler_state"]:
            scaler.load_state_dict(checkpoint["scaler_state"])
        Log_file.info(f"scaler loaded from {ckpt_path}")
    Log_screen.info('{} train start'.format(args.label_screen))
    eval_gap, early_stop_gap = 1, 10

    def batches(now_epoch):
        if use_device_batches:
            order = batcher.epoch_order(now_epoch, rank, world)
            for s in range(0, order.numel(), args.batch_size):
                users = order[s:s + args.batch_size]
                ids, items, lm = batcher.batch(users)
                if kind != "text" and use_modal:
                    items = batcher.images[ids.reshape(-1)].float()
                yield ids, items, lm, batcher.host_ids(users.numpy())
        else:
            train_dl.sampler.set_epoch(now_epoch)
            for ids, items, lm in train_dl:                   # the DataLoader's ids are host tensors already
                yield ids.to(dev, non_blocking=True), items.to(dev, non_blocking=True), lm.to(dev, non_blocking=True), ids

    for ep in range(args.epoch):
        now_epoch = start_epoch + ep + 1
        Log_file.info('\n')
        Log_file.info('epoch {} start'.format(now_epoch))
        Log_file.info('')
        loss, batch_index, need_break = torch.zeros((), device=dev), 1, False
        model.train()
        for sample_items_id, sample_items, log_mask, host_ids in batches(now_epoch):
            if kind == "text":
                sample_items = sample_items.view(-1, sample_items.size(-1)) if use_modal else sample_items.view(-1)
            elif use_modal:
                sample_items = sample_items.view(-1, *sample_items.shape[-3:])
            else:
                sample_items = sample_items.view(-1)
            sample_items_id = sample_items_id.view(-1)

            optimizer.zero_grad(set_to_none=True)
            bz_loss = model(sample_items_id, sample_items, log_mask, local_rank, host_ids=host_ids)
            loss += bz_loss.detach().float()
            scaler.scale(bz_loss).backward()
            scaler.step(optimizer)
            scaler.update()

            if batch_index % steps_for_log == 0:
                if torch.isnan(loss):
                    need_break = True
                    break
                Log_file.info('cnt: {}, Ed: {}, batch loss: {:.5f}, sum loss: {:.5f}'.format(
                    batch_index, batch_index * args.batch_size, float(loss) / batch_index, float(loss)))
            batch_index += 1
            if args.max_steps and batch_index > args.max_steps:
                break
        if torch.isnan(loss):
            need_break = True

        if not need_break and now_epoch % eval_gap == 0:
            Log_file.info('')
            max_eval_value, max_epoch, early_stop_epoch, early_stop_count, need_break, need_save = \
                run_eval(now_epoch, max_epoch, early_stop_epoch, max_eval_value, early_stop_count, model, item_content,
                         hist_valid, users_valid, args.eval_batch_size, item_num, use_modal, args.mode, is_early_stop,
                         local_rank, early_stop_gap, args=args, ctx=ctx)
            model.train()
            if use_modal and need_save and rank == 0:
                save_model(now_epoch, model, ctx.model_dir, optimizer, torch.get_rng_state(), torch.cuda.get_rng_state(),
                           scaler, Log_file)
        Log_file.info('')
        next_set_start_time = report_time_train(batch_index, now_epoch, float(loss), next_set_start_time, ctx.start_time,
                                                Log_file)
        Log_screen.info('{} training: epoch {}/{}'.format(args.label_screen, now_epoch, args.epoch))
        if need_break:
            break
    Log_file.info('\n')
    Log_file.info('%' * 90)
    Log_file.info(' max eval Hit10 {:0.5f}  in epoch {}'.format(max_eval_value * 100, max_epoch))
    Log_file.info(' early stop in epoch {}'.format(early_stop_epoch))
    Log_file.info('the End')
    Log_screen.info('{} train end in epoch {}'.format(args.label_screen, early_stop_epoch))
    return dict(max_eval_value=max_eval_value, max_epoch=max_epoch, last_loss=float(loss) / max(batch_index - 1, 1),
                model=module, optimizer=optimizer, scaler=scaler)


def run_eval(now_epoch, max_epoch, early_stop_epoch, max_eval_value, early_stop_count, model, item_content,
             user_history, users_eval, batch_size, item_num, use_modal, mode, is_early_stop, local_rank, early_stop_gap,
             args=None, ctx=_Ctx):
    Log_file = ctx.Log_file
    t0 = time.time()
    Log_file.info('Validating...')
    item_embeddings = get_item_embeddings(model, item_content, batch_size, args, use_modal, local_rank)
    valid_Hit10 = eval_model(model, user_history, users_eval, item_embeddings, batch_size, args, item_num, Log_file, mode,
                             local_rank)
    report_time_eval(t0, Log_file)
    Log_file.info('')
    need_break = need_save = False
    if valid_Hit10 > max_eval_value:
        max_eval_value, max_epoch, early_stop_count, need_save = valid_Hit10, now_epoch, 0, True
    else:
        early_stop_count += 1
        if early_stop_count > early_stop_gap:
            if is_early_stop:
                need_break = True
            early_stop_epoch = now_epoch
    return max_eval_value, max_epoch, early_stop_epoch, early_stop_count, need_break, need_save


def main(kind="text", argv=None):
    """`python run.py ...` under torch.distributed.launch / torchrun, as the reference's launchers start it"""
    from .params import parse_args
    args = parse_args(argv, kind)
    local_rank = max(args.local_rank, 0)
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group(backend='nccl', device_id=torch.device("cuda", local_rank))
    setup_seed(12345)
    gpus = dist.get_world_size()
    if 'modal' in args.item_tower:
        is_use_modal = True
        model_load = args.bert_model_load if kind == "text" else args.CV_model_load
        dir_label = str(args.item_tower) + f'_{model_load}_freeze_{args.freeze_paras_before}'
    else:
        is_use_modal, model_load, dir_label = False, 'id', str(args.item_tower)
    batch_size = args.batch_size * gpus
    log_paras = f'{model_load}_ed_{args.embedding_dim}_bs_{batch_size}_lr_{args.lr}_Flr_{args.fine_tune_lr}' \
                f'_L2_{args.l2_weight}_FL2_{args.fine_tune_l2_weight}'
    _Ctx.model_dir = os.path.join('./checkpoint_' + dir_label, 'cpt_' + log_paras)
    time_run = time.strftime('-%Y%m%d-%H%M%S', time.localtime())
    args.label_screen = args.label_screen + time_run
    _Ctx.Log_file, _Ctx.Log_screen = setuplogger(dir_label, log_paras, time_run, args.mode, dist.get_rank(), args.behaviors)
    _Ctx.Log_file.info(args)
    Path(_Ctx.model_dir).mkdir(parents=True, exist_ok=True)
    _Ctx.start_time = time.time()
    out = None
    if 'train' in args.mode:
        out = train(args, is_use_modal, local_rank, kind=kind)
    hour, minu, secon = get_time(_Ctx.start_time, time.time())
    _Ctx.Log_file.info("##### (time) all: {} hours {} minutes {} seconds #####".format(hour, minu, secon))
    return out
