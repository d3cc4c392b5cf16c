"""Logging, timing and checkpoint helpers with the reference's names and file layout
(inbatch_sasrec_e2e_text/data_utils/utils.py): `setuplogger`, `para_and_log`, `save_model` (-> `epoch-N.pt` holding
`model_state_dict`, `optimizer`, `rng_state`, `cuda_rng_state`, `scaler_state`), `get_checkpoint`,
`latest_checkpoint`, `report_time_train`, `report_time_eval`, `get_time`, `str2bool`."""
import argparse
import logging
import math
import os
import time

import torch
import torch.distributed as dist


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = str(v).lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def setuplogger(dir_label, log_paras, time_run, mode, rank, behaviors):
    """two loggers: `Log_file` (file + console) and `Log_screen` (console); only rank 0 / -1 emits INFO"""
    fmt = logging.Formatter("[%(levelname)s %(asctime)s] %(message)s")
    log_file, log_screen = logging.getLogger('Log_file'), logging.getLogger('Log_screen')
    for lg in (log_file, log_screen):
        for h in list(lg.handlers):
            lg.removeHandler(h)
    if rank not in (-1, 0):
        log_file.setLevel(logging.WARN)
        log_screen.setLevel(logging.WARN)
        return log_file, log_screen
    if 'train' in mode:
        log_dir = './logs_' + dir_label + '_train'
        os.makedirs(log_dir, exist_ok=True)
        name = os.path.join(log_dir, 'log_' + log_paras + time_run + '.log')
    else:
        name = ('log_test_all_' if 'test' in mode else 'log_other_') + behaviors.split('_')[0] + '.log'
    fh = logging.FileHandler(filename=name, encoding='utf-8')
    sh = logging.StreamHandler()
    for h in (fh, sh):
        h.setLevel(logging.INFO)
        h.setFormatter(fmt)
    log_file.setLevel(logging.INFO)
    log_screen.setLevel(logging.INFO)
    log_file.addHandler(fh)
    log_file.addHandler(sh)
    log_screen.addHandler(sh)
    return log_file, log_screen


def get_time(start_time, end_time):
    t = int(end_time - start_time)
    return t // 3600, (t // 60) % 60, t % 60


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def para_and_log(model, seq_num, batch_size, Log_file, logging_num, testing_num):
    total = sum(p.numel() for p in model.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    Log_file.info("##### total_num {} #####".format(total))
    Log_file.info("##### trainable_num {} #####".format(trainable))
    step_num = math.ceil(seq_num / _world() / batch_size)
    Log_file.info("##### all {} steps #####".format(step_num))
    steps_for_log = max(int(step_num / logging_num), 1)
    steps_for_test = max(int(step_num / testing_num), 1)
    Log_file.info("##### {} logs/epoch; {} steps/log #####".format(logging_num, steps_for_log))
    Log_file.info("##### {} tests/epoch; {} steps/test #####".format(testing_num, steps_for_test))
    return steps_for_log, steps_for_test


def save_model(now_epoch, model, model_dir, optimizer, rng_state, cuda_rng_state, scaler, Log_file):
    """reference layout (data_utils/utils.py:107-114); `model` is the DDP wrapper (or anything with .module)"""
    ckpt_path = os.path.join(model_dir, f'epoch-{now_epoch}.pt')
    module = model.module if hasattr(model, "module") else model
    torch.save({'model_state_dict': module.state_dict(),
                'optimizer': optimizer.state_dict(),
                'rng_state': rng_state,
                'cuda_rng_state': cuda_rng_state,
                'scaler_state': scaler.state_dict()}, ckpt_path)
    Log_file.info(f"Model saved to {ckpt_path}")
    return ckpt_path


def get_checkpoint(directory, ckpt_name):
    path = os.path.join(directory, ckpt_name)
    return path if os.path.exists(path) else None


def latest_checkpoint(directory, Log_file):
    if not os.path.isdir(directory):
        return None
    names = [n for n in os.listdir(directory) if n.startswith("epoch-") and n.endswith(".pt")]
    Log_file.info(f"[{names}]")
    if not names:
        return None
    best = max(names, key=lambda n: int(n[len("epoch-"):-len(".pt")]))
    return os.path.join(directory, best)


def report_time_train(batch_index, now_epoch, loss, next_set_start_time, start_time, Log_file):
    loss = loss / max(batch_index, 1)
    Log_file.info('epoch: {} end, train_loss: {:.5f}'.format(now_epoch, float(loss)))
    now = time.time()
    Log_file.info("##### (time) this epoch set: {} hours {} minutes {} seconds #####".format(*get_time(next_set_start_time, now)))
    Log_file.info("##### (time) start until now: {} hours {} minutes {} seconds #####".format(*get_time(start_time, now)))
    return time.time()


def report_time_eval(start_time, Log_file):
    Log_file.info("##### (time) eval(valid and test): {} hours {} minutes {} seconds #####".format(*get_time(start_time, time.time())))
