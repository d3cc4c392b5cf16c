"""Command line of run.py -- every flag of the reference's parameters.py (inbatch_sasrec_e2e_text/parameters.py:4-50,
inbatch_sasrec_e2e_vision/parameters.py) with the same names, types and defaults, plus the launcher-compat fixes of
SURVEY.md §8(b):

  * `--news` exists (the text launchers pass it, train_bert_base.py:42, and run.py:79,100 read args.news; the
    reference parser lacks it and argparse silently abbreviates it to --news_attributes);
  * both spellings `--local_rank` (torch < 2.0 launch) and `--local-rank` (torch >= 2.0) are accepted and the
    LOCAL_RANK environment variable (torchrun) is the fallback;
  * abbreviations are disabled so a typo can no longer alias another flag.

New, optional flags (defaults keep the reference's behaviour): --compute_dtype, --parallel_mode, --optimizer,
--device_batches, --eval_batch_size, --max_steps, --synthetic.
"""
import argparse
import os


def _common(p):
    p.add_argument("--mode", type=str, default="train")
    p.add_argument("--item_tower", type=str, default="id")
    p.add_argument("--root_data_dir", type=str, default="../")
    p.add_argument("--batch_size", type=int, default=64)
    p.add_argument("--epoch", type=int, default=1)
    p.add_argument("--fine_tune_lr", type=float, default=1e-5)
    p.add_argument("--l2_weight", type=float, default=0)
    p.add_argument("--fine_tune_l2_weight", type=float, default=0)
    p.add_argument("--drop_rate", type=float, default=0.1)
    p.add_argument("--num_attention_heads", type=int, default=2)
    p.add_argument("--transformer_block", type=int, default=2)
    p.add_argument("--min_seq_len", type=int, default=5)
    p.add_argument("--num_workers", type=int, default=12)
    p.add_argument("--load_ckpt_name", type=str, default='None')
    p.add_argument("--label_screen", type=str, default='None')
    p.add_argument("--logging_num", type=int, default=8)
    p.add_argument("--testing_num", type=int, default=1)
    p.add_argument("--local_rank", "--local-rank", dest="local_rank", default=-1, type=int)
    # ---- morec_b200 additions
    p.add_argument("--compute_dtype", type=str, default="fp16", choices=["fp32", "tf32", "bf16", "fp16"],
                   help="fp16 = the arithmetic of the reference's autocast loop (run.py:242); fp32 = 3xTF32 parity mode")
    p.add_argument("--parallel_mode", type=str, default="local", choices=["local", "global"],
                   help="local = reference DDP semantics (rank-local negatives); global = items encoded once per "
                        "global batch, one all-gather of item embeddings, global negatives")
    p.add_argument("--optimizer", type=str, default="fused", choices=["fused", "torch"],
                   help="fused = idvs.morec_b200.optim.FusedAdamW; torch = torch.optim.AdamW as in the reference")
    p.add_argument("--device_batches", type=int, default=1,
                   help="1: item content / user sequences live on the GPU and batches are assembled there; "
                        "0: DataLoader workers exactly as the reference")
    p.add_argument("--eval_batch_size", type=int, default=512)
    p.add_argument("--max_steps", type=int, default=0, help="stop every epoch after this many steps (0 = full epoch)")
    p.add_argument("--synthetic", type=str, default="None",
                   help="'users,items' : generate a synthetic dataset of that size instead of reading TSV files")


def build_parser(kind="text"):
    p = argparse.ArgumentParser(allow_abbrev=False)
    _common(p)
    if kind == "text":
        p.add_argument("--dataset", type=str, default='MIND-small')
        p.add_argument("--behaviors", type=str, default='behaviors_l5_tr_v.tsv')
        p.add_argument("--news", type=str, default='news.tsv')
        p.add_argument("--lr", type=float, default=1e-5)
        p.add_argument("--bert_model_load", type=str, default='bert-base-uncased')
        p.add_argument("--freeze_paras_before", type=int, default=165)
        p.add_argument("--word_embedding_dim", type=int, default=768)
        p.add_argument("--embedding_dim", type=int, default=256)
        p.add_argument("--max_seq_len", type=int, default=20)
        p.add_argument("--num_words_title", type=int, default=30)
        p.add_argument("--num_words_abstract", type=int, default=50)
        p.add_argument("--num_words_body", type=int, default=50)
        p.add_argument("--news_attributes", type=str, default='title')
    else:
        p.add_argument("--dataset", type=str, default='pinterest')
        p.add_argument("--behaviors", type=str, default='users_log.tsv')
        p.add_argument("--images", type=str, default='images_log.tsv')
        p.add_argument("--lmdb_data", type=str, default='image.lmdb')
        p.add_argument("--cold_seqs", type=str, default='None')
        p.add_argument("--new_seqs", type=str, default='None')
        p.add_argument("--new_items", type=str, default='None')
        p.add_argument("--new_lmdb_data", type=str, default='None')
        p.add_argument("--lr", type=float, default=1e-3)
        p.add_argument("--accumulation_step", type=int, default=1)
        p.add_argument("--CV_model_load", type=str, default='resnet-50')
        p.add_argument("--freeze_paras_before", type=int, default=45)
        p.add_argument("--CV_resize", type=int, default=224)
        p.add_argument("--embedding_dim", type=int, default=64)
        p.add_argument("--max_seq_len", type=int, default=10)
    return p


def parse_args(argv=None, kind="text"):
    args = build_parser(kind).parse_args(argv)
    if kind == "text":
        args.news_attributes = args.news_attributes.split(',')
    if args.local_rank is None or args.local_rank < 0:
        args.local_rank = int(os.environ.get("LOCAL_RANK", args.local_rank if args.local_rank is not None else -1))
    return args
