"""Host-side odds and ends that need no GPU: the bench's workload table, the launch-list aggregator the profiles are made
with, and the global-mode plan when a rank ends up without items."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_bench_workload_table():
    import bench
    a = bench.make_args(dict(bench.CFG_TINY, bert=bench.BERT_TINY))
    assert a.num_words_title == 128 and a.word_embedding_dim == 128 and a.bert_model_load == "bert_tiny"
    a = bench.make_args(dict(bench.CFG, bert=bench.BERT_BASE))
    assert a.num_words_title == 30 and a.word_embedding_dim == 768 and a.max_seq_len == 25 and a.embedding_dim == 512
    assert bench.SWIN_B["depths"] == [2, 2, 18, 2] and bench.SWIN_B["embed_dim"] == 128 and bench.CFG_VISION_B["B"] == 16
    assert bench.TRAFFIC["bytes"] > 0 and os.path.exists(os.path.join(ROOT, bench.TRAFFIC["source"]))


def test_launch_list_aggregator_on_committed_profile():
    csv = os.path.join(ROOT, "profiles", "r02_launches_text_final.csv")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "agg_step.py"), csv], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    head = r.stdout.splitlines()[0]
    assert head.startswith("one step:") and "launches" in head
    n = int(head.split()[2])
    assert 250 < n < 500                                       # a BERT-base step is ~320 launches
    assert "gemm2_kernel" in r.stdout.splitlines()[1]          # the tcgen05 GEMM is the dominant kernel


def test_global_plan_with_fewer_items_than_ranks():
    from idvs.morec_b200 import parallel as par
    ids = np.array([0, 7, 7, 0, 7, 0, 0, 7], dtype=np.int64)    # ONE distinct item, 4 ranks
    plans = [par.plan_global_batch(ids, 4, r) for r in range(4)]
    assert [p.my_first_slots.size for p in plans] == [1, 0, 0, 0] and all(p.u_max == 1 for p in plans)
    # every real slot reads the single encoded row (owner 0, local 0); pad slots read nothing
    assert np.array_equal(plans[2].slot_to_row, np.where(ids != 0, 0, -1))
    # all padding: nobody encodes anything, the gathered table still has one (zero) row per rank
    p = par.plan_global_batch(np.zeros(8, dtype=np.int64), 4, 1)
    assert p.n_unique == 0 and p.u_max == 1 and p.my_first_slots.size == 0


def test_fused_adamw_declares_the_grad_scaler_contract():
    """torch.amp.GradScaler.step hands itself to optimizers whose step() has a `grad_scaler` parameter and then does NOT
    run its own pass over the gradients (torch/amp/grad_scaler.py): the signature is the contract"""
    import inspect
    from idvs.morec_b200.optim import FusedAdamW
    params = inspect.signature(FusedAdamW.step).parameters
    assert "grad_scaler" in params and params["grad_scaler"].default is None
    assert getattr(FusedAdamW, "_step_supports_amp_scaling", False) is True


def test_scoring_arithmetic_follows_the_precision_mode():
    """Model._ce_inputs: 3xTF32 on fp32 operands in the parity mode, one TF32 pass in tf32, the 16-bit activations as
    they are in fp16 / bf16 (what autocast does with the logits matmul, model/model.py:49)"""
    import types
    import torch
    from idvs.morec_b200.model import Model
    a = types.SimpleNamespace(max_seq_len=4, embedding_dim=8, num_attention_heads=2, drop_rate=0.0, transformer_block=1)
    m = Model(a, 10, False, None, np.ones(11) / 11)
    P, E = torch.randn(8, 8), torch.randn(10, 8)
    for mode, dt, x3, out_dt in (("fp32", torch.float32, True, torch.float32), ("tf32", torch.float32, False, torch.float32),
                                 ("fp16", torch.float16, False, torch.float16), ("bf16", torch.bfloat16, False, torch.bfloat16)):
        m.compute_dtype = mode
        meta, p, e = m._ce_inputs(P.to(dt), E.to(dt))
        assert meta["x3"] is x3 and p.dtype == out_dt and e.dtype == out_dt, mode
    m.compute_dtype = "fp16"                       # mixed inputs fall back to fp32 operands
    meta, p, e = m._ce_inputs(P, E.half())
    assert p.dtype == torch.float32 and e.dtype == torch.float32
