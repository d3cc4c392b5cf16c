"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads without a GPU, exports every symbol that
include/morec_b200.h declares, and the product path fails loudly (no CPU fallback) when asked to compute without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from idvs.morec_b200 import lib as L
    return L


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "morec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(morec_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    h = lib.load()
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/morec_b200.h but not exported"
    assert h.morec_abi_version() == 2


def test_adam_chunk_struct_layout(lib):
    # MorecAdamTensor: 5 pointers + int + 5 floats = 64 bytes (8-byte aligned)
    assert ctypes.sizeof(lib.AdamChunk) == 64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    h = lib.load()
    assert h.morec_device_sms() < 0
    with pytest.raises(Exception):
        x = torch.zeros(4, 4)
        lib.gemm(x, x, x, M=4, N=4, K=4, lda=4, ldb=4, ldc=4)   # device tensors only: asserts / raises on CPU


def test_model_module_surface():
    """same class names / ctor signature / state-dict keys as the reference package"""
    import inspect
    import types
    from idvs.morec_b200.model import Model
    from idvs.morec_b200.model import encoders, modules
    assert list(inspect.signature(Model.__init__).parameters)[1:] == ["args", "item_num", "use_modal", "bert_model", "pop_prob_list"]
    fwd = inspect.signature(Model.forward).parameters
    assert list(fwd)[1:5] == ["sample_items_id", "sample_items", "log_mask", "local_rank"]          # the reference's call
    assert all(p.default is not inspect.Parameter.empty for p in list(fwd.values())[5:])        # extras are optional
    for name in ("User_Encoder", "Text_Encoder", "Bert_Encoder"):
        assert hasattr(encoders, name)
    for name in ("TransformerEncoder", "TransformerBlock", "MultiHeadedAttention", "PositionwiseFeedForward"):
        assert hasattr(modules, name)
    a = types.SimpleNamespace(max_seq_len=5, embedding_dim=16, num_attention_heads=2, drop_rate=0.1, transformer_block=2,
                              num_words_title=4, num_words_abstract=0, num_words_body=0, news_attributes=["title"],
                              bert_model_load="x", word_embedding_dim=8)
    m = Model(a, 10, False, None, [1.0] + [0.1] * 10)
    keys = set(m.state_dict().keys())
    pre = "user_encoder.transformer_encoder."
    want = {"id_embedding.weight", pre + "position_embedding.weight", pre + "layer_norm.weight", pre + "layer_norm.bias"}
    for b in range(2):
        q = pre + f"transformer_blocks.{b}."
        want |= {q + f"multi_head_attention.{n}.weight" for n in ("w_Q", "w_K", "w_V", "fc")}
        want |= {q + "multi_head_attention.layer_norm.weight", q + "multi_head_attention.layer_norm.bias",
                 q + "feed_forward.w_1.weight", q + "feed_forward.w_1.bias", q + "feed_forward.w_2.weight",
                 q + "feed_forward.w_2.bias", q + "feed_forward.layer_norm.weight", q + "feed_forward.layer_norm.bias"}
    assert keys == want


def test_state_dict_keys_match_reference_goldens(goldens):
    """the text-tower model exposes exactly the reference's state-dict keys (golden state dicts come from the reference)"""
    import types
    from transformers import BertConfig, BertModel
    from idvs.morec_b200.model import Model
    g = goldens["text_tiny"]
    m = g["meta"]
    a = types.SimpleNamespace(max_seq_len=m["L"], embedding_dim=m["D"], num_attention_heads=m["heads"], drop_rate=0.1,
                              transformer_block=m["blocks"], num_words_title=m["T"], num_words_abstract=50, num_words_body=50,
                              news_attributes=["title"], bert_model_load="bert_tiny", word_embedding_dim=128)
    model = Model(a, m["N"], True, BertModel(BertConfig(**m["bert_cfg"])), g["pop_prob"].numpy())
    assert set(model.state_dict().keys()) == set(g["state_dict"].keys())


def test_seeded_construction_matches_reference_weights(goldens):
    """same construction order + initialisers as the reference => same weights for the same seed (ID tower golden)"""
    import random
    import types
    import numpy as np
    from idvs.morec_b200.model import Model
    g = goldens["id_small_collide"]
    m = g["meta"]
    torch.manual_seed(m["seed"]); np.random.seed(m["seed"]); random.seed(m["seed"])
    a = types.SimpleNamespace(max_seq_len=m["L"], embedding_dim=m["D"], num_attention_heads=m["heads"], drop_rate=0.1,
                              transformer_block=m["blocks"], num_words_title=0, num_words_abstract=50, num_words_body=50,
                              news_attributes=["title"], bert_model_load="bert_tiny", word_embedding_dim=128)
    model = Model(a, m["N"], False, None, g["pop_prob"].numpy())
    for k, v in g["state_dict"].items():
        assert torch.equal(model.state_dict()[k], v), k
