"""Golden fixture for the VISION package: runs the UNMODIFIED reference inbatch_sasrec_e2e_vision/model on CPU fp32
(authoring container only) with a tiny Swin configuration (56x56 images, 2 stages incl. a shifted-window block and a
patch merge).  Separate script because both reference packages are called `model`."""
import os
import sys
import types
import random

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/inbatch_sasrec_e2e_vision"
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from idvs.morec_b200.synth import synth_batch  # noqa: E402

SWIN_CFG = dict(image_size=56, patch_size=4, num_channels=3, embed_dim=32, depths=[2, 2], num_heads=[2, 4], window_size=7,
                mlp_ratio=4.0, qkv_bias=True, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                drop_path_rate=0.1, hidden_act="gelu", layer_norm_eps=1e-5, num_labels=16)


def main():
    assert os.path.isdir(REF)
    sys.path.insert(0, REF)
    from model import Model                      # the reference's own vision Model
    from transformers import SwinConfig, SwinForImageClassification
    seed, B, L, N, D = 31, 3, 4, 20, 32
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    data = synth_batch(B, L, N, 0, seed, modal=False, n_users_pop=50)
    net = SwinForImageClassification(SwinConfig(**SWIN_CFG))
    net.classifier = torch.nn.Linear(net.classifier.in_features, D)            # run.py:49-54
    torch.nn.init.xavier_normal_(net.classifier.weight.data)
    torch.nn.init.constant_(net.classifier.bias.data, 0)
    a = types.SimpleNamespace(max_seq_len=L, embedding_dim=D, num_attention_heads=2, drop_rate=0.1, transformer_block=2,
                              CV_model_load="swin_tiny")
    content = torch.randn(N + 1, 3, 56, 56)
    content[0] = 0                                                              # pad item = all-zero image
    ids = data["ids"]
    images = content[ids.reshape(-1)]
    model = Model(a, N, True, net, data["pop_prob"].numpy()).eval()
    cap = {}
    h1 = model.cv_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("score_embs", o.detach().clone()))
    h2 = model.user_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("prec_vec", o.detach().clone()))
    loss = model(ids.reshape(-1), images, data["log_mask"], "cpu")
    loss.backward()
    h1.remove(); h2.remove()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    out = dict(meta=dict(name="vision_tiny", B=B, L=L, N=N, D=D, heads=2, blocks=2, seed=seed, swin_cfg=SWIN_CFG,
                         torch=torch.__version__, transformers=__import__("transformers").__version__,
                         reference_commit="ce372cf"),
               ids=ids, images=images, log_mask=data["log_mask"], pop_prob=data["pop_prob"],
               state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
               loss=loss.detach().clone(), score_embs=cap["score_embs"], prec_vec=cap["prec_vec"], grads=grads)
    path = os.path.join(HERE, "vision_tiny.pt")
    torch.save(out, path)
    print(f"vision_tiny: loss={float(loss):.6f} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
