"""Item image encoders, mirroring inbatch_sasrec_e2e_vision/model/encoders.py on the morec_b200 CUDA kernels."""
import torch
import torch.nn as nn

from .. import swin
from ..model.encoders import User_Encoder, _adt, _x3  # noqa: F401  (User_Encoder is identical in both packages)


class Vit_Encoder(torch.nn.Module):                 # reference: encoders.py:24-31
    """GELU(image_net(pixel_values)[0]) for HF SwinForImageClassification whose classifier was replaced by
    nn.Linear(num_features, embedding_dim) (inbatch_sasrec_e2e_vision/run.py:47-54).  `image_net` is kept as a
    sub-module for its parameters (names, order, objects unchanged: run.py:58-60,125-131 address them by index and
    by the substrings 'image_net' / 'classifier'); its forward is never called."""

    def __init__(self, image_net):
        super().__init__()
        self.image_net = image_net
        self.compute_dtype = "fp32"

    def forward(self, item_content):
        cfg = self.image_net.config
        assert getattr(cfg, "hidden_dropout_prob", 0.0) == 0.0 and getattr(cfg, "attention_probs_dropout_prob", 0.0) == 0.0, \
            "Swin hidden/attention dropout are 0 in every reference config (pretrained_models/swin_*/config.json)"
        meta = dict(net=self.image_net, adt=_adt(self), x3=_x3(self), training=self.training)
        x = item_content if item_content.dtype == torch.float32 else item_content.float()
        return swin.SwinTowerFn.apply(meta, x.contiguous(), *list(self.image_net.parameters()))
