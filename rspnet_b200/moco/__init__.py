"""``ModelFactory`` with the interface of the reference's ``moco/__init__.py`` (:14-55)."""
import logging

import torch

from ..models import get_model_class
from .builder_diffspeed_diffloss import Loss, MoCoDiffLoss, MoCoDiffLossTwoFc, concat_all_gather
from .ddp import FlatDDP
from .split_wrapper import MultiTaskWrapper

logger = logging.getLogger(__name__)

__all__ = ["ModelFactory", "MoCoDiffLossTwoFc", "MoCoDiffLoss", "MultiTaskWrapper", "Loss", "concat_all_gather",
           "FlatDDP"]


class ModelFactory:
    def __init__(self, cfg):
        self.cfg = cfg

    def build_moco_diffloss(self):
        cfg = self.cfg
        moco_dim = cfg.get_int('moco.dim')
        moco_t = cfg.get_float('moco.t')
        moco_k = cfg.get_int('moco.k')
        moco_m = cfg.get_float('moco.m')
        moco_fc_type = cfg.get_string('moco.fc_type')
        moco_diff_speed = cfg.get_list('moco.diff_speed')
        base_model_class = get_model_class(**cfg.get_config('model'))

        def model_class(num_classes=128):
            return MultiTaskWrapper(base_model_class, num_classes=num_classes, fc_type=moco_fc_type, finetune=False,
                                    groups=1)

        model = MoCoDiffLossTwoFc(model_class, dim=moco_dim, K=moco_k, m=moco_m, T=moco_t,
                                  diff_speed=moco_diff_speed)
        if not torch.cuda.is_available():
            raise RuntimeError("rspnet_b200: build_moco_diffloss needs a B200 (there is no CPU path)")
        model.cuda()
        return FlatDDP(model)
