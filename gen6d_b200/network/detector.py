import torch.nn as nn
from .params import VGG11BNParams, detector_heads

class Detector(nn.Module):
    default_cfg = {
        'vgg_score_stats': [[36.264317, 13.151907], [13910.291, 5345.965], [829.70807, 387.98788]],
        'vgg_score_max': 10,
        'detection_scales': [-1.0, -0.5, 0.0, 0.5],
        'train_feats': False,
    }
    def __init__(self, cfg):
        self.cfg = {**self.default_cfg, **cfg}
        super().__init__()
        self.backbone = VGG11BNParams()
        for k, m in detector_heads(64, 3 * len(self.cfg['detection_scales'])).items():
            setattr(self, k, m)
