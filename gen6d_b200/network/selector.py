import torch.nn as nn
from .params import VGG11BNParams, selector_modules

class ViewpointSelector(nn.Module):
    default_cfg = {'selector_angle_num': 5}
    def __init__(self, cfg):
        self.cfg = {**self.default_cfg, **cfg}
        super().__init__()
        self.backbone = VGG11BNParams()
        for k, m in selector_modules(self.cfg['selector_angle_num']).items():
            setattr(self, k, m)
