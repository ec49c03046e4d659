import torch.nn as nn
from .params import RefineFeatureParams, RefineVolumeParams, RefineRegressorParams

class VolumeRefiner(nn.Module):
    default_cfg = {'refiner_sample_num': 32}
    def __init__(self, cfg):
        self.cfg = {**self.default_cfg, **cfg}
        super().__init__()
        self.feature_net = RefineFeatureParams()
        self.volume_net = RefineVolumeParams()
        self.regressor = RefineRegressorParams()
