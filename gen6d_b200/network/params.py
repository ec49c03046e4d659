"""Parameter containers with the reference's state-dict names and shapes.

The drop-in boundary includes checkpoint compatibility: a Gen6D
`model_best.pth['network_state_dict']` must `load_state_dict` into our classes unchanged.  The
modules below therefore only *hold* parameters under the reference's names -- their torch
`forward` is never called; the compute runs through the C-ABI kernels on packed copies (see
`ops.pack_conv` and each network's `_pack`).  Layer tables are written as compact specs rather than literal module listings.

Reference layouts mirrored: network/pretrain_models.py:86-111 (VGG11-BN 'A' features),
network/detector.py:159-184, network/selector.py:27-111, network/attention.py:28-48,
network/refiner.py:24-52,88-134,153-159.
"""
import torch.nn as nn


def sparse_sequential(length, layers):
    """nn.Sequential of `length` slots; slots absent from `layers` are parameter-free."""
    return nn.Sequential(*[layers[i] if i in layers else nn.Identity() for i in range(length)])


_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}


def conv_slots(dim, table, length=None, kernel=3, padding=1):
    """table: {slot: (cin, cout)} or {slot: (cin, cout, kernel, stride, padding)}."""
    layers = {}
    for slot, t in table.items():
        cin, cout = t[0], t[1]
        k, s, p = (t[2], t[3], t[4]) if len(t) == 5 else (kernel, 1, padding)
        layers[slot] = _CONV[dim](cin, cout, k, s, p)
    return sparse_sequential(length if length is not None else max(table) + 1, layers)


# conv slot -> (cin, cout) of torchvision vgg11_bn.features; BatchNorm2d follows at slot+1
VGG11_CONVS = {0: (3, 64), 4: (64, 128), 8: (128, 256), 11: (256, 256),
               15: (256, 512), 18: (512, 512), 22: (512, 512), 25: (512, 512)}
# conv slots grouped by resolution (1/1, 1/2, 1/4, 1/8, 1/16); slot 25 has BN but no ReLU on the path
VGG11_BLOCKS = ((0,), (4,), (8, 11), (15, 18), (22, 25))


class VGG11BNParams(nn.Module):
    """`features.{i}` parameters of vgg11_bn (29 slots, no classifier, no download)."""

    def __init__(self):
        super().__init__()
        layers = {}
        for slot, (cin, cout) in VGG11_CONVS.items():
            layers[slot] = nn.Conv2d(cin, cout, 3, padding=1)
            layers[slot + 1] = nn.BatchNorm2d(cout)
        self.features = sparse_sequential(29, layers)
        for p in self.parameters():
            p.requires_grad = False


def detector_heads(d=64, n_in=12):
    head = lambda cout: conv_slots(2, {0: (d, d), 2: (d, d), 4: (d, cout)})
    return {
        'score_conv': conv_slots(3, {0: (n_in, d, 1, 1, 0), 2: (d, d, 1, 1, 0)}),
        'score_predict': head(1), 'scale_predict': head(1), 'offset_predict': head(2),
    }


# selector correlation towers: conv slot -> (cin, cout); kernel (1,3,3), padding (0,1,1)
SEL_TOWERS = (
    {1: (512, 64), 4: (64, 64), 7: (64, 128), 10: (128, 128), 13: (128, 256), 16: (256, 256)},
    {1: (512, 128), 4: (128, 128), 7: (128, 256), 10: (256, 256)},
    {1: (512, 256), 4: (256, 256)},
)
# what follows each tower conv: n = InstanceNorm3d, r = ReLU, p = MaxPool3d((1,2,2))
SEL_TOWER_POST = (
    {1: 'nr', 4: 'np', 7: 'nr', 10: 'np', 13: 'nr', 16: ''},
    {1: 'nr', 4: 'np', 7: 'nr', 10: ''},
    {1: 'nr', 4: ''},
)


def selector_tower(level):
    table = {s: (ci, co, (1, 3, 3), 1, (0, 1, 1)) for s, (ci, co) in SEL_TOWERS[level].items()}
    return conv_slots(3, table)


class AttentionParams(nn.Module):
    def __init__(self, dim=512):
        super().__init__()
        for name in ('conv_key', 'conv_query', 'conv_feats', 'conv_merge'):
            setattr(self, name, nn.Conv1d(dim, dim, 1))
        self.norm = nn.Module()
        self.norm.norm = nn.LayerNorm(dim)


def selector_modules(angle_num, dim=512):
    one = lambda table: conv_slots(1, {s: (ci, co, 1, 1, 0) for s, (ci, co) in table.items()})
    return {
        'corr_conv_list': nn.ModuleList([selector_tower(l) for l in range(3)]),
        'corr_feats_conv': conv_slots(3, {0: (768, dim, 1, 1, 0), 3: (dim, dim, 1, 1, 0)}),
        'score_process': conv_slots(2, {0: (dim + 3, dim, 1, 1, 0), 2: (dim, dim, 1, 1, 0)}),
        'atts': nn.ModuleList([AttentionParams(dim) for _ in range(2)]),
        'mlps': nn.ModuleList([one({0: (2 * dim, dim), 3: (dim, dim)}) for _ in range(2)]),
        'score_predict': one({0: (dim, dim), 2: (dim, 1)}),
        'angle_predict': one({0: ((dim + 3) * angle_num, dim), 2: (dim, dim), 4: (dim, 1)}),
        'view_point_encoder': sparse_sequential(5, {0: nn.Linear(3, 128), 2: nn.Linear(128, 256),
                                                    4: nn.Linear(256, dim)}),
    }


class RefineFeatureParams(nn.Module):
    """feature_net.{conv0,conv1,conv2,conv_out}.{0,3} + feature_net.backbone.features.*"""
    BRANCHES = {'conv0': (256, 64, 64), 'conv1': (512, 256, 64), 'conv2': (512, 256, 64),
                'conv_out': (192, 128, 128)}

    def __init__(self):
        super().__init__()
        for name, (cin, mid, cout) in self.BRANCHES.items():
            setattr(self, name, conv_slots(2, {0: (cin, mid), 3: (mid, cout)}, length=5))
        self.backbone = VGG11BNParams()


class RefineVolumeParams(nn.Module):
    """volume_net.{mean_embed,var_embed}.{0,3}, volume_net.conv{0..4}.0, volume_net.conv5.{0,3}"""
    # name -> (cin, cout, stride)
    TRUNK = (('conv0', 128, 64, 1), ('conv1', 64, 128, 2), ('conv2', 128, 128, 1),
             ('conv3', 128, 256, 2), ('conv4', 256, 256, 1), ('conv5', 256, 512, 2))

    def __init__(self):
        super().__init__()
        self.mean_embed = conv_slots(3, {0: (256, 64), 3: (64, 64)})
        self.var_embed = conv_slots(3, {0: (128, 64), 3: (64, 64)})
        for name, cin, cout, stride in self.TRUNK:
            table = {0: (cin, cout, 3, stride, 1)}
            if name == 'conv5':
                table[3] = (cout, cout, 3, 1, 1)
            setattr(self, name, conv_slots(3, table, length=max(3, max(table) + 1)))


class RefineRegressorParams(nn.Module):
    def __init__(self, in_feats=512 * 4 ** 3):
        super().__init__()
        self.fc = nn.Sequential(sparse_sequential(2, {0: nn.Linear(in_feats, 512)}),
                                sparse_sequential(2, {0: nn.Linear(512, 512)}))
        self.fcr = nn.Linear(512, 4)
        self.fct = nn.Linear(512, 2)
        self.fcs = nn.Linear(512, 1)
