"""Deterministic synthetic checkpoints.

There is no network in the build/bench environment, so pretrained Gen6D weights cannot be
fetched; BASELINE.json's configs all say "random weights".  `seeded_state_dict` fills a
module's state dict with values that depend only on (seed, tensor name, shape), so the oracle,
the golden-vector generator (which loads them into the unmodified reference) and the GPU path
all see bit-identical parameters.  Initialisation keeps activations O(1) through the ReLU
stacks (He-style fan-in scaling) and gives the eval-mode BatchNorm non-trivial statistics so
that BN folding is actually exercised.
"""
import zlib

import torch


def _gen(seed, name):
    g = torch.Generator(device='cpu')
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def seeded_tensor(seed, name, ref):
    g = _gen(seed, name)
    shape = tuple(ref.shape)
    leaf = name.rsplit('.', 1)[-1]
    if leaf == 'num_batches_tracked':
        return torch.zeros(shape, dtype=ref.dtype)
    if leaf == 'running_mean':
        return torch.randn(shape, generator=g) * 0.1
    if leaf == 'running_var':
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    if leaf == 'bias':
        b = (torch.rand(shape, generator=g) - 0.5) * 0.1
        if name in _OUTPUT_BIAS:
            b = b * 0.2 + torch.tensor(_OUTPUT_BIAS[name])
        return b
    if leaf == 'weight' and len(shape) == 1:  # BatchNorm / LayerNorm scale
        return torch.rand(shape, generator=g) * 0.4 + 0.8
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    w = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return w * _OUTPUT_GAIN.get(name, 1.0)


# Output heads that regress geometric quantities are given small gains so that the synthetic
# checkpoints behave like a (weakly) trained model: the refiner predicts near-identity pose
# updates (quaternion ~ (1,0,0,0), sub-pixel offsets, log2-scale ~ 0) and the detector predicts
# scale factors near 1.  Without this, random heads throw every pose far off the object and the
# refinement chain becomes an ill-conditioned (chaotic) map on featureless background crops.
_OUTPUT_GAIN = {
    'regressor.fcr.weight': 0.02, 'regressor.fct.weight': 0.05, 'regressor.fcs.weight': 0.02,
    'scale_predict.4.weight': 0.03, 'offset_predict.4.weight': 0.03,
}
_OUTPUT_BIAS = {'regressor.fcr.bias': (1.0, 0.0, 0.0, 0.0)}


def seeded_state_dict(module_or_spec, seed=0):
    """module_or_spec: an nn.Module, or a {name: tensor-like with .shape/.dtype} mapping."""
    spec = module_or_spec.state_dict() if hasattr(module_or_spec, 'state_dict') else module_or_spec
    return {k: seeded_tensor(seed, k, v) for k, v in spec.items()}
