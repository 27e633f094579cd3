"""Weight initialisers with the reference's names and distributions
(/root/reference/torchrl/networks/init.py:5-47)."""
import math

import torch.nn as nn


def _fanin_uniform_(tensor, alpha=0.0):
    """U(-b, b), b = sqrt(1 / ((1+alpha^2) * fan_in)); fan_in = size(0) for 2-D weights
    (the reference's convention, init.py:5-15), prod(size[1:]) for conv kernels."""
    if tensor.dim() == 2:
        fan_in = tensor.size(0)
    elif tensor.dim() > 2:
        fan_in = 1
        for s in tensor.shape[1:]:
            fan_in *= s
    else:
        raise ValueError("need a tensor with at least 2 dims")
    bound = math.sqrt(1.0 / ((1.0 + alpha * alpha) * fan_in))
    return tensor.data.uniform_(-bound, bound)


def _small_uniform_(tensor, param=3e-3):
    return tensor.data.uniform_(-param, param)


def _const_(tensor, constant=0.1):
    tensor.data.fill_(constant)


def layer_init(layer, weight_init=_fanin_uniform_, bias_init=_const_):
    weight_init(layer.weight)
    bias_init(layer.bias)


def basic_init(layer):
    """fan-in uniform weights, bias 0.1 (init.py:33-34)."""
    layer_init(layer, _fanin_uniform_, _const_)


def uniform_init(layer):
    """U(-3e-3, 3e-3) for weights and bias (init.py:37-38)."""
    layer_init(layer, _small_uniform_, _small_uniform_)


def orthogonal_init(layer, scale=math.sqrt(2), constant=0):
    """orthogonal weights with gain `scale`, zero bias (init.py:45-47)."""
    nn.init.orthogonal_(layer.weight, gain=scale)
    layer.bias.data.fill_(0)
