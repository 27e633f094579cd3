"""Weight initialisers with the reference's names and distributions
(/root/reference/torchrl/networks/init.py:5-47).

An initialiser is a `LayerInit(weight_rule, bias_rule)`; a rule fills one tensor in place.  The random rules
consume the global torch generator exactly like the reference's (one `uniform_` / `orthogonal_` call per tensor,
weight before bias), so equal seeds give bit-identical networks (tests/test_oracle_vs_reference.py).
"""
import math

import torch.nn as nn


def fanin_uniform(alpha=0.0):
    """U(-b, b) with b = sqrt(1 / ((1 + alpha^2) * fan_in)).  fan_in follows the reference's convention
    (init.py:5-15): size(0) -- the OUTPUT width -- for 2-D weights, prod(size[1:]) for conv kernels."""
    def rule(tensor):
        if tensor.dim() < 2:
            raise ValueError("fan-in initialisation needs a tensor with at least 2 dims")
        fan_in = tensor.size(0) if tensor.dim() == 2 else math.prod(tensor.shape[1:])
        bound = math.sqrt(1.0 / ((1.0 + alpha * alpha) * fan_in))
        tensor.data.uniform_(-bound, bound)
    return rule


def uniform(bound=3e-3):
    def rule(tensor):
        tensor.data.uniform_(-bound, bound)
    return rule


def fill(value=0.1):
    def rule(tensor):
        tensor.data.fill_(value)
    return rule


def orthogonal(gain=math.sqrt(2)):
    def rule(tensor):
        nn.init.orthogonal_(tensor, gain=gain)
    return rule


class LayerInit:
    def __init__(self, weight_rule, bias_rule):
        self.weight_rule, self.bias_rule = weight_rule, bias_rule

    def __call__(self, layer):
        self.weight_rule(layer.weight)
        self.bias_rule(layer.bias)


def layer_init(layer, weight_init=fanin_uniform(), bias_init=fill(0.1)):
    LayerInit(weight_init, bias_init)(layer)


basic_init = LayerInit(fanin_uniform(), fill(0.1))       # hidden layers (init.py:33-34)
uniform_init = LayerInit(uniform(3e-3), uniform(3e-3))       # output layers (init.py:37-38)


def orthogonal_init(layer, scale=math.sqrt(2), constant=0):
    """orthogonal weights with gain `scale`, constant bias (init.py:45-47)."""
    LayerInit(orthogonal(scale), fill(constant))(layer)


# the reference's private rule names, for code that passes them to layer_init
_fanin_init = fanin_uniform()
_uniform_init = uniform(3e-3)
_constant_bias_init = fill(0.1)
