"""Weight initialisation with the reference's distributions (constructor parity).

SubgraphCountingMatching/utils/init.py:17-75,125-193: `init_weight(w, activation, init="uniform")` draws
U(-a, a) with a = sqrt(3) * gain * sqrt(2 / (fan_in + fan_out)), where the gain is
`nn.init.calculate_gain` of the activation family with leaky slope 1/5.5; Linear biases are zeroed.
"""
import math

import torch.nn as nn

from .constants import LEAKY_RELU_A

_FAMILY = {
    "none": "linear", "maximum": "linear", "minimum": "linear",
    "relu": "relu", "relu6": "relu", "elu": "relu", "selu": "relu", "celu": "relu", "gelu": "relu",
    "leaky_relu": "leaky_relu", "prelu": "leaky_relu",
    "softmax": "sigmoid", "sparsemax": "sigmoid", "gumbel_softmax": "sigmoid",
    "sigmoid": "sigmoid", "tanh": "tanh",
}


def calculate_gain(activation):
    if not isinstance(activation, str):
        raise ValueError("activation must be given by name")
    if activation not in _FAMILY:
        raise NotImplementedError(activation)
    return nn.init.calculate_gain(_FAMILY[activation], LEAKY_RELU_A)


def xavier_uniform_init(x, gain=1.0):
    t = x if x.dim() >= 2 else x.unsqueeze(-1)
    fan_in, fan_out = t.size(1), t.size(0)
    if t.dim() > 2:
        rf = t[0][0].numel()
        fan_in, fan_out = fan_in * rf, fan_out * rf
    a = math.sqrt(3.0) * gain * math.sqrt(2.0 / float(fan_in + fan_out))
    return nn.init.uniform_(x, -a, a)


def init_weight(x, activation="none", init="uniform"):
    if init != "uniform":
        raise NotImplementedError("DMPNN layers are built with init='uniform' (dmpnn.py:64-73)")
    return xavier_uniform_init(x, gain=calculate_gain(activation))


def init_module(m, activation="none", init="uniform"):
    if isinstance(m, nn.Linear):
        init_weight(m.weight, activation, init)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, (nn.BatchNorm1d, nn.LayerNorm)):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)
