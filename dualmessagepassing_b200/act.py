"""Activation registry with the reference's names and sharing behaviour.

SubgraphCountingMatching/utils/act.py:457-489 keeps ONE module instance per name and hands the same
instance to every layer (so e.g. a `prelu` weight is shared model-wide); `map_activation_str_to_layer`
returns that shared instance.  Only the elementwise activations a DMPNN layer can be built with are
provided; the reference's sparsemax / gumbel_softmax / maximum / minimum are not on the hot path.
"""
import torch.nn as nn

from .constants import LEAKY_RELU_A

supported_act_funcs = {
    "none": nn.Identity(),
    "softmax": nn.Softmax(dim=-1),
    "sigmoid": nn.Sigmoid(),
    "tanh": nn.Tanh(),
    "relu": nn.ReLU(),
    "relu6": nn.ReLU6(),
    "leaky_relu": nn.LeakyReLU(negative_slope=LEAKY_RELU_A),
    "prelu": nn.PReLU(init=LEAKY_RELU_A),
    "elu": nn.ELU(),
    "celu": nn.CELU(),
    "selu": nn.SELU(),
    "gelu": nn.GELU(),
}


def map_activation_str_to_layer(act_func, **kw):
    if act_func not in supported_act_funcs:
        raise NotImplementedError("activation %r is not available in dualmessagepassing_b200" % (act_func,))
    act = supported_act_funcs[act_func]
    for k, v in kw.items():
        if hasattr(act, k):
            try:
                setattr(act, k, v)
            except Exception:
                pass
    return act
