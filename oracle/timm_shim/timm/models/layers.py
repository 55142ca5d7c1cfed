"""Import-time-only names (reference SOccDPT/model/dpt.py:3; only called for LeViT)."""
import torch.nn as nn


def get_act_layer(name="relu"):
    table = {"relu": nn.ReLU, "gelu": nn.GELU, "hard_swish": nn.Hardswish,
             "silu": nn.SiLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh}
    return table[name]


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
