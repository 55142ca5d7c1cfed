"""Import-time-only name (reference SOccDPT/model/backbones/beit.py:8; BEiT is out of scope)."""


def gen_relative_position_index(window_size):
    raise NotImplementedError("timm shim: BEiT is outside the SOccDPT hot path")
