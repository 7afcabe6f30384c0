"""Small host-side helpers."""
import torch


def to_channels_last(model):
    """Put every 4-D parameter (2D conv weights) in channels_last memory format, in place.  Unlike
    `module.to(memory_format=torch.channels_last)` this leaves the 5-D MVF tap tensors alone (torch rejects
    rank-5 tensors for that format).  Activations follow the input's format: feed channels_last frames."""
    for p in model.parameters():
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    return model
