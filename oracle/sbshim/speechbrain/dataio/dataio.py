"""speechbrain.dataio.dataio stand-in: only ``length_to_mask`` (used at TransformerASR.py:160-162)."""
import torch


def length_to_mask(length, max_len=None, dtype=None, device=None):
    assert len(length.shape) == 1
    if max_len is None:
        max_len = length.max().long().item()
    mask = torch.arange(max_len, device=length.device, dtype=length.dtype).expand(
        len(length), max_len
    ) < length.unsqueeze(1)
    if dtype is None:
        dtype = length.dtype
    if device is None:
        device = length.device
    return torch.as_tensor(mask, dtype=dtype, device=device)
