"""Bit-unpacking of Lytro raw sensor data on the device: the array work of `LfpDecoder.comp_bayer`
(lfp_reader/lfp_decoder.py:273-326).  Container parsing (sections, sha-1, json; lfp_decoder.py:76-271) is file
I/O and stays with the reference's LfpDecoder; a maintainer swaps the body of comp_bayer for

    self._bay_img = comp_bayer(self._img_buf, self._shape, self.cfg.lfpimg.get('bit', 10))
"""
import numpy as np
import torch

from .. import _native as nat
from .. import ops


def comp_bayer_device(img_buf, shape, bit_pac=10):
    """Packed bytes (bytes / bytearray / uint8 array / uint8 CUDA tensor) -> (shape[1], shape[0]) uint16 CUDA tensor."""
    if bit_pac not in (10, 12):
        raise ValueError("bit packing %r (10 or 12, lfp_decoder.py:280-312)" % (bit_pac,))
    dev = ops.require_cuda()
    if hasattr(img_buf, 'is_cuda'):
        buf = img_buf.contiguous()
    else:
        host = np.frombuffer(bytes(img_buf), dtype=np.uint8) if not isinstance(img_buf, np.ndarray) \
            else np.ascontiguousarray(img_buf, dtype=np.uint8)
        pinned = torch.empty(host.shape, dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[...] = host
        buf = pinned.to(dev, non_blocking=True)
    nbytes = int(buf.numel())
    npx = (nbytes // 5) * 4 if bit_pac == 10 else (nbytes // 3) * 2
    rows, cols = int(shape[1]), int(shape[0])
    if npx != rows * cols:
        # np.reshape in the reference raises the same way (lfp_decoder.py:315)
        raise ValueError("cannot reshape array of size %d into shape (%d,%d)" % (npx, rows, cols))
    out = torch.empty((rows, cols), dtype=torch.uint16, device=dev)
    nat.check(nat.lib().pcb_unpack_bayer(ops._ptr(buf), nbytes, int(bit_pac), ops._ptr(out), ops._stream()))
    return out


def comp_bayer(img_buf, shape, bit_pac=10):
    """float64 (shape[1], shape[0]) Bayer image, like `LfpDecoder._bay_img` after comp_bayer (:315-318)."""
    return ops.to_host(comp_bayer_device(img_buf, shape, bit_pac)).astype('float')
