"""The step right before the hot path: what the reference's dataset
classes do on the host between the decoded files and the network input.

* `crop_multiple` - xview/datasets/augmentation.py:244-262: spatial size forced to a multiple of 16
  (4 poolings) by cropping at the bottom / right.
* `collate` - xview/datasets/data_baseclass.py:64-79 (`_get_batch`): crop every item, stack per
  modality.  Unlike the reference, image modalities keep their raw integer dtype (uint8 rgb,
  uint16 depth): the float32 cast happens on the device after the host->device copy
  (`device.convert_to_f32`), which cuts the PCIe bytes of a 768x384 RGB-D frame from 4.7 MB to
  1.5 MB; labels become int32 as in the reference.
"""
import numpy as np


def crop_multiple(data, multiple_of=16, batched=False):
    """Crop the two spatial dims (0,1 - or 1,2 for a batched array) to multiples of `multiple_of`."""
    try:
        shape = data.shape
    except AttributeError:
        return data
    a = 1 if batched else 0
    if len(shape) < a + 2:
        return data
    h, w = shape[a], shape[a + 1]
    h_c, w_c = h - (h % multiple_of), w - (w % multiple_of)
    if h_c == h and w_c == w:
        return data
    return data[:, :h_c, :w_c, ...] if batched else data[:h_c, :w_c, ...]


def collate(items, modalities=None, keep_raw_dtype=True):
    """List of per-image dicts -> dict of batched arrays (axis 0 = image)."""
    modalities = list(modalities or items[0].keys())
    batch = {}
    for mod in modalities:
        stacked = np.stack([crop_multiple(np.asarray(item[mod])) for item in items])
        if mod == 'labels':
            stacked = stacked.astype('int32')
        elif not keep_raw_dtype or stacked.dtype.kind == 'f':
            stacked = stacked.astype('float32')
        batch[mod] = stacked
    return batch
