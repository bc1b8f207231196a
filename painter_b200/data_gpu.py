"""Per-sample work of the training data path on the GPU: what `PairDataset.__getitem__`
(Painter/data/pairdataset.py:106-190) does after the images are decoded and augmented, without the host.

  DeviceMaskingGenerator   util/masking_generator.py:15-93 + the half-mask draw of pairdataset.py:149,183-186
  valid_maps               pairdataset.py:154-181 (per-task `valid` weighting of the loss)
  combine_pairs            pairdataset.py:100-104,136-146 (second pair stitched under the first)

Decoding and the PIL augmentations stay on the host (file I/O is out of scope)."""
import ctypes
import math

import torch

from ._lib import check, lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceMaskingGenerator:
    """Same constructor as the reference MaskingGenerator; `gen(B, seed)` returns int32 [B, h, w] on the device."""

    def __init__(self, input_size, num_masking_patches, min_num_patches=4, max_num_patches=None, min_aspect=0.3,
                 max_aspect=None, half_mask_ratio=0.0, device="cuda"):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.num_masking_patches = num_masking_patches
        self.min_num_patches = min_num_patches
        self.max_num_patches = num_masking_patches if max_num_patches is None else max_num_patches
        max_aspect = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(max_aspect))
        self.half_mask_ratio = half_mask_ratio
        self.device = torch.device(device)
        self._calls = 0

    def get_shape(self):
        return self.height, self.width

    def __call__(self, B=1, seed=None):
        if seed is None:
            seed = 0x5EED0000 + self._calls
        self._calls += 1
        out = torch.empty((B, self.height, self.width), dtype=torch.int32, device=self.device)
        check(lib().pk_block_masks(ctypes.c_void_p(out.data_ptr()), B, self.height, self.width,
                                   self.num_masking_patches, self.min_num_patches, self.max_num_patches,
                                   ctypes.c_float(self.log_aspect_ratio[0]), ctypes.c_float(self.log_aspect_ratio[1]),
                                   ctypes.c_float(self.half_mask_ratio), ctypes.c_ulonglong(seed), _stream()),
              "pk_block_masks")
        return out


def _rule_and_threshold(pair_type):
    """pairdataset.py:157-181 -> (rule code, threshold before normalisation)."""
    if "nyuv2_image2depth" in pair_type:
        return 1, 1e-3 * 0.1
    if "ade20k_image2semantic" in pair_type or "coco_image2panoptic_sem_seg" in pair_type:
        return 1, 1e-5
    if "image2pose" in pair_type:
        return 2, 1e-5
    if "image2panoptic_inst" in pair_type:
        return 3, 1e-5
    return 0, 0.0


def valid_maps(targets, pair_types):
    """targets: fp32 [B, 3, H, W] on the device (normalised); pair_types: list of B type strings -> valid [B,3,H,W]."""
    assert targets.is_cuda and targets.dtype == torch.float32 and targets.is_contiguous()
    B, _, H, W = targets.shape
    rules, thr = [], []
    mean = torch.tensor(IMAGENET_MEAN)
    std = torch.tensor(IMAGENET_STD)
    for t in pair_types:
        r, th = _rule_and_threshold(t)
        rules.append(r)
        thr.append((torch.ones(3) * th - mean) / std)      # the reference's fp32 torch arithmetic, bit for bit
    rule_d = torch.tensor(rules, dtype=torch.int32).to(targets.device)
    thr_d = torch.stack(thr).to(targets.device).contiguous()
    fg = torch.zeros(B, dtype=torch.int32, device=targets.device)
    valid = torch.empty_like(targets)
    check(lib().pk_valid_maps(ctypes.c_void_p(targets.data_ptr()), ctypes.c_void_p(rule_d.data_ptr()),
                              ctypes.c_void_p(thr_d.data_ptr()), ctypes.c_void_p(fg.data_ptr()),
                              ctypes.c_void_p(valid.data_ptr()), B, H, W, _stream()), "pk_valid_maps")
    return valid


def combine_pairs(first, second):
    """pairdataset.py:100-104: the second pair goes under the first ([.., 3, H, W] x 2 -> [.., 3, 2H, W])."""
    return torch.cat([first, second], dim=-2)
