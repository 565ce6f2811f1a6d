"""Test-time occluders — drop-in for the two transforms of ref datasets/augment/rand_occ.py that the verification path
uses (BASELINE config 5): ``RandomBlock`` (:25-72, the 0-90 % square occlusion of ref eval/qeval_mxnet.py:528-547)
and ``NoneOcc`` (:78-87).  Host code on PIL / numpy, as in the reference: it runs in the data pipeline before the images
reach the GPU.  The random draws come from numpy's global generator in the reference's order (ratio, [3 Gaussian planes],
x, y), so ``np.random.seed(s)`` reproduces the reference's occluded set pixel for pixel
(tests/test_oracle_golden.py::test_random_block_matches_reference).

``random_block_batch`` applies the same transform to an (N, 3, H, W) uint8 array without the PIL round trip (same draws,
same pixels): that is what bench.py's config-5 workload and msml_b200.eval use to build a 12,000-image set quickly.
The training-time occluders of the reference (rectangles, ellipses, glasses, scarves, objects, 3-D masks; :89-600) need its
PNG assets and are out of scope.
"""
import copy

import numpy as np
from PIL import Image

__all__ = ["RandomBlock", "NoneOcc", "random_block_batch"]


class RandomBlock(object):
    """lo, hi: occluded AREA in percent, drawn from [lo, hi); fill: 'black' | 'white' | 'gauss'."""
    fill_list = ['black', 'white', 'gauss', ]

    def __init__(self, lo: int, hi: int, fill: str = 'black'):
        self.lo = lo
        self.hi = hi
        self.fill = fill
        assert fill in RandomBlock.fill_list

    def __call__(self, img):
        ratio = np.random.randint(self.lo, self.hi) * 0.01
        return self._block_occ(img, ratio)

    def _block_occ(self, img, ratio):
        width = img.size[0]
        img_occ = copy.deepcopy(img)
        if ratio == 0:
            return img_occ
        block_width = int((ratio * width * width) ** 0.5)
        occ = Image.fromarray(_block_pixels(block_width, self.fill, img.mode))
        randx = np.random.randint(0, width - block_width + 1)
        randy = np.random.randint(0, width - block_width + 1)
        img_occ.paste(occ, (randx, randy))
        return img_occ


def _block_pixels(block_width, fill, mode):
    """The occluder patch, consuming numpy's generator exactly as ref :52-66 does."""
    if fill == 'black':
        return np.zeros([block_width, block_width], dtype=np.uint8)
    if fill == 'white':
        return np.ones([block_width, block_width], dtype=np.uint8) * 255
    if mode == 'L':
        return np.random.randn(block_width, block_width) * 255
    if mode == 'RGB':
        occ_r = np.random.randn(block_width, block_width)
        occ_g = np.random.randn(block_width, block_width)
        occ_b = np.random.randn(block_width, block_width)
        return (np.stack((occ_r, occ_g, occ_b), axis=2) * 255).astype(np.uint8)
    raise ValueError('Error Image type.')


class NoneOcc(object):
    """No occlusion; returns (img, all-white mask) (ref :78-87)."""

    def __init__(self, ret_msk: bool = True):
        self.ret_msk = ret_msk

    def __call__(self, img):
        width, height = img.size[0], img.size[1]
        assert width == height
        return img, Image.fromarray(np.ones((height, width), dtype=np.uint8) * 255)


def random_block_batch(images, lo, hi, fill='black'):
    """images (N, 3, H, W) uint8 numpy (RGB planes) -> occluded copy; image i gets exactly what
    ``RandomBlock(lo, hi, fill)(PIL image i)`` would produce when the transforms are applied in order."""
    images = np.asarray(images)
    if images.ndim != 4 or images.shape[1] != 3 or images.dtype != np.uint8:
        raise ValueError("random_block_batch expects an (N, 3, H, W) uint8 array")
    assert fill in RandomBlock.fill_list
    out = images.copy()
    width = images.shape[3]
    for i in range(images.shape[0]):
        ratio = np.random.randint(lo, hi) * 0.01
        if ratio == 0:
            continue
        bw = int((ratio * width * width) ** 0.5)
        patch = _block_pixels(bw, fill, 'RGB')
        x = np.random.randint(0, width - bw + 1)
        y = np.random.randint(0, width - bw + 1)
        if patch.ndim == 2:                        # a gray patch pasted into an RGB image: PIL replicates it over the planes
            out[i, :, y:y + bw, x:x + bw] = patch[None]
        else:
            out[i, :, y:y + bw, x:x + bw] = patch.transpose(2, 0, 1)
    return out
