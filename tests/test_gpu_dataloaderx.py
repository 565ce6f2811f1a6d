"""GPU check of msml_b200.datasets.DataLoaderX (drop-in for ref datasets/dataloaderx.py:40-67)."""

import pytest
import torch
from torch.utils.data import TensorDataset

from gpu_util import need_gpu

pytestmark = pytest.mark.gpu


def test_dataloaderx_delivers_every_batch_on_the_device():
    need_gpu()
    from msml_b200.datasets import DataLoaderX
    torch.manual_seed(0)
    img = torch.randn(44, 3, 16, 16)
    msk = torch.randint(0, 2, (44, 16, 16))
    label = torch.arange(44)
    for cl in (False, True):
        loader = DataLoaderX(local_rank=0, channels_last=cl, dataset=TensorDataset(img, msk, label), batch_size=8, shuffle=False,
                             num_workers=0, pin_memory=True, drop_last=False)
        for epoch in range(2):                                   # the loader is re-iterable, as in ref train.py:240
            seen = 0
            for x, m, y in loader:
                assert x.is_cuda and m.is_cuda and y.is_cuda
                n = y.numel()
                assert torch.equal(y.cpu(), label[seen:seen + n])
                assert torch.equal(x.cpu(), img[seen:seen + n]) and torch.equal(m.cpu(), msk[seen:seen + n])
                assert x.is_contiguous(memory_format=torch.channels_last) == cl
                (x * 2).sum().item()                             # consumed on the current stream
                seen += n
            assert seen == 44
