"""CPU tests of the host-side logic that needs no GPU: gradient-buffer ordering for the overlapped all-reduce, dense-layout
detection, channel padding helper, stale-gradient hygiene."""
import pytest
import torch

from msml_b200 import ops


class _FakePFC:
    sample_rate = 1.0


def test_flat_buffer_is_ordered_by_backward_completion():
    from msml_b200.backbones import MSML
    from msml_b200.engine import TrainStep
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), 10, header_type=None, fm_params=(3, 2, "sigmoid", "mul"))
    step = TrainStep(net, _FakePFC(), None, None, (2, 3, 112, 112), use_graph=False)
    names = {id(p): n for n, p in net.named_parameters()}
    used = [p for n, p in net.named_parameters() if n.startswith("frb.") and p.requires_grad]
    ordered, buckets = step._order_for_overlap(used)
    assert sorted(id(p) for p in ordered) == sorted(id(p) for p in used)          # a permutation: nothing lost or duplicated
    assert [t for t, _ in buckets] == [3, 2, 1, 0, -1] and sum(c for _, c in buckets) == len(used)
    pos = 0
    for tag, cnt in buckets:
        for p in ordered[pos:pos + cnt]:
            n = names[id(p)]
            if tag >= 0:
                assert (n.startswith("frb.layer%d." % (tag + 1)) or n.startswith("frb.fm_ops.%d." % tag)
                        or (tag == 3 and n.startswith(("frb.bn2.", "frb.fc.", "frb.features.")))), (tag, n)
            else:
                assert n.startswith(("frb.conv1.", "frb.bn1.", "frb.prelu.")), n
        pos += cnt
    # stage 3 (finished first in backward) sits at the front, the stem at the very end
    assert names[id(ordered[0])].startswith(("frb.layer4.", "frb.fm_ops.3.", "frb.bn2.", "frb.fc.", "frb.features."))
    assert names[id(ordered[-1])].startswith(("frb.conv1.", "frb.bn1.", "frb.prelu."))


def test_is_dense_accepts_any_gap_free_permutation():
    a = torch.zeros(4, 6, 3, 3)
    assert ops._is_dense(a)
    assert ops._is_dense(a.contiguous(memory_format=torch.channels_last))
    assert ops._is_dense(a.permute(1, 0, 2, 3))
    assert not ops._is_dense(a[:, :4])                      # a channel slice has gaps
    assert not ops._is_dense(a[::2])
    assert ops._is_dense(torch.zeros(5, 1, 1, 7))           # size-1 dims carry no stride information


def test_cat_channels_padded_is_plain_cat_on_cpu_and_pads_nothing_when_aligned():
    x, y = torch.randn(2, 5, 3, 3), torch.randn(2, 3, 3, 3)
    out, pad = ops.cat_channels_padded((x, y))
    assert pad == 0 and torch.equal(out, torch.cat((x, y), 1))          # 8 channels: nothing to pad
    out, pad = ops.cat_channels_padded((x,))
    assert pad == 0 and out.shape[1] == 5                                # CPU tensors are never padded (no cuDNN alignment)


def test_pending_weight_gradients_can_be_discarded():
    ops._PENDING_WGRADS.append((torch.zeros(1), torch.zeros(1)))
    ops.discard_pending_weight_grads()
    assert not ops._PENDING_WGRADS
    ops.flush_weight_grads()                                 # nothing queued: no library call, no error without a GPU


def test_grad_marker_is_identity_without_a_callback():
    x = torch.randn(3, requires_grad=True)
    assert ops.grad_marker(x, 0) is x
    seen = []
    ops.set_grad_marker_callback(seen.append)
    try:
        y = ops.grad_marker(x, 7)
        y.sum().backward()
    finally:
        ops.set_grad_marker_callback(None)
    assert seen == [7] and torch.equal(x.grad, torch.ones(3))


def test_background_generator_order_end_and_errors():
    """ref datasets/dataloaderx.py:12-38: items in order, StopIteration at the end (and again afterwards); unlike the
    reference, a failing producer re-raises in the consumer instead of leaving it blocked."""
    from msml_b200.datasets import BackgroundGenerator
    g = BackgroundGenerator(iter(range(20)), None, max_prefetch=3)
    assert list(g) == list(range(20))
    with pytest.raises(StopIteration):
        next(g)

    def boom():
        yield 1
        raise KeyError("bad record")
    g = BackgroundGenerator(boom(), None)
    assert next(g) == 1
    with pytest.raises(KeyError):
        next(g)
    assert next(g, "done") == "done"
