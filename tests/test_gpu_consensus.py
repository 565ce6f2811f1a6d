"""GPU parity of the consensus-loss kernels (csrc/seg_loss.cu) against the reference goldens
(tests/golden/consensus.npz, ref tricks/consensus_loss.py:63-178) and the oracle.  First passed on a B200 in round 2
(gpurun_out/r02_check_consensus.log: 15 passed)."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import assert_close, dev, host, need_gpu
from oracle import consensus

pytestmark = pytest.mark.gpu

CASES = ["binary_missing", "four_blobs", "four_blobs_all_all", "four_blobs_idx_all", "underflow", "seg_shape"]


def case(g, name):
    alpha, beta, rp, rkl = [str(v) for v in g[name + ".cfg"]]
    blobs = g[name + ".blobs"].astype(np.int64)
    ids, dense = np.unique(blobs, return_inverse=True)          # the kernel wants ids 0 .. K-1 (the mirror's documented contract)
    return (g[name + ".logit"], dense.reshape(blobs.shape), g[name + ".target"].astype(np.int64), len(ids),
            (float(alpha), float(beta), rp, rkl))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cl", [False, True])
def test_consensus_matches_reference_golden(name, cl):
    need_gpu()
    from msml_b200.tricks import StructureConsensuLossFunction
    g = load_golden("consensus")
    logit, blobs, target, K, cfg = case(g, name)
    crit = StructureConsensuLossFunction(*cfg, num_blobs=K)
    z = dev(logit).contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format).requires_grad_(True)
    loss = crit(z, dev(blobs), dev(target))
    assert loss.shape == () and loss.dtype == torch.float32
    loss.backward()
    want = float(g[name + ".loss"])
    assert abs(float(loss) - want) <= 5e-5 * abs(want), (float(loss), want)
    assert_close(host(z.grad), g[name + ".dlogit"], 2e-3, atol_frac=2e-5, what="dlogit")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_consensus_half_precision_inputs_vs_oracle(dtype):
    need_gpu()
    from msml_b200 import ops
    rng = np.random.default_rng(3)
    N, C, H, W = 5, 3, 33, 41                                   # ragged chunk (1353 pixels), three classes
    z0 = torch.from_numpy(rng.normal(size=(N, C, H, W)).astype(np.float32)).to(dtype)
    blobs = rng.integers(0, 3, size=(N, H, W))
    blobs[4][blobs[4] == 1] = 2                                 # blob 1 is missing from the last sample
    target = np.array([2, 0, 1])[blobs]
    z = z0.cuda().requires_grad_(True)
    loss = ops.consensus_loss(z, dev(blobs), dev(target), 10.0, 5.0, "idx", "idx", num_blobs=3)
    (2.5 * loss).backward()                                     # upstream gradient is applied inside the kernel
    want, dz = consensus.consensus_loss(z0.float().numpy(), blobs, target, 10.0, 5.0)
    assert abs(float(loss) - want) <= 1e-4 * abs(want)
    assert z.grad.dtype == dtype
    assert_close(host(z.grad), 2.5 * dz, 2e-2, atol_frac=1e-2, what="dlogit")


def test_consensus_full_size_vs_oracle_and_poison():
    need_gpu()
    from msml_b200 import ops
    from msml_b200.tricks import StructureConsensuLossFunction
    rng = np.random.default_rng(4)
    N, H, W = 128, 112, 112                                     # final_seg of BASELINE config 3
    z0 = rng.normal(size=(N, 2, H, W)).astype(np.float32)
    msk = np.zeros((N, H, W), np.int64)
    for n in range(N - 1):                                      # the last image is unoccluded
        h0, w0 = rng.integers(0, 60, size=2)
        msk[n, h0:h0 + 40, w0:w0 + 40] = 1
    z = dev(z0).requires_grad_(True)
    crit = StructureConsensuLossFunction(10.0, 5.0, "idx", "idx")
    loss = crit(z, dev(msk)[:, None], dev(msk))
    loss.backward()
    want, dz = consensus.consensus_loss(z0, msk, msk, 10.0, 5.0)
    assert abs(float(loss) - want) <= 5e-5 * abs(want)
    assert_close(host(z.grad), dz, 2e-3, atol_frac=2e-5, what="dlogit")
    # deterministic: partial sums are combined in a fixed order
    assert float(crit(z, dev(msk)[:, None], dev(msk))) == float(loss)
    # one blob through the per-blob entry point (pixels outside the mask belong to no blob)
    one = crit.structure_via_consensus_over_blob(dev(msk) == 1, dev(msk), z.detach())
    want1, _ = consensus.consensus_loss(z0, msk, msk, 10.0, 5.0, want_grad=False, ids=[1])
    assert abs(float(one) - want1) <= 5e-5 * abs(want1)
    # poisoned inputs: an id outside [0, K), labels that differ inside a blob, a label outside [0, C)
    bad = msk.copy(); bad[0, 0, 0] = 7
    assert np.isnan(float(crit(z.detach(), dev(bad), dev(msk))))
    lab = msk.copy(); lab[3, 5, 5] = 1 - lab[3, 5, 5]
    assert np.isnan(float(crit(z.detach(), dev(msk), dev(lab))))
    assert np.isnan(float(crit(z.detach(), dev(msk), dev(msk * 5))))
    with pytest.raises(RuntimeError):
        ops.consensus_loss(torch.zeros(2, 7, 4, 4, device="cuda"), dev(msk[:2, :4, :4]), dev(msk[:2, :4, :4]))     # C > 4
    with pytest.raises(RuntimeError):
        ops.consensus_loss(torch.zeros(2, 2, 4, 4), torch.zeros(2, 4, 4), torch.zeros(2, 4, 4))                       # CPU tensors
