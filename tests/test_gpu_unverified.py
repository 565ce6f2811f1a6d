"""Kernels that have not yet executed on a GPU (written after the round's GPU budget was spent).

Their parity checks live in tests/unverified/ and run here in a SUBPROCESS, so that a faulting kernel cannot take the
CUDA context of the main test process with it; the outcome is reported as xfail / xpass and never fails the run.
Once a file has passed on a B200 its tests move into the regular test_gpu_*.py files.
"""
import os
import subprocess
import sys

import pytest

from gpu_util import need_gpu

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_checks(name):
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(HERE, "unverified", name)], capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0


@pytest.mark.xfail(strict=False, reason="csrc/seg_loss.cu (consensus loss, SURVEY 8f-4) has never run on a GPU; its logic is "
                                        "checked under CPU emulation (tests/test_emu_kernels.py)")
def test_consensus_loss_kernels_first_gpu_run():
    need_gpu()
    run_checks("check_consensus.py")


@pytest.mark.xfail(strict=False, reason="csrc/pfc_sgd_kernels.cuh (fused PartialFC SGD, SURVEY 8f-2) has never run on a GPU; its "
                                        "logic is checked under CPU emulation (tests/test_emu_kernels.py)")
def test_fused_pfc_sgd_first_gpu_run():
    need_gpu()
    run_checks("check_pfc_sgd.py")


@pytest.mark.xfail(strict=False, reason="msml_b200/datasets/dataloaderx.py (SURVEY 8f-4 data path) has never run on a GPU")
def test_dataloaderx_first_gpu_run():
    need_gpu()
    run_checks("check_dataloaderx.py")


@pytest.mark.xfail(strict=False, reason="edge-case probes of the PartialFC head written without GPU access (rank with no positive row, "
                                        "batch of one, repeated class, two-class shard); never run yet")
def test_head_edge_case_probes_first_gpu_run():
    need_gpu()
    run_checks("check_head_edge_cases.py")
