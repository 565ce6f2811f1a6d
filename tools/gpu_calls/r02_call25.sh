#!/bin/bash
# Round-2 GPU call 25 (1 GPU): K-P peer-branch kernels (fm_peer_mul, mse) + FlatSGD test after its fix: parity suites that touch them.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_model.py tests/test_gpu_engine.py -m gpu -q --durations=5 > $O/r02z_pytest_peer.log 2>&1
echo "pytest rc=$? : $(tail -1 $O/r02z_pytest_peer.log)"
grep -E "FAILED|Error" $O/r02z_pytest_peer.log | head -20
