"""CPU oracle for the MSML hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy fp64 / plain torch-CPU fp32) of the
reference's algorithm for the mask-fusion + PartialFC hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  Nothing under ``msml_b200/`` imports it: the product path calls
the sm_100a CUDA library through the C ABI and raises if that library is
missing.

Parity pinning: the reference ships no tests, golden vectors or KATs for this
path, so the oracle is pinned against outputs of the
reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` read-only)
and committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
checks every oracle function against those vectors.

Modules
  fm_tail      FMCnn elementwise tail          ref backbones/fm/fmoperator.py:277-311
  dap          DAP (PixelShuffle+AvgPool)      ref backbones/osb/unet.py:158-161,223
  margins      AMArcFace / AMCosFace / Softmax ref headers/margin_losses.py:41-68,241-305,356-418
  partial_fc   PartialFC sharding/sample/step  ref headers/partial_fc.py:19-177
  consensus    structure-via-consensus seg loss ref tricks/consensus_loss.py:63-178
  model_cpu    torch-CPU fp32 MSML backbone    ref backbones/{msml,frb/iresnet,osb/unet,fm/fmoperator}.py
  detfill      deterministic weight fill shared by the golden generator and tests
"""
