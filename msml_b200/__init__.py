"""msml_b200 — B200-native (sm_100a) implementation of the MSML data-parallel hot path.

Drop-in mirrors of the reference's operator API (ygtxr1997/MSML):
  msml_b200.backbones.MSML, msml_b200.backbones.fm.FMCnn / FMNone,
  msml_b200.headers.partial_fc.PartialFC, msml_b200.headers.margin_losses.{Softmax,AMCosFace,AMArcFace}
all of which call hand-written CUDA kernels through the C ABI in include/msml_b200.h
(libmsml_b200.so, loaded with ctypes).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
