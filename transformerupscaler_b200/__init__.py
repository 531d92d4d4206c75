"""transformerupscaler_b200 — B200 (sm_100a) engine for the forward pass of TransformerUpscaler's models.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/tu_b200.h) -> libtu_b200.so (built in-tree)
  _lib.py    ctypes binding of the C ABI (fails loudly when the library is missing)
  packing.py state_dict -> packed device weights
  engine.py  forward dispatch (torch.library custom op ``tu::forward``) and single-op wrappers
  models/    drop-in ``TransformerModel`` classes (same ctor, state_dict and forward signature as the reference)
"""
__version__ = "0.1.0"
