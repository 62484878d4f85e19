"""tfhe-rs_b200 -- Blackwell-native programmable-bootstrapping engine behind the
tfhe-cuda-backend C ABI (see include/tfhe_b200.h).

Layout:
  csrc/   hand-written sm_100a CUDA kernels + the C-ABI entry points
  lib/    the built libtfhe_cuda_backend_b200.so (git-ignored artefact)
  gpu.py  host-side mirror of tfhe-rs core_crypto::gpu (Rust is not in this
          image) -- same names / argument order as the Rust callers
  server_key.py  KS->PBS atomic pattern on device batches (shortint
          StandardAtomicPatternServerKey, atomic_pattern/standard.rs:162-199)
  multi_gpu.py   rank-sharded batches + one NCCL broadcast per key
"""
from . import _lib  # noqa: F401
from ._lib import build, lib  # noqa: F401

__all__ = ["build", "lib"]
