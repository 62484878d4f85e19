"""Multi-GPU plumbing for the PBS path: one process per GPU.

The path shards with no exchange step: every LWE is independent, keys are
replicated.  The reference splits the batch with
`get_num_inputs_on_gpu` / `get_gpu_offset`
(backends/tfhe-cuda-backend/cuda/src/utils/helper_multi_gpu.cu:64-101) and
uploads each key from the host once per GPU (core_crypto/gpu/ffi.rs:744-784).
Here each rank owns a contiguous slice of the LWE list and the keys are
replicated with ONE NCCL broadcast per key over NVLink at set-up; the steady
state has no collective.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import gpu


def get_num_inputs_on_gpu(total: int, rank: int, world: int) -> int:
    """helper_multi_gpu.cu:64-83: the first `total % world` ranks get one more."""
    base, rem = divmod(total, world)
    return base + (1 if rank < rem else 0)


def get_gpu_offset(total: int, rank: int, world: int) -> int:
    """helper_multi_gpu.cu:85-101."""
    base, rem = divmod(total, world)
    return rank * base + min(rank, rem)


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    off = get_gpu_offset(total, rank, world)
    return off, off + get_num_inputs_on_gpu(total, rank, world)


def broadcast_vec(vec: Optional[gpu.CudaVec], length: int, np_dtype, streams: gpu.CudaStreams, src: int = 0,
                  group=None) -> gpu.CudaVec:
    """Replicate a device buffer from rank `src` to every rank (NCCL over
    NVLink when the process group is nccl; gloo moves it through the host in
    the CPU tests)."""
    import torch.distributed as dist

    backend = dist.get_backend(group)
    if vec is None:
        vec = gpu.CudaVec.new(length, streams, np_dtype=np_dtype)
    assert len(vec) == length
    if backend == "nccl":
        streams.synchronize()
        dist.broadcast(vec.t, src=src, group=group)
        torch.cuda.current_stream(vec.t.device).synchronize()
    else:
        h = vec.t.cpu()
        dist.broadcast(h, src=src, group=group)
        vec.t.copy_(h)
    return vec


def broadcast_host_array(arr: Optional[np.ndarray], length: int, np_dtype, src: int = 0, group=None) -> np.ndarray:
    """Host-side twin used by the gloo CPU tests of the sharding logic."""
    import torch.distributed as dist

    if arr is None:
        arr = np.zeros(length, dtype=np_dtype)
    t = torch.from_numpy(arr.view(np.int64) if arr.dtype == np.uint64 else arr)
    dist.broadcast(t, src=src, group=group)
    return arr
