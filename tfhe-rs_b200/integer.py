"""Radix-integer multiplication on top of the batched KS -> PBS entry points
(FheUint64 x FheUint64, 32 blocks of 2+2 bits).

Host-side orchestration only: every ciphertext operation that is not a plain
wrapping add is a keyswitch + programmable bootstrap issued through the C ABI
with per-sample LUT indexes, exactly what the reference's integer layer does
with `execute_keyswitch_async` + `execute_pbs_async`
(backends/tfhe-cuda-backend/cuda/src/integer/integer.cuh:869-1000).

The algorithm restates the reference's CPU `unchecked_mul_parallelized`:
  * terms: for every block i of rhs, lhs shifted by i blocks with the bivariate
    LUT (x*y) % 4, and shifted by i+1 blocks with (x*y) / 4
    (tfhe/src/integer/server_key/radix_parallel/mul.rs:333-398; bivariate
    packing lhs*4 + rhs, shortint/server_key/bivariate_pbs.rs);
  * column-wise partial sums in chunks of 5 blocks (15 / 3), message and carry
    extraction, until every column holds at most 5 blocks
    (radix_parallel/sum.rs:14-160);
  * full carry propagation (sequential variant of full_propagate_parallelized).

The arithmetic on encrypted blocks is engine-agnostic (`BlockEngine`): the GPU
engine keeps blocks in device memory (torch int64 tensors: wrapping adds are
plumbing) and sends every LUT evaluation to libtfhe_cuda_backend_b200.so; the
tests additionally run the same host logic over a CPU engine backed by the
oracle (tests/ only).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

MSG_MOD = 4      # message modulus of PARAM_MESSAGE_2_CARRY_2
TOTAL_MOD = 16   # message * carry modulus
CHUNK = (TOTAL_MOD - 1) // (MSG_MOD - 1)  # = 5, ServerKey::max_sum_size(Degree(3))

LUT_MUL_LSB, LUT_MUL_MSB, LUT_MSG, LUT_CARRY = 0, 1, 2, 3
LUT_STATE, LUT_COMBINE, LUT_CARRY_BIT = 4, 5, 6
ST_NONE, ST_PROPAGATE, ST_GENERATE = 0, 1, 2


def lut_functions() -> List[List[int]]:
    """f tables over the 4-bit padded message space, in LUT index order."""
    lsb = [((x // MSG_MOD) * (x % MSG_MOD)) % MSG_MOD for x in range(TOTAL_MOD)]
    msb = [((x // MSG_MOD) * (x % MSG_MOD)) // MSG_MOD for x in range(TOTAL_MOD)]
    msg = [x % MSG_MOD for x in range(TOTAL_MOD)]
    carry = [x // MSG_MOD for x in range(TOTAL_MOD)]
    # carry-propagation states of a block value (message + incoming carry <= 7)
    state = [ST_GENERATE if x >= MSG_MOD else (ST_PROPAGATE if x == MSG_MOD - 1 else ST_NONE)
             for x in range(TOTAL_MOD)]
    # prefix operator on packed (cur * 4 + prev): a propagating block inherits the previous state
    combine = [((x % MSG_MOD) if (x // MSG_MOD) == ST_PROPAGATE else min(x // MSG_MOD, ST_GENERATE))
               for x in range(TOTAL_MOD)]
    carry_bit = [1 if x == ST_GENERATE else 0 for x in range(TOTAL_MOD)]
    return [lsb, msb, msg, carry, state, combine, carry_bit]


class BlockEngine:
    """What the radix algorithms need from a backend.  Blocks are rows of an
    array-like `[count, big_lwe_dimension + 1]` with wrapping u64 arithmetic."""

    def zeros(self, count: int):
        raise NotImplementedError

    def stack(self, rows: Sequence):
        raise NotImplementedError

    def add(self, a, b):
        raise NotImplementedError

    def scalar_mul(self, a, s: int):
        raise NotImplementedError

    def apply_luts(self, blocks, lut_ids: Sequence[int]):
        """KS -> PBS of every row of `blocks` with LUT `lut_ids[row]`."""
        raise NotImplementedError

    # batched plumbing (defaults work for numpy arrays and torch tensors alike)
    def take(self, blocks, idx: Sequence[int]):
        """rows `idx` of `blocks` as a new array"""
        return blocks[list(idx)]

    def cat(self, parts: Sequence):
        return self.stack([row for part in parts for row in part])

    def sum_groups(self, blocks, group: int):
        """wrapping sum of every `group` consecutive rows"""
        return blocks.reshape(blocks.shape[0] // group, group, blocks.shape[1]).sum(1)


def _bivariate_pack(engine: BlockEngine, lhs_rows, rhs_rows):
    """lhs * 4 + rhs (unchecked_apply_lookup_table_bivariate's linear part)."""
    return engine.add(engine.scalar_mul(lhs_rows, MSG_MOD), rhs_rows)


def compute_terms_for_mul_low(engine: BlockEngine, lhs, rhs, num_blocks: int):
    """All partial-product blocks in ONE batched LUT evaluation.  Returns
    `columns`: list (per output block) of lists of row handles."""
    lhs_idx, rhs_idx, lut_ids, col = [], [], [], []
    for i in range(num_blocks):          # rhs block
        for j in range(num_blocks - i):  # lhs block, result column i + j
            lhs_idx.append(j)
            rhs_idx.append(i)
            lut_ids.append(LUT_MUL_LSB)
            col.append(i + j)
            if i + j + 1 < num_blocks:   # carry part lands one block higher
                lhs_idx.append(j)
                rhs_idx.append(i)
                lut_ids.append(LUT_MUL_MSB)
                col.append(i + j + 1)
    packed = _bivariate_pack(engine, lhs[lhs_idx], rhs[rhs_idx])
    out = engine.apply_luts(packed, lut_ids)
    columns = [[] for _ in range(num_blocks)]
    for r, c in enumerate(col):
        columns[c].append(out[r])
    return columns


def partial_sum_columns(engine: BlockEngine, columns):
    """radix_parallel/sum.rs:14-160 on column lists; every round is one batched
    LUT evaluation (message + carry extraction of every full chunk).  The chunk
    sums of a round are ONE stacked gather + grouped sum, not one add per row."""
    n = len(columns)
    while any(len(c) > CHUNK for c in columns):
        flat, meta = [], []
        new_cols = [[] for _ in range(n)]
        for ci, colm in enumerate(columns):
            if len(colm) < CHUNK:
                new_cols[ci].extend(colm)
                continue
            full = (len(colm) // CHUNK) * CHUNK
            flat.extend(colm[:full])
            meta.extend([ci] * (full // CHUNK))
            new_cols[ci].extend(colm[full:])
        sums = engine.sum_groups(engine.stack(flat), CHUNK)
        pick, ids, dest = [], [], []
        for r, ci in enumerate(meta):
            pick.append(r)
            ids.append(LUT_MSG)
            dest.append(ci)
            if ci + 1 < n:
                pick.append(r)
                ids.append(LUT_CARRY)
                dest.append(ci + 1)
        out = engine.apply_luts(engine.take(sums, pick), ids)
        for r, d in enumerate(dest):
            new_cols[d].append(out[r])
        columns = new_cols
    # final column sums: pad every column to the longest with zero blocks
    width = max(1, max(len(c) for c in columns))
    zero = engine.zeros(1)[0]
    flat = [row for colm in columns for row in (list(colm) + [zero] * (width - len(colm)))]
    return engine.sum_groups(engine.stack(flat), width)


def _split_and_shift_carries(engine: BlockEngine, blocks, n: int):
    """Blocks may hold any value <= 15: extract message and carry of every
    block in one batched round and add each carry one block up (value <= 6).
    Returns the stacked [n, L] result."""
    out = engine.apply_luts(engine.take(blocks, [i for i in range(n) for _ in (0, 1)]), [LUT_MSG, LUT_CARRY] * n)
    msgs = engine.take(out, list(range(0, 2 * n, 2)))
    if n == 1:
        return msgs
    carries = engine.take(out, list(range(1, 2 * (n - 1), 2)))
    return engine.add(msgs, engine.cat([engine.zeros(1), carries]))


def full_propagate(engine: BlockEngine, blocks, num_blocks: int):
    """Sequential carry ripple (one LUT round per block) after the parallel
    message/carry split; kept as the simple cross-check of the parallel scan."""
    vals = _split_and_shift_carries(engine, blocks, num_blocks)
    out = []
    cur = vals[0]
    for i in range(num_blocks):
        if i + 1 < num_blocks:
            res = engine.apply_luts(engine.stack([cur, cur]), [LUT_MSG, LUT_CARRY])
            out.append(res[0])
            cur = engine.add(vals[i + 1], res[1])
        else:
            res = engine.apply_luts(engine.stack([cur]), [LUT_MSG])
            out.append(res[0])
    return engine.stack(out)


def full_propagate_parallel(engine: BlockEngine, blocks, num_blocks: int):
    """Carry propagation in 4 + ceil(log2(blocks)) batched LUT rounds instead of
    `blocks` sequential ones (role of the reference's parallel
    `propagate_single_carry_parallelized`, integer/server_key/radix_parallel/
    add.rs): split every block into message and carry, add the carry one block
    up (value <= 6), classify blocks as generate / propagate / none, run a
    Hillis-Steele prefix scan with the bivariate operator, turn the incoming
    state into a carry bit, add it and extract the message."""
    n = num_blocks
    vals = _split_and_shift_carries(engine, blocks, n)
    states = engine.apply_luts(vals, [LUT_STATE] * n)
    d = 1
    while d < n:
        idx = list(range(d, n))
        packed = engine.add(engine.scalar_mul(engine.take(states, idx), MSG_MOD),
                            engine.take(states, [i - d for i in idx]))
        comb = engine.apply_luts(packed, [LUT_COMBINE] * len(idx))
        states = engine.cat([states[:d], comb])
        d *= 2
    if n == 1:
        return engine.apply_luts(vals, [LUT_MSG])
    # carry into block i = [inclusive prefix state of blocks 0..i-1 == generate]
    bits = engine.apply_luts(states[: n - 1], [LUT_CARRY_BIT] * (n - 1))
    finals = engine.add(vals, engine.cat([engine.zeros(1), bits]))
    return engine.apply_luts(finals, [LUT_MSG] * n)


def unchecked_mul(engine: BlockEngine, lhs, rhs, parallel_carry: bool = True):
    """radix_parallel/mul.rs:437-470 (clean inputs: carries empty)."""
    num_blocks = lhs.shape[0]
    assert rhs.shape[0] == num_blocks
    columns = compute_terms_for_mul_low(engine, lhs, rhs, num_blocks)
    summed = partial_sum_columns(engine, columns)
    if parallel_carry:
        return full_propagate_parallel(engine, summed, num_blocks)
    return full_propagate(engine, summed, num_blocks)


# ---------------------------------------------------------------------------
# GPU engine
# ---------------------------------------------------------------------------
class CudaBlockEngine(BlockEngine):
    """Blocks live on the GPU as an int64 tensor [count, kN + 1]; LUT
    evaluations go through CudaServerKey.apply_lookup_table (C ABI)."""

    def __init__(self, server_key, lut_glwe_list_np: np.ndarray, glwe_dimension: int, polynomial_size: int):
        import torch

        from . import gpu

        self.torch, self.gpu = torch, gpu
        self.skey = server_key
        self.streams = server_key.streams
        self.luts = gpu.CudaGlweCiphertextList.from_glwe_ciphertext_list(
            lut_glwe_list_np, glwe_dimension, polynomial_size, self.streams)
        self.big = server_key.big_dim
        self.pbs_count = 0

    def from_numpy(self, a: np.ndarray):
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64))
        with self.torch.cuda.stream(self.streams.streams[0]):
            return t.to(self.streams.device(0))

    def to_numpy(self, t) -> np.ndarray:
        self.streams.synchronize()
        return t.cpu().numpy().view(np.uint64)

    def zeros(self, count):
        with self.torch.cuda.stream(self.streams.streams[0]):
            return self.torch.zeros((count, self.big + 1), dtype=self.torch.int64, device=self.streams.device(0))

    def stack(self, rows):
        with self.torch.cuda.stream(self.streams.streams[0]):
            return self.torch.stack(list(rows))

    def add(self, a, b):
        with self.torch.cuda.stream(self.streams.streams[0]):
            return a + b  # int64 add wraps mod 2^64

    def scalar_mul(self, a, s):
        with self.torch.cuda.stream(self.streams.streams[0]):
            return a * int(s)

    def take(self, blocks, idx):
        with self.torch.cuda.stream(self.streams.streams[0]):
            index = self.torch.tensor(list(idx), dtype=self.torch.int64).to(self.streams.device(0), non_blocking=True)
            return blocks.index_select(0, index)

    def cat(self, parts):
        with self.torch.cuda.stream(self.streams.streams[0]):
            return self.torch.cat(list(parts), dim=0)

    def sum_groups(self, blocks, group):
        with self.torch.cuda.stream(self.streams.streams[0]):
            return blocks.view(blocks.shape[0] // group, group, blocks.shape[1]).sum(1)

    def apply_luts(self, blocks, lut_ids):
        gpu = self.gpu
        count = blocks.shape[0]
        with self.torch.cuda.stream(self.streams.streams[0]):
            flat = blocks.contiguous().view(-1)
            ids = self.torch.tensor(list(lut_ids), dtype=self.torch.int64).to(self.streams.device(0))
        cts = gpu.CudaLweCiphertextList(gpu.CudaVec(flat, np.uint64), self.big, count)
        out = self.skey.apply_lookup_table(cts, self.luts, gpu.CudaVec(ids, np.uint64))
        self.pbs_count += count
        return out.d_vec.t.view(count, self.big + 1)


class CudaUnsignedRadixCiphertext:
    """integer/gpu/ciphertext: a vector of shortint blocks on one GPU."""

    def __init__(self, blocks):
        self.blocks = blocks

    @property
    def num_blocks(self):
        return self.blocks.shape[0]


class CudaRadixServerKey:
    """Mirror of integer::gpu::CudaServerKey for the multiplication path
    (integer/gpu/server_key/radix/mul.rs:167)."""

    def __init__(self, server_key, lut_glwe_list_np=None, glwe_dimension=1, polynomial_size=2048):
        if lut_glwe_list_np is None:  # the 7 accumulators of lut_functions(), 2+2-bit blocks, one padding bit
            from . import algorithms

            delta = (1 << 63) // TOTAL_MOD
            lut_glwe_list_np = np.stack([
                algorithms.generate_programmable_bootstrap_glwe_lut(polynomial_size, glwe_dimension + 1, TOTAL_MOD,
                                                                    delta, f) for f in lut_functions()])
        self.engine = CudaBlockEngine(server_key, lut_glwe_list_np, glwe_dimension, polynomial_size)

    def unchecked_mul(self, lhs: CudaUnsignedRadixCiphertext, rhs: CudaUnsignedRadixCiphertext):
        return CudaUnsignedRadixCiphertext(unchecked_mul(self.engine, lhs.blocks, rhs.blocks))
