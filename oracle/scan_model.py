"""Numpy model of the parallel evaluation of numpy's sequential float32 cumsum (csrc/seed.cu::seed_scan_kernel, the k-means++
scan behind sklearn `_kmeans_plusplus`, sklearn/cluster/_kmeans.py:252-257).  TEST INFRASTRUCTURE (see oracle/__init__.py): it
states the arithmetic argument of the kernel in executable form and is held to `np.cumsum` bit for bit by
tests/test_host_logic.py; the kernel itself is held to `np.cumsum` on the GPU (tests/test_gpu_fit.py).

Inside one binade [2^e, 2^(e+1)) the running sum is S * u with u = 2^(e-23) and S an integer in [2^23, 2^24), and
RN(S*u + v) = (S + q) * u with q = v / u rounded to the nearest integer, an exact tie going to the side that makes S + q even.
So a block's contribution is an integer that depends on the block's start only through the PARITY of S; blocks are summed
independently for both parities and combined by a scan over the maps "parity in -> (increment, parity out)".  Where the sum
leaves the binade the block that crosses is redone with the float chain itself from its exact start value.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
BLOCK = 128


def _block_increments(v: np.ndarray, scale: np.float32):
    """(T0, T1): units added by the values `v` when the block starts on an even / odd S."""
    x = np.minimum(v * scale, F32(16777216.0)).astype(F32)          # exact scaling; beyond the binade anyway
    fl = np.floor(x)
    fr = x - fl
    q = fl.astype(np.int64)
    up = (fr > F32(0.5)).astype(np.int64)
    tie = fr == F32(0.5)
    out = []
    for p in (0, 1):
        t = 0
        for qi, ui, ti in zip(q.tolist(), up.tolist(), tie.tolist()):
            t += qi + (((p + t + qi) & 1) if ti else ui)
        out.append(t)
    return out


def boundary_sums(values: np.ndarray, blocks_per_pass: int = 512) -> np.ndarray:
    """The sequential float32 sum of `values` at the end of every BLOCK values (and at the end), computed block-parallel."""
    v = np.asarray(values, F32)
    n_blocks = (len(v) + BLOCK - 1) // BLOCK
    out = np.zeros(n_blocks, F32)
    s, blk = F32(0.0), 0

    def chain(s0, b):
        for x in v[b * BLOCK:(b + 1) * BLOCK]:
            s0 = F32(s0 + x)
        return s0
    while blk < n_blocks:
        m, e = np.frexp(s)                                           # s = m * 2^e, m in [0.5, 1)  ->  binade exponent e - 1
        if s == 0 or e - 1 < -100 or e - 1 > 123:
            s = chain(s, blk)
            out[blk] = s
            blk += 1
            continue
        scale, u = F32(2.0 ** (23 - (e - 1))), F32(2.0 ** ((e - 1) - 23))
        S = int(s * scale)
        nb = min(n_blocks - blk, blocks_per_pass, 2 * blk + 4)
        maps = [_block_increments(v[(blk + t) * BLOCK:(blk + t + 1) * BLOCK], scale) for t in range(nb)]   # independent
        par, acc, done = S & 1, 0, nb
        for t, (t0, t1) in enumerate(maps):                          # the prefix scan, written as the fold it computes
            inc = t1 if par else t0
            if S + acc + inc >= 1 << 24:                             # leaves the binade inside this block
                s = chain(F32(F32(S + acc) * u), blk + t)
                out[blk + t] = s
                done = t
                break
            acc += inc
            par ^= inc & 1
            out[blk + t] = F32(F32(S + acc) * u)
        if done == nb:
            s = out[blk + nb - 1]
            blk += nb
        else:
            blk += done + 1
    return out
