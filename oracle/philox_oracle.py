"""ORACLE (test infrastructure): numpy restatement of the arithmetic behind the reference's
`torch.rand(N, C, H-6, W-6, device='cuda')` (R/utils/utils_modules.py:49).  That arithmetic
lives in PyTorch/cuRAND, not in /root/reference:

* Philox4x32-10 (Salmon et al., SC'11; cuRAND `curand_philox4x32_x.h`), pinned here by the
  Random123 known-answer vectors in `tests/test_oracle_golden.py`.
* torch 2.11 `ATen/native/cuda/DistributionTemplates.h:50-62` (`calc_execution_policy`),
  `:65-90` (`distribution_elementwise_grid_stride_kernel`) and `:485-505` (uniform transform):
  block 256, grid = min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256)); thread `idx` runs
  `curand_init(seed, idx, offset)` and on trip t writes component ii of its t-th
  `curand_uniform4` draw to element `idx + Tn*(4t + ii)`, Tn = 256*grid; the generator offset
  then advances by ((numel-1)/(Tn*4)+1)*4.
* cuRAND `_curand_uniform`: x * 2^-32 + 2^-33 in fp32 -> (0, 1]; torch maps 1.0 -> 0.0.

The bit-exact check against the real `torch.rand` on a B200 is `tests/test_gpu_dropblock.py`.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over counter arrays (uint32).  Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def curand_uniform_from_u32(x: np.ndarray) -> np.ndarray:
    """cuRAND `_curand_uniform`: fp32 x * 2^-32 + 2^-33 (the product is exact, one rounding)."""
    xf = x.astype(np.float32)                      # round-to-nearest-even, like I2F.U32
    return (xf * np.float32(2.3283064e-10) + np.float32(2.3283064e-10 / 2.0)).astype(np.float32)


def torch_rand_grid(numel: int, sm_count: int = 148, max_threads_per_sm: int = 2048) -> int:
    return min(sm_count * (max_threads_per_sm // 256), (numel + 255) // 256)


def torch_rand_offset_increment(numel: int, sm_count: int = 148, max_threads_per_sm: int = 2048) -> int:
    tn = 256 * torch_rand_grid(numel, sm_count, max_threads_per_sm)
    return ((numel - 1) // (tn * 4) + 1) * 4


def torch_cuda_rand_u32(seed: int, offset: int, numel: int, sm_count: int = 148,
                        max_threads_per_sm: int = 2048) -> np.ndarray:
    """Raw 32-bit Philox words in torch.rand's element order."""
    tn = 256 * torch_rand_grid(numel, sm_count, max_threads_per_sm)
    p = np.arange(numel, dtype=np.int64)
    t = p // (4 * tn)
    rem = p % (4 * tn)
    ii = rem // tn
    idx = rem % tn
    ctr = (offset // 4) + t                         # curand skipahead: offset counts 32-bit words
    outs = philox4x32_10(ctr & 0xFFFFFFFF, ctr >> 32, idx & 0xFFFFFFFF, idx >> 32,
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    stacked = np.stack(outs, axis=0)
    return stacked[ii, np.arange(numel)]


def torch_cuda_rand(seed: int, offset: int, numel: int, sm_count: int = 148,
                    max_threads_per_sm: int = 2048) -> np.ndarray:
    u = curand_uniform_from_u32(torch_cuda_rand_u32(seed, offset, numel, sm_count, max_threads_per_sm))
    return np.where(u == np.float32(1.0), np.float32(0.0), u)


def threshold_hi_u32() -> int:
    """Smallest x whose uniform rounds to exactly 1.0f; torch then maps it to 0.0
    (DistributionTemplates.h `uniform_real` reverse-bound), which IS < gamma.  About 7 such
    words occur per 237 M-uniform forward, so the mask builder must honour them."""
    lo, hi = 0, (1 << 32) - 1
    while lo < hi:
        mid = (lo + hi) // 2
        if curand_uniform_from_u32(np.array([mid], dtype=np.uint32))[0] == np.float32(1.0):
            hi = mid
        else:
            lo = mid + 1
    return lo


def threshold_u32(gamma: float) -> int:
    """Smallest T such that (uniform(x) < float32(gamma)) == (x < T) for every uint32 x below
    `threshold_hi_u32()` (uniform() is monotone non-decreasing in x).  The CUDA mask builder
    compares raw Philox words: centre = (x < T) or (x >= threshold_hi_u32())."""
    g = np.float32(gamma)
    lo, hi = 0, 1 << 32                      # invariant: f(lo-1) true (or lo == 0), f(hi) false (or hi == 2^32)
    while lo < hi:
        mid = (lo + hi) // 2
        if curand_uniform_from_u32(np.array([mid], dtype=np.uint32))[0] < g:
            lo = mid + 1
        else:
            hi = mid
    return lo
