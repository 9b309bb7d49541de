"""Oracle: the exact evaluation ORDER of the ATen CPU kernels the bit-exact contracts hang on,
spelled out in numpy (test infrastructure; small cases only — pure-Python loops).

The reference computes sample indices with torch.sum / torch.cumsum / torch.searchsorted
(src/models/SimpleNeRF17.py:388-401) and the occupancy mask with F.grid_sample
(src/models/SimpleTensoRF09.py:1344).  Those live in PyTorch, not in the reference checkout
(pinned there as pytorch 2.0.0, EnvironmentData/SimpleRF.yml:187; the oracle runs torch 2.11).
The CUDA kernels reproduce these orders; tests check  numpy-restatement == torch op == CUDA.

* row_sum_f32:  ATen/native/cpu/SumKernel.cpp `vectorized_inner_sum` -> `row_sum` ->
  `multi_row_sum`: 8 fp32 lanes (the Sum kernel is built for AVX2 even when the process
  reports AVX512), 4 ILP accumulators, cascade levels of 2^max(4, ceil_log2(n/32)/4) groups,
  left-over vectors into accumulator 0, accumulators 1..3 folded into 0, scalar tail summed
  first, then the 8 lanes added to it in lane order.
* cumsum_f32:  ATen/native/cpu/ReduceOpsKernel.cpp `cumsum_cpu_kernel`: sequential
  accumulation in double (`at::acc_type<float,false>`), every prefix rounded to fp32.
* searchsorted_right: count of elements <= value (upper bound).
* trilinear_positive:  `grid_sampler_3d` (align_corners=True, zeros padding) > 0 on a {0,1}
  volume == OR over the <= 8 in-bounds corners whose fp32 weight product is > 0.
"""
import math

import numpy as np

F32 = np.float32
LANES = 8
ILP = 4
LEVELS = 4


def _ceil_log2(v):
    return 0 if v <= 1 else int(math.ceil(math.log2(v)))


def row_sum_f32(x):
    """x: 1-D float32 array (a contiguous reduced row with len(x) >= 8)."""
    x = np.asarray(x, dtype=F32)
    n = x.shape[0]
    if n < LANES:                       # scalar_inner_sum -> row_sum: 4 ILP partial sums, no SIMD lanes
        part = np.zeros(ILP, dtype=F32)
        full = n // ILP
        for i in range(full):
            part += x[i * ILP:(i + 1) * ILP]
        for k in range(full * ILP, n):
            part[0] = F32(part[0] + x[k])
        for k in range(1, ILP):
            part[0] = F32(part[0] + part[k])
        return part[0]
    nvec = n // LANES
    vecs = x[:nvec * LANES].reshape(nvec, LANES)
    groups = nvec // ILP
    power = max(4, _ceil_log2(groups) // LEVELS)
    step = 1 << power
    acc = np.zeros((LEVELS, ILP, LANES), dtype=F32)
    i = 0
    while i + step <= groups:
        for _ in range(step):
            acc[0] += vecs[i * ILP:(i + 1) * ILP]
            i += 1
        for j in range(1, LEVELS):
            acc[j] += acc[j - 1]
            acc[j - 1] = 0
            if i & ((step - 1) << (j * power)):
                break
    while i < groups:
        acc[0] += vecs[i * ILP:(i + 1) * ILP]
        i += 1
    for j in range(1, LEVELS):
        acc[0] += acc[j]
    part = acc[0].copy()
    for k in range(groups * ILP, nvec):
        part[0] += vecs[k]
    for k in range(1, ILP):
        part[0] += part[k]
    total = F32(0)
    for k in range(nvec * LANES, n):
        total = F32(total + x[k])
    for k in range(LANES):
        total = F32(total + part[0][k])
    return total


def cumsum_f32(x):
    acc = np.float64(0)
    out = np.empty(len(x), dtype=F32)
    for i, v in enumerate(x):
        acc = acc + np.float64(v)
        out[i] = F32(acc)
    return out


def searchsorted_right(sorted_row, v):
    lo, hi = 0, len(sorted_row)
    while lo < hi:
        mid = (lo + hi) // 2
        if sorted_row[mid] <= v:
            lo = mid + 1
        else:
            hi = mid
    return lo


def inverse_cdf_row(bins, weights, u):
    """One ray of SimpleNeRF17.py:385-417 in explicit fp32/fp64 steps (no FMA anywhere).
    Returns (samples, below, above, cdf)."""
    w = (np.asarray(weights, dtype=F32) + F32(1e-5)).astype(F32)
    tot = row_sum_f32(w)
    pdf = (w / tot).astype(F32)
    cdf = np.concatenate([np.zeros(1, dtype=F32), cumsum_f32(pdf)])
    nb = cdf.shape[0]
    samples = np.empty(len(u), dtype=F32)
    below = np.empty(len(u), dtype=np.int64)
    above = np.empty(len(u), dtype=np.int64)
    for j, uj in enumerate(np.asarray(u, dtype=F32)):
        ind = searchsorted_right(cdf, uj)
        b = max(ind - 1, 0)
        a = min(ind, nb - 1)
        denom = F32(cdf[a] - cdf[b])
        if denom < F32(1e-5):
            denom = F32(1)
        t = F32(F32(uj - cdf[b]) / denom)
        samples[j] = F32(bins[b] + F32(t * F32(bins[a] - bins[b])))
        below[j], above[j] = b, a
    return samples, below, above, cdf


def trilinear_positive(volume, bbox, pts):
    """volume: bool/float [Z,Y,X]; bbox float32 [2,3]; pts float32 [N,3] -> bool [N].
    Coordinate arithmetic is ATen's: normalise ((p-b0)/(b1-b0))*2-1 in fp32, unnormalise
    ((c+1)/2)*(size-1) (GridSampler.h:27-31), floor, corner weights as fp32 products."""
    vol = np.asarray(volume) > 0
    Z, Y, X = vol.shape
    b0 = np.asarray(bbox[0], dtype=F32)
    size = (np.asarray(bbox[1], dtype=F32) - b0).astype(F32)
    out = np.zeros(len(pts), dtype=bool)
    dims = (X, Y, Z)
    for n, p in enumerate(np.asarray(pts, dtype=F32)):
        c = ((((p - b0).astype(F32) / size).astype(F32) * F32(2)).astype(F32) - F32(1)).astype(F32)
        ix = [F32(F32(F32(c[a] + F32(1)) / F32(2)) * F32(dims[a] - 1)) for a in range(3)]
        f0 = [F32(np.floor(v)) for v in ix]
        hit = False
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    d = (dx, dy, dz)
                    idx = [int(f0[a]) + d[a] for a in range(3)]
                    if any(idx[a] < 0 or idx[a] >= dims[a] for a in range(3)):
                        continue
                    wts = []
                    for a in range(3):
                        hi_c = F32(f0[a] + F32(1))
                        wts.append(F32(ix[a] - f0[a]) if d[a] else F32(hi_c - ix[a]))
                    wgt = F32(F32(wts[0] * wts[1]) * wts[2])
                    if wgt > 0 and vol[idx[2], idx[1], idx[0]]:
                        hit = True
        out[n] = hit
    return out
