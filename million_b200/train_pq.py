"""Codebook training on the GPU and the `.fvecs` sample format.

The reference trains its codebooks with faiss on the CPU — `train_pq` (scripts/utils/pq_utils.py:586-609):
`faiss.IndexPQ(d, M, nbits).train(X)` with `cp.niter = 25` — from key/value samples dumped as `.fvecs`
(scripts/utils/fvecio.py:23-43) and saves them as `(M, 2**nbits, d/M)` fp32 tensors `*_cent_{M}_{nbits}.pq.pt`
(scripts/modeldb/main_pq.py:224-236).  faiss (faiss-cpu 1.9.0.post1, requirements.txt:18) is not in this image and its
sources are not part of the reference, so this module restates the PUBLISHED algorithm of `ProductQuantizer::train` /
`Clustering::train`: an independent k-means per sub-space — at most 256 training points per centroid (random sub-sample),
centroids initialised with k distinct random training points, `niter` Lloyd iterations, an empty cluster re-seeded by splitting
the most populated one with a relative perturbation of 1/1024.  The random streams differ from faiss's, so trained centroids
are not bit-comparable with faiss's (parity for this row is unpinned; the invariants are tested instead: monotone objective,
agreement with a numpy restatement of the same iteration from the same initialisation).

What is B200-native here: the ASSIGNMENT step is the library's own exact encoder (`million_pq_encode`: arg-min of the fp32
squared distance, lowest index on ties — the very operation the trained codebook will be used with), run over all sub-spaces at
once; the update step is two `index_add_` reductions.  There is no CPU fallback.
"""
import os

import numpy as np
import torch

from . import ops


# ---------------------------------------------------------------------------------------------- .fvecs (fvecio.py:23-43)


def read_fvecs(filename):
    """(n, d) float32 from the reference's `.fvecs` layout: per vector an int32 `d` (little endian) followed by d float32."""
    raw = np.fromfile(filename, dtype=np.int32)
    if raw.size == 0:
        return np.zeros((0, 0), dtype=np.float32)
    d = int(raw[0])
    if d <= 0 or raw.size % (d + 1):
        raise ValueError(f"{filename}: not a fixed-dimension .fvecs file (first header {d}, {raw.size} words)")
    rows = raw.reshape(-1, d + 1)
    if not (rows[:, 0] == d).all():
        raise ValueError(f"{filename}: vectors of different dimensions")
    return rows[:, 1:].view(np.float32).copy()


def write_fvecs(filename, vecs, mode='ab'):
    """Append (default, as fvecio.py:36-43) or write vectors in `.fvecs` layout."""
    vecs = np.ascontiguousarray(np.asarray(vecs, dtype=np.float32))
    if vecs.ndim != 2:
        raise ValueError("write_fvecs expects (n, d)")
    if mode == 'ab' and not os.path.exists(filename):
        mode = 'wb'
    n, d = vecs.shape
    out = np.empty((n, d + 1), dtype=np.int32)
    out[:, 0] = d
    out[:, 1:] = vecs.view(np.int32)
    with open(filename, mode) as f:
        f.write(out.tobytes())


def save_centroids(cent, path):
    """`(M, 2**nbits, d/M)` fp32 tensor, the `.pq.pt` file main_pq.py:224-236 writes and :262-270 loads."""
    torch.save(cent.detach().to('cpu', torch.float32).contiguous(), path)


def load_centroids(path, device='cuda'):
    return torch.load(path, map_location=device)


# ---------------------------------------------------------------------------------------------- k-means per sub-space


def _assign(X4, cent):
    """codes (n, M) int64 of X4 (1, 1, n, d) fp32 against cent (M, C, dm) fp32: the library's exact encoder."""
    C = cent.shape[1]
    codes = ops.pq_encode(X4, cent, out_dtype=torch.uint8 if C <= 256 else torch.uint16)
    return codes[0, 0].to(torch.int64)


def kmeans_step(X, cent):
    """One Lloyd iteration for all sub-spaces at once.  X (n, d) fp32 CUDA, cent (M, C, dm) fp32 -> (new cent, codes (n, M),
    objective = sum of squared distances to the OLD centroids, counts (M, C))."""
    n, d = X.shape
    M, C, dm = cent.shape
    codes = _assign(X.view(1, 1, n, d), cent)
    flat = (codes + torch.arange(M, device=X.device) * C).reshape(-1)               # (n*M,) row of the (M*C, dm) table
    pts = X.view(n, M, dm).reshape(n * M, dm)
    sums = torch.zeros(M * C, dm, dtype=torch.float32, device=X.device).index_add_(0, flat, pts)
    counts = torch.bincount(flat, minlength=M * C).view(M, C)
    obj = float(((pts - cent.view(M * C, dm)[flat]) ** 2).sum())
    new = torch.where(counts.view(M, C, 1) > 0, sums.view(M, C, dm) / counts.clamp(min=1).view(M, C, 1).float(), cent)
    return new, codes, obj, counts


def _split_empty(cent, counts, gen):
    """faiss's rule for empty clusters (Clustering.cpp `split_clusters`): an empty cluster takes a copy of a populated one
    (picked with probability proportional to its size), the two are moved apart by a relative 1/1024, the points are shared."""
    M, C, dm = cent.shape
    eps = 1.0 / 1024.0
    counts = counts.clone().cpu()
    cent_h = cent.cpu()
    n_split = 0
    for m in range(M):
        empty = (counts[m] == 0).nonzero().flatten().tolist()
        for ci in empty:
            p = (counts[m].float() - 1).clamp(min=0)
            if float(p.sum()) <= 0:
                break
            cj = int(torch.multinomial(p / p.sum(), 1, generator=gen))
            c = cent_h[m, cj].clone()
            sign = torch.where(torch.arange(dm) % 2 == 0, 1.0, -1.0)
            cent_h[m, ci] = c * (1 + sign * eps)
            cent_h[m, cj] = c * (1 - sign * eps)
            counts[m, ci] = counts[m, cj] // 2
            counts[m, cj] -= counts[m, ci]
            n_split += 1
    return cent_h.to(cent.device), n_split


def train_pq(X, M, nbits, niter=25, seed=1234, device='cuda', max_points_per_centroid=256, verbose=False, return_stats=False):
    """train_pq of the reference (pq_utils.py:586-609): X (n, d) float32 (numpy or torch) -> cent (M, 2**nbits, d/M) fp32 torch
    tensor on the CPU (as the reference returns it, ready for `torch.save`)."""
    if not torch.cuda.is_available():
        raise RuntimeError("million_b200.train_pq runs on the GPU (there is no CPU path)")
    X = torch.as_tensor(np.asarray(X) if not isinstance(X, torch.Tensor) else X, dtype=torch.float32)
    n, d = X.shape
    assert d > M and d % M == 0, "d must be divisible by M"      # pq_utils.py:598
    C, dm = 2 ** nbits, d // M
    if n < C:
        raise ValueError(f"{n} training vectors for {C} centroids")
    gen = torch.Generator().manual_seed(seed)
    if n > max_points_per_centroid * C:                             # faiss: "Sampling a subset of %d / %d for training"
        keep = torch.randperm(n, generator=gen)[:max_points_per_centroid * C]
        X = X[keep]
        n = X.shape[0]
    X = X.to(device).contiguous()
    # k distinct random training points per sub-space (one permutation per sub-space, like faiss's per-sub-quantizer training)
    cent = torch.empty(M, C, dm, dtype=torch.float32, device=device)
    for m in range(M):
        idx = torch.randperm(n, generator=gen)[:C].to(device)
        cent[m] = X[idx, m * dm:(m + 1) * dm]
    stats = []
    for it in range(niter):
        cent, codes, obj, counts = kmeans_step(X, cent)
        n_split = 0
        if int((counts == 0).sum()):
            cent, n_split = _split_empty(cent, counts, gen)
        stats.append({"iteration": it, "objective": obj, "empty_split": n_split})
        if verbose:
            print(f"  k-means iteration {it}: objective {obj:.6g}, {n_split} empty clusters re-seeded")
    out = cent.cpu()
    return (out, stats) if return_stats else out
