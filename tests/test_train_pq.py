"""Codebook training on the GPU (million_b200/train_pq.py) and the .fvecs sample format (SURVEY 8(f)2;
scripts/utils/pq_utils.py:586-609, scripts/utils/fvecio.py:23-43, scripts/modeldb/main_pq.py:224-236)."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import pq_oracle as O


def test_fvecs_layout_matches_the_reference_format(tmp_path):
    """int32 d (little endian) + d float32 per vector, append mode by default (fvecio.py:23-43)."""
    from million_b200 import train_pq as T
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((5, 7), dtype=np.float32), rng.standard_normal((3, 7), dtype=np.float32)
    f = tmp_path / "key_0_0.fvecs"
    T.write_fvecs(f, a)
    T.write_fvecs(f, b)                      # appends, like the reference's sampling hook
    raw = open(f, "rb").read()
    assert len(raw) == 8 * (4 + 7 * 4)
    assert struct.unpack("<i", raw[:4])[0] == 7
    assert np.frombuffer(raw[4:32], dtype="<f4").tolist() == a[0].tolist()
    # byte-for-byte what the reference's per-vector loop writes
    want = b"".join(np.int32(7).tobytes() + v.astype(np.float32).tobytes() for v in np.concatenate([a, b]))
    assert raw == want
    back = T.read_fvecs(f)
    assert back.dtype == np.float32 and np.array_equal(back, np.concatenate([a, b]))
    T.write_fvecs(f, a, mode="wb")
    assert T.read_fvecs(f).shape == (5, 7)
    with open(tmp_path / "bad.fvecs", "wb") as g:
        g.write(np.int32(3).tobytes() + np.zeros(3, np.float32).tobytes() + np.int32(2).tobytes() + np.zeros(2, np.float32).tobytes() + b"\0" * 4)
    with pytest.raises(ValueError):
        T.read_fvecs(tmp_path / "bad.fvecs")


def test_centroid_file_round_trip(tmp_path):
    from million_b200 import train_pq as T
    cent = torch.randn(64, 256, 2)
    T.save_centroids(cent.half(), tmp_path / "key_cent_64_8.pq.pt")
    back = torch.load(tmp_path / "key_cent_64_8.pq.pt")       # what main_pq.py:262-270 does
    assert back.dtype == torch.float32 and back.shape == (64, 256, 2) and torch.equal(back, cent.half().float())


def test_oracle_kmeans_step_decreases_the_objective():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((2000, 8), dtype=np.float32)
    cent = X[rng.permutation(2000)[:16]].reshape(16, 4, 2).transpose(1, 0, 2).copy()     # (M=4, C=16, dm=2)
    objs = []
    for _ in range(6):
        cent, codes, obj, counts = O.kmeans_step(X, cent)
        objs.append(obj)
        assert counts.sum() == 4 * 2000
    assert all(b <= a * (1 + 1e-6) for a, b in zip(objs, objs[1:])) and objs[-1] < 0.9 * objs[0]


@pytest.mark.gpu
@pytest.mark.parametrize("M,nbits", [(64, 8), (32, 8), (64, 10)])
def test_gpu_kmeans_step_matches_the_oracle(M, nbits):
    """One Lloyd iteration on the GPU (assignment by the library's encoder) against the numpy restatement, from the same
    centroids: codes bit-exact, counts equal, centroids equal up to fp32 summation order."""
    from million_b200 import train_pq as T
    rng = np.random.default_rng(M + nbits)
    n, d, C = 6000, 128, 2 ** nbits
    X = rng.standard_normal((n, d), dtype=np.float32)
    cent = np.stack([X[rng.permutation(n)[:C], m * (d // M):(m + 1) * (d // M)] for m in range(M)]).astype(np.float32)
    new, codes, obj, counts = T.kmeans_step(torch.from_numpy(X).cuda(), torch.from_numpy(cent).cuda())
    onew, ocodes, oobj, ocounts = O.kmeans_step(X, cent)
    assert np.array_equal(codes.cpu().numpy(), ocodes)
    assert np.array_equal(counts.cpu().numpy(), ocounts)
    np.testing.assert_allclose(new.cpu().numpy(), onew, rtol=0, atol=2e-6)
    assert abs(obj - oobj) <= 1e-4 * oobj


@pytest.mark.gpu
def test_train_pq_end_to_end(tmp_path):
    """train_pq(X, M, nbits) as main_pq.py:216-236 uses it: samples from .fvecs, centroids saved as .pq.pt; the objective falls
    monotonically, no cluster stays empty, and the trained codebook quantizes the training set better than random points do."""
    from million_b200 import ops, train_pq as T
    rng = np.random.default_rng(5)
    n, d, M, nbits = 20000, 128, 64, 8
    basis = rng.standard_normal((16, d)).astype(np.float32)
    X = (rng.standard_normal((n, 16)).astype(np.float32) @ basis + 0.3 * rng.standard_normal((n, d)).astype(np.float32))
    T.write_fvecs(tmp_path / "key_sampled_64_8.fvecs", X, mode="wb")
    key = T.read_fvecs(tmp_path / "key_sampled_64_8.fvecs")
    cent, stats = T.train_pq(key, M, nbits, niter=12, return_stats=True)
    assert cent.shape == (M, 256, 2) and cent.dtype == torch.float32 and cent.device.type == "cpu"
    objs = [s["objective"] for s in stats]
    assert all(b <= a * (1 + 1e-5) for a, b in zip(objs[1:], objs[2:])), objs       # (iteration 0 measures the random init)
    assert objs[-1] < 0.7 * objs[0]
    T.save_centroids(cent, tmp_path / "key_cent_64_8.pq.pt")
    # quantization error of the trained codebook vs the initial one (random training points)
    Xg = torch.from_numpy(X).cuda().view(1, 1, n, d)
    codes = ops.pq_encode(Xg.half(), cent.cuda())
    rec = ops.pq_decode(codes, cent.cuda().half()).float()
    err = float(((rec - Xg.half().float()) ** 2).mean())
    assert err < 0.5 * float((Xg ** 2).mean())
    counts = torch.stack([torch.bincount(codes[0, 0, :, m].long(), minlength=256) for m in range(M)])
    assert int((counts == 0).sum()) <= 0.02 * counts.numel()
