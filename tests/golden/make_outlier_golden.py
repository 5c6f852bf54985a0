"""Fixed regression vectors for the outlier side store (extension: the reference has no outlier code, so these come from OUR
oracle, oracle/pq_oracle.py section A.6, not from the reference — they pin the definition against drift, nothing more).

    python tests/golden/make_outlier_golden.py      # rewrites tests/golden/outlier_golden.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pq_oracle as O  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    for name, (M, k_out) in {"m64k2": (64, 2), "m32k1": (32, 1), "m64k4": (64, 4)}.items():
        X = rng.standard_normal((1, 2, 96, 128)).astype(np.float16)
        X[..., [7, 100]] *= np.float16(12)
        X[0, 0, 0, :] = 1.0                      # ties everywhere: lowest dims win
        X[0, 0, 1, :] = 0.0
        C = rng.standard_normal((M, 256, 128 // M)).astype(np.float16).astype(np.float32)
        codes, idx, val = O.pq_encode_outliers(X, C, k_out)
        q = rng.standard_normal((1, 4, 1, 128)).astype(np.float16)
        V = rng.standard_normal((1, 2, 96, 128)).astype(np.float16)
        vcodes, vidx, vval = O.pq_encode_outliers(V, C, 1)
        kres = rng.standard_normal((1, 2, 128, 128)).astype(np.float16)
        vres = rng.standard_normal((1, 2, 128, 128)).astype(np.float16)
        attn = O.pq_decode_attn_outliers(q, codes, vcodes, C, C, kres, vres, 9, kout=(idx, val), vout=(vidx, vval))
        for k, v in dict(X=X, C=C, codes=codes, idx=idx, val=val, q=q, V=V, vcodes=vcodes, vidx=vidx, vval=vval, kres=kres, vres=vres,
                         recon=O.pq_decode_outliers(codes, C, idx, val), attn=attn).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "outlier_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith("m64k2")})


if __name__ == "__main__":
    main()
