"""Generate golden fixtures by running the REFERENCE's own code (read-only, /root/reference) in the
build container.  The fixtures (`*.npz`, small) are committed; this script is what made them.

    python tests/golden/make_golden.py

The reference's `scripts/utils/pq_utils.py` imports `pykeops` (not installed here, pinned 2.2.3 in
requirements.txt:80).  We put a *dense torch stand-in* for `pykeops.torch.LazyTensor` into
sys.modules: it implements only `-`, `** 2`, `.sum(-1)` and `.argmin(dim=2)` and evaluates them
eagerly in the dtype it is given (fp32, per pq_utils.py:483-484), so the reference's
`sa_encode_4d_keops` runs UNCHANGED (its reshapes/permutes/casts are the reference's; only the
KeOps reduction is substituted by its published semantics).  `sa_encode_4d`, `sa_decode_4d`,
`l2Ns`, `nbits2dtype` run unchanged with no substitution at all.

Nothing under tests/ or the product reads /root/reference at run time; only this script does.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MILLION_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


class _Lazy:
    def __init__(self, t):
        self.t = t

    def __sub__(self, o):
        return _Lazy(self.t - o.t)

    def __pow__(self, p):
        assert p == 2
        return _Lazy(self.t * self.t)

    def sum(self, dim):
        return _Lazy(self.t.sum(dim))

    def argmin(self, dim):
        return self.t.argmin(dim=dim)


def load_reference():
    mod = types.ModuleType("pykeops")
    sub = types.ModuleType("pykeops.torch")
    sub.LazyTensor = _Lazy
    mod.torch = sub
    sys.modules["pykeops"] = mod
    sys.modules["pykeops.torch"] = sub
    sys.path.insert(0, REF)
    import scripts.utils.pq_utils as ref  # noqa
    return ref


def main():
    ref = load_reference()
    torch.manual_seed(42)                     # scripts/modeldb/configs/default.json:19
    out = {}

    # --- l2Ns / nbits2dtype tables
    ls = [0, 1, 64, 65, 128, 129, 256, 257, 2048, 2049, 32768, 131072]
    out["l2Ns_l"] = np.array(ls)
    out["l2Ns_v"] = np.array([ref.l2Ns(l) for l in ls])
    nb = [1, 4, 8, 9, 12, 16, 17, 32]
    out["nbits"] = np.array(nb)
    out["nbits_itemsize"] = np.array([torch.empty(0, dtype=ref.nbits2dtype(b)).element_size() for b in nb])

    # --- encode / decode, several shapes (bs, nh_k, n, d, M, C)
    cases = [(1, 2, 37, 128, 64, 256), (2, 3, 16, 128, 32, 256), (1, 1, 130, 64, 16, 128),
             (1, 2, 8, 128, 16, 256), (1, 1, 0, 128, 64, 256)]
    for i, (bs, nh, n, d, M, C) in enumerate(cases):
        X = torch.randn(bs, nh, n, d).half()
        cent = torch.randn(M, C, d // M).half()
        if n:
            codes_keops = ref.sa_encode_4d_keops(X, cent)               # pq_utils.py:451
            codes_cdist = ref.sa_encode_4d(X.float(), cent.float())     # pq_utils.py:410
        else:
            codes_keops = torch.zeros(bs, nh, 0, M, dtype=torch.uint8)
            codes_cdist = codes_keops
        dec = ref.sa_decode_4d(codes_keops, cent)                       # pq_utils.py:501
        out[f"enc{i}_shape"] = np.array([bs, nh, n, d, M, C])
        out[f"enc{i}_X"] = X.numpy()
        out[f"enc{i}_cent"] = cent.numpy()
        out[f"enc{i}_codes_keops"] = codes_keops.numpy()
        out[f"enc{i}_codes_cdist"] = codes_cdist.numpy()
        out[f"enc{i}_decoded"] = dec.numpy()
    out["n_enc"] = np.array(len(cases))

    # --- an exact-tie case: duplicated centroids -> lowest index must win (torch.argmin semantics)
    cent = torch.randn(4, 8, 2).half()
    cent[:, 5] = cent[:, 2]
    X = cent[:, 2].reshape(1, 1, 1, 8).clone()
    out["tie_X"], out["tie_cent"] = X.numpy(), cent.numpy()
    out["tie_codes"] = ref.sa_encode_4d_keops(X, cent).numpy()

    # --- nbits > 8 -> uint16 codes (pq_utils.py:542-552)
    cent = torch.randn(8, 512, 16).half()
    X = torch.randn(1, 1, 9, 128).half()
    out["wide_X"], out["wide_cent"] = X.numpy(), cent.numpy()
    out["wide_codes"] = ref.sa_encode_4d_keops(X, cent, target_dtype=torch.int32).numpy()

    # --- decode attention: the invariant the reference writes down for its kernel
    #     (pq_utils.py:360-368; test_kernel.py:86-90 without the stray is_causal), fp32 on CPU,
    #     built ONLY from reference functions + torch SDPA.
    att = [(1, 8, 2, 300, 17, 128, 64, 256), (2, 4, 4, 65, 128, 128, 32, 256), (1, 4, 1, 0, 5, 64, 16, 128),
           (1, 32, 8, 257, 1, 128, 64, 256)]
    for i, (bs, nh, nh_k, nk, r, d, M, C) in enumerate(att):
        q = torch.randn(bs, nh, 1, d).half()
        kc = torch.randint(0, C, (bs, nh_k, nk, M), dtype=torch.uint8)
        vc = torch.randint(0, C, (bs, nh_k, nk, M), dtype=torch.uint8)
        kcent = torch.randn(M, C, d // M).half()
        vcent = torch.randn(M, C, d // M).half()
        kres = torch.randn(bs, nh_k, d, d).half()
        vres = torch.randn(bs, nh_k, d, d).half()
        K = torch.cat([ref.sa_decode_4d(kc, kcent.float()), kres[:, :, :r].float()], dim=2)
        V = torch.cat([ref.sa_decode_4d(vc, vcent.float()), vres[:, :, :r].float()], dim=2)
        G = nh // nh_k
        o = torch.nn.functional.scaled_dot_product_attention(
            q.float(), K.repeat_interleave(G, 1), V.repeat_interleave(G, 1))
        for k_, v_ in dict(q=q, kc=kc, vc=vc, kcent=kcent, vcent=vcent, kres=kres, vres=vres, out=o).items():
            out[f"att{i}_{k_}"] = v_.numpy()
        out[f"att{i}_shape"] = np.array([bs, nh, nh_k, nk, r, d, M, C])
    out["n_att"] = np.array(len(att))

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "reference_golden.npz"))
    print(f"wrote reference_golden.npz ({sz/1024:.0f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
