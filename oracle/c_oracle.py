"""TEST INFRASTRUCTURE ONLY — ctypes front end of oracle/pq_oracle.c (liboracle.so, `make -C oracle`)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return lib().oracle_num_threads()


def pq_encode(X, cent):
    X = np.ascontiguousarray(X, dtype=np.float32)
    cent = np.ascontiguousarray(cent, dtype=np.float32)
    M, C, d_m = cent.shape
    n_vec = X.size // (M * d_m)
    codes = np.empty(X.shape[:-1] + (M,), dtype=np.uint8)
    lib().oracle_pq_encode(_p(X), _p(cent), _p(codes), ctypes.c_int64(n_vec), M, C, d_m)
    return codes


def pq_decode(codes, cent):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    cent = np.ascontiguousarray(cent, dtype=np.float32)
    M, C, d_m = cent.shape
    out = np.empty(codes.shape[:-1] + (M * d_m,), dtype=np.float32)
    lib().oracle_pq_decode(_p(codes), _p(cent), _p(out), ctypes.c_int64(codes.size // M), M, C, d_m)
    return out


def pq_decode_attn(q, kc, vc, kcent, vcent, kres, vres, r):
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    q, kcent, vcent, kres, vres = f(q), f(kcent), f(vcent), f(kres), f(vres)
    kc, vc = np.ascontiguousarray(kc, dtype=np.uint8), np.ascontiguousarray(vc, dtype=np.uint8)
    bs, nh = q.shape[0], q.shape[1]
    d = q.shape[-1]
    nh_k, nk, M = kc.shape[1], kc.shape[2], kc.shape[3]
    C = kcent.shape[1]
    Lt = kres.shape[2]
    out = np.empty((bs, nh, 1, d), dtype=np.float32)
    lib().oracle_pq_decode_attn(_p(q), _p(kc), _p(vc), _p(kcent), _p(vcent), _p(kres), _p(vres),
                                int(r), _p(out), bs, nh, nh_k, ctypes.c_int64(nk), d, M, C, Lt)
    return out
