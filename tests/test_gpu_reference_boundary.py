"""GPU: the zero-edit boundary proof.  The REFERENCE's own scripts/utils/pq_utils.py — unmodified, staged byte-for-byte by
oracle/stage_ref.py into the git-ignored oracle/_ref/ref_py/ — drives ITS DynamicPQCache.prefill / decoding (pq_utils.py:222-327)
and ITS KernelRegistry (pq_utils.py:32-96), and the only native code underneath is ours: `million_b200.bindings.install()` makes
the `__import__('bindings')` of pq_utils.py:65 resolve to our closures over the C ABI.

pykeops (pinned 2.2.3, not installed in this image) is replaced by the same dense torch stand-in tests/golden/make_golden.py
uses for LazyTensor; a second variant swaps in OUR encoder/reconstruct under the reference's cache class.  Both are checked
against the oracle step by step.  Skipped when the staged files are absent."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "oracle", "_ref", "ref_py")
ATOL, RTOL = 2e-3, 1e-2


class _Lazy:
    """Dense stand-in for pykeops.torch.LazyTensor: exactly the four operations pq_utils.py:487-494 uses."""

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


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(os.path.join(REF_PY, "scripts", "utils", "pq_utils.py")):
        pytest.skip("reference modules not staged (python oracle/stage_ref.py in the build container)")
    mod, sub = types.ModuleType("pykeops"), types.ModuleType("pykeops.torch")
    sub.LazyTensor = _Lazy
    mod.torch = sub
    saved = {k: sys.modules.get(k) for k in ("pykeops", "pykeops.torch", "bindings", "scripts", "scripts.utils")}
    sys.modules["pykeops"], sys.modules["pykeops.torch"] = mod, sub
    sys.path.insert(0, REF_PY)
    for k in [k for k in sys.modules if k == "scripts" or k.startswith("scripts.")]:
        del sys.modules[k]
    from million_b200 import bindings
    bindings.install()
    import scripts.utils.pq_utils as ref_mod
    assert os.path.realpath(ref_mod.__file__).startswith(os.path.realpath(REF_PY))
    yield ref_mod
    sys.path.remove(REF_PY)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


@pytest.mark.parametrize("our_codec", [False, True])
def test_reference_cache_class_runs_on_our_kernels(ref, our_codec, monkeypatch):
    from million_b200 import pq_utils as ours
    if our_codec:
        # the reference's cache policy over OUR encoder / reconstruct too (same names, same signatures: pq_utils.py:451, 501)
        monkeypatch.setattr(ref, "sa_encode_4d_keops", ours.sa_encode_4d_keops)
        monkeypatch.setattr(ref, "sa_decode_4d", ours.sa_decode_4d)
    ref.Singleton.clear_instance()                 # scripts/utils/Singleton.py:24-27
    bs, nh, nh_k, d = 1, 32, 8, 128                 # the reference's registry hard-wires bs = 1 (pq_utils.py:56)
    kw = dict(bs=bs, nh=nh, num_key_value_heads=nh_k, M=64, layer_num=1, d=d)
    cache = ref.DynamicPQCache(scalar_t=torch.float16, **kw)
    cache.init_cache()
    rng = np.random.default_rng(12)
    f = lambda *s: rng.standard_normal(s, dtype=np.float32).astype(np.float16)
    kc, vc = f(64, 256, 2), f(64, 256, 2)
    cache.set_cent(torch.from_numpy(kc).cuda(), torch.from_numpy(vc).cuda())
    oracle = O.DynamicPQCacheOracle(**kw)
    oracle.set_cent(kc, vc)
    T0 = 300
    q, k, v = f(bs, nh, T0, d), f(bs, nh_k, T0, d), f(bs, nh_k, T0, d)
    out = cache.prefill(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), 0)
    np.testing.assert_allclose(out.float().cpu().numpy(), oracle.prefill(q, k, v, 0), atol=ATOL, rtol=RTOL)
    for step in range(140):                        # crosses the flush at 128 (pq_utils.py:288-301)
        q, k, v = f(bs, nh, 1, d), f(bs, nh_k, 1, d), f(bs, nh_k, 1, d)
        out = cache.decoding(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), 0)
        want = oracle.decoding(q, k, v, 0)
        np.testing.assert_allclose(out.float().cpu().numpy(), want, atol=ATOL, rtol=RTOL, err_msg=f"step {step}")
    assert cache.seen_tokens == oracle.seen_tokens and cache.residualed_tokens == oracle.residualed_tokens
    assert np.array_equal(cache.key_cache[0].cpu().numpy(), oracle.key_cache[0])
    assert np.array_equal(cache.value_cache[0].cpu().numpy(), oracle.value_cache[0])
    # the kernel really came from our module, by the reference's own name grammar
    import bindings as b
    assert b is sys.modules["million_b200.bindings"]
    assert 32 in cache.registery.kernels or 16 in cache.registery.kernels
