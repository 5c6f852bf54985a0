"""GPU: randomized differential test of the fast decode-attention kernels (attn_fast.cu, attn_fast_dm4.cu) against the
independently written all-shapes kernel (attn_generic.cu) and, for a subset, the oracle — random shapes, group sizes, value
layouts, outlier record counts, window lengths, fused append / device-resident window length / PDL, ragged token counts.
Seeded: failures reproduce from the printed case."""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu
ATOL, RTOL = 2e-3, 1e-2


def _case(rng):
    Mm = int(rng.choice([64, 64, 32]))
    G = int(rng.choice([1, 2, 4, 4, 8]))
    nh_k = int(rng.choice([1, 2, 3, 8]))
    bs = int(rng.choice([1, 1, 2, 3, 5]))
    nk = int(rng.choice([0, 1, 31, 64, 65, 127, 500, 1023, 2049, 4100, 9000]))
    r = int(rng.choice([0, 1, 2, 17, 64, 127, 128])) if nk else int(rng.choice([1, 5, 128]))
    v_layout = int(rng.choice([0, 0, 1, 2]))
    k_out = int(rng.choice([0, 0, 1, 2, 3, 4]))
    v_out = int(rng.choice([0, 0, 0, 1, 2, 4]))
    fused = bool(rng.integers(0, 2)) and r >= 1
    use_rdev = fused and bool(rng.integers(0, 2))
    return dict(Mm=Mm, G=G, nh_k=nh_k, bs=bs, nk=nk, r=r, v_layout=v_layout, k_out=k_out, v_out=v_out, fused=fused, use_rdev=use_rdev,
                pdl=bool(rng.integers(0, 2)), bf16=bool(rng.integers(0, 4) == 0), n_splits=int(rng.choice([0, 0, 0, 1, 3, 7])))


@pytest.mark.parametrize("seed", range(20))
def test_fast_kernels_agree_with_the_all_shapes_kernel(seed):
    from million_b200 import _lib as L, ops
    rng = np.random.default_rng(1000 + seed)
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    checked = 0
    for it in range(40):
        c = _case(rng)
        Mm, G, nh_k, bs, nk, r = c["Mm"], c["G"], c["nh_k"], c["bs"], c["nk"], c["r"]
        nh, d, dt = G * nh_k, 128, (torch.bfloat16 if c["bf16"] else torch.float16)
        rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).to(dt)
        kcent, vcent = rnd(Mm, 256, d // Mm), rnd(Mm, 256, d // Mm)
        q, kres, vres = rnd(bs, nh, 1, d), rnd(bs, nh_k, 128, d), rnd(bs, nh_k, 128, d)
        cap = max(nk, 1)
        kc = torch.randint(0, 256, (bs, nh_k, cap, Mm), dtype=torch.uint8, device="cuda", generator=g)
        vc = torch.randint(0, 256, (bs, nh_k, cap, Mm), dtype=torch.uint8, device="cuda", generator=g)
        kw = {}
        if c["v_layout"] == 1:
            ld = (cap + 127) // 128 * 128
            vca = torch.zeros(bs, nh_k, Mm, ld, dtype=torch.uint8, device="cuda")
            vca[..., :cap] = vc.transpose(2, 3)
        elif c["v_layout"] == 2:
            pool, table = O.build_page_pool(vc.cpu().numpy(), 64)
            vca, kw = torch.from_numpy(pool).cuda(), dict(v_page_ids=torch.from_numpy(table).cuda(), page_size=64)
        else:
            vca = vc
        mk = lambda n: (torch.randint(0, d, (bs, nh_k, cap, n), dtype=torch.uint8, device="cuda", generator=g), (0.3 * torch.randn(bs, nh_k, cap, n, device="cuda", generator=g)).to(dt))
        ko = mk(c["k_out"]) if c["k_out"] and nk else None
        vo = mk(c["v_out"]) if c["v_out"] and nk else None
        common = dict(nk=nk, v_layout=c["v_layout"], k_outliers=ko, v_outliers=vo, n_splits=c["n_splits"], **kw)
        # reference: all-shapes kernel on a window that already holds the new token
        k_new, v_new = rnd(bs, nh_k, 1, d), rnd(bs, nh_k, 1, d)
        kw_ref, vw_ref = kres.clone(), vres.clone()
        if c["fused"]:
            ops.window_append(kw_ref, vw_ref, k_new, v_new, r - 1)
        want = ops.pq_decode_attn(q, kc, vca, kcent, vcent, kw_ref, vw_ref, r, impl=L.IMPL_GENERIC, **common)
        kw_t, vw_t = kres.clone(), vres.clone()
        extra = {}
        if c["fused"]:
            extra = dict(k_new=k_new, v_new=v_new)
            if c["use_rdev"]:
                extra["r_dev"] = torch.tensor([r - 1], dtype=torch.int32, device="cuda")
        try:
            got = ops.pq_decode_attn(q, kc, vca, kcent, vcent, kw_t, vw_t, 1 if c["use_rdev"] else r, impl=L.IMPL_FAST, pdl=c["pdl"], **common, **extra)
        except L.MillionError as e:
            assert e.status == L.MILLION_ERR_UNSUPPORTED, (c, str(e))
            continue
        checked += 1
        a, b = got.float().cpu().numpy(), want.float().cpu().numpy()
        if nk + r == 0:
            assert not a.any(), c
            continue
        np.testing.assert_allclose(a, b, atol=ATOL, rtol=RTOL, err_msg=f"seed {seed} case {it}: {c}")
        if c["fused"]:
            assert torch.equal(kw_t, kw_ref) and torch.equal(vw_t, vw_ref), f"window after the fused append, seed {seed} case {it}: {c}"
    assert checked >= 25
