"""GPU: compare with the REFERENCE's own CUDA kernels (flash_decoding_split/residual/reduce, Kernel.cuh), compiled for
sm_100a from /root/reference by oracle/build_ref.sh into oracle/_ref/libref_kernels.so (travels with the snapshot).
The reference accumulates in fp16 (core/Scalar.cuh:34-78), so it is a secondary comparator: both implementations must
sit around the fp32 oracle, ours much closer.  Skipped when the comparator library is absent."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_kernels.so")


def ref_kernel(q, kc, vc, kcent, vcent, kres, vres, r, Ns=32):
    lib = ctypes.CDLL(REF_SO)
    bs, nh, _, d = q.shape
    nh_k, nk, M = kc.shape[1], kc.shape[2], kc.shape[3]
    # Interface.cu:49-50
    lut = torch.matmul(q.reshape(bs, nh, 1, M, d // M).transpose(2, 3), kcent.transpose(1, 2)).contiguous()
    pout = torch.empty(bs, nh, Ns + 1, d, dtype=torch.float16, device="cuda")
    plse = torch.empty(bs, nh, Ns + 1, dtype=torch.float16, device="cuda")
    out = torch.empty(bs, nh, 1, d, dtype=torch.float16, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.ref_flash_decoding_f16u8_Lt128d128M64C256(Ns, p(lut), p(kc), p(vc), p(vcent), p(q), p(kres), p(vres), int(r), p(pout), p(plse),
                                                       p(out), bs, nh, nh_k, nk, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    return out


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_kernels.so not built (oracle/build_ref.sh)")
@pytest.mark.parametrize("nh,nh_k,nk,r", [(32, 8, 4096, 17), (32, 32, 2048, 128)])
def test_against_reference_cuda_kernel(nh, nh_k, nk, r):
    from million_b200 import ops
    inp = O.make_inputs(bs=1, nh=nh, nh_k=nh_k, nk=nk, seed=21)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
    ours = ops.pq_decode_attn(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], r)
    theirs = ref_kernel(t["q"], t["kc"], t["vc"], t["kcent"], t["vcent"], t["kres"], t["vres"], r)
    oracle = O.pq_decode_attn(inp["q"], inp["kc"], inp["vc"], inp["kcent"], inp["vcent"], inp["kres"], inp["vres"], r)
    e_ours = np.abs(ours.float().cpu().numpy() - oracle).max()
    e_theirs = np.abs(theirs.float().cpu().numpy() - oracle).max()
    print(f"max |err| vs fp32 oracle: ours {e_ours:.2e}, reference kernel (fp16 accumulation) {e_theirs:.2e}")
    assert e_ours <= 2e-3
    assert e_theirs <= 5e-2                       # the reference kernel itself is only this close to its own invariant
    np.testing.assert_allclose(ours.float().cpu().numpy(), theirs.float().cpu().numpy(), atol=5e-2)
