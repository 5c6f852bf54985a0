"""GPU: million_rope_qk (decode side) — the rotary embedding of a decode token's q and k in one launch — against the
oracle and against transformers' apply_rotary_pos_emb on the same device, bit for bit; and the patched HF Llama produces the same
logits with it as with the torch expression."""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(bs, nh, nh_k, d, dtype, seed=5):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(bs, 1, nh, d, generator=g).to(dtype).cuda().transpose(1, 2)
    k = torch.randn(bs, 1, nh_k, d, generator=g).to(dtype).cuda().transpose(1, 2)
    pos = torch.randint(0, 100000, (bs, 1), generator=g).float()
    inv = 1.0 / (500000.0 ** (torch.arange(0, d, 2).float() / d))
    ang = torch.cat([pos[..., None] * inv, pos[..., None] * inv], dim=-1)
    return q, k, ang.cos().to(dtype).cuda(), ang.sin().to(dtype).cuda()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(8, 32, 8, 128), (1, 8, 8, 64), (3, 4, 1, 16), (2, 32, 32, 256)])
def test_rope_qk_bit_identical_to_transformers_and_oracle(dtype, shape):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    from million_b200 import ops
    bs, nh, nh_k, d = shape
    q, k, cos, sin = _inputs(bs, nh, nh_k, d, dtype)
    qr, kr = apply_rotary_pos_emb(q, k, cos, sin)
    oq, ok = ops.rope_qk(q, k, cos, sin)
    assert oq.shape == q.shape and ok.shape == k.shape
    assert torch.equal(oq, qr) and torch.equal(ok, kr)
    if dtype != torch.bfloat16:                      # numpy has no bf16; the torch comparison above covers it
        nq, nk = O.rope_qk(q.reshape(bs, nh, d).cpu().numpy(), k.reshape(bs, nh_k, d).cpu().numpy(),
                           cos.reshape(bs, d).cpu().numpy(), sin.reshape(bs, d).cpu().numpy())
        assert np.array_equal(nq, oq.reshape(bs, nh, d).cpu().numpy()) and np.array_equal(nk, ok.reshape(bs, nh_k, d).cpu().numpy())


def test_rope_qk_in_place_and_errors():
    from million_b200 import ops
    from million_b200._lib import MillionError
    q, k, cos, sin = _inputs(2, 8, 2, 128, torch.float16)
    q, k = q.contiguous(), k.contiguous()
    ref_q, ref_k = ops.rope_qk(q, k, cos, sin)
    ops.rope_qk(q, k, cos, sin, q_out=q, k_out=k)                     # each thread reads its pair before writing it
    assert torch.equal(q, ref_q) and torch.equal(k, ref_k)
    with pytest.raises(ValueError):
        ops.rope_qk(torch.zeros(1, 2, 3, 8, device="cuda", dtype=torch.float16), torch.zeros(1, 2, 3, 8, device="cuda", dtype=torch.float16),
                    torch.zeros(1, 3, 8, device="cuda", dtype=torch.float16), torch.zeros(1, 3, 8, device="cuda", dtype=torch.float16))
    with pytest.raises(MillionError):
        x = torch.zeros(1, 1, 7, device="cuda", dtype=torch.float16)
        ops.rope_qk(x, x, torch.zeros(1, 7, device="cuda", dtype=torch.float16), torch.zeros(1, 7, device="cuda", dtype=torch.float16))


@torch.no_grad()
def test_patched_llama_logits_identical_with_fused_rope():
    from test_gpu_hf_llama import caches, tiny_llama
    from million_b200.hf_llama import GraphDecoder, decode_step, patched_llama
    model = tiny_llama().half().cuda()
    ids = torch.randint(1, 512, (1, 200), generator=torch.Generator().manual_seed(3)).cuda()
    runs = {}
    for name, fused, graph in (("torch", False, False), ("fused", True, False), ("torch_graph", False, True), ("fused_graph", True, True)):
        cache, _ = caches()
        with patched_llama(model, cache, fused_rope=fused):
            tok = model(input_ids=ids, use_cache=False).logits[:, -1:].argmax(-1)
            dec = GraphDecoder(model, cache, fused_rope=fused) if graph else None
            out = []
            for i in range(140):                                           # crosses a window flush
                lg = dec.step(tok, 200 + i) if graph else decode_step(model, tok, 200 + i)
                out.append(lg[:, -1].float().clone())
                tok = lg[:, -1:].argmax(-1)
        runs[name] = torch.stack(out)
    assert torch.equal(runs["torch"], runs["fused"])                    # same launches except the rotary embedding: bit-identical
    assert torch.equal(runs["torch_graph"], runs["fused_graph"])
