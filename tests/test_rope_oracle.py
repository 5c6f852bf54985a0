"""CPU: the oracle's rotary-embedding restatement against transformers' own apply_rotary_pos_emb — the function the reference
calls in its attention forward (scripts/modeldb/models/modeling_llama.py:500-512) — bit for bit, fp16 and fp32."""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(3, 32, 8, 128), (1, 8, 8, 64), (2, 4, 1, 16)])
def test_oracle_rope_is_bit_identical_to_transformers(dtype, shape):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    bs, nh, nh_k, d = shape
    g = torch.Generator().manual_seed(5)
    q = torch.randn(bs, 1, nh, d, generator=g).to(dtype).transpose(1, 2)          # the views the attention forward holds at q_len = 1
    k = torch.randn(bs, 1, nh_k, d, generator=g).to(dtype).transpose(1, 2)
    pos = torch.randint(0, 100000, (bs, 1), generator=g).float()
    inv = 1.0 / (500000.0 ** (torch.arange(0, d, 2).float() / d))
    ang = torch.cat([pos[..., None] * inv, pos[..., None] * inv], dim=-1)           # (bs, 1, d), LlamaRotaryEmbedding
    cos, sin = ang.cos().to(dtype), ang.sin().to(dtype)
    qr, kr = apply_rotary_pos_emb(q, k, cos, sin)
    oq, ok = O.rope_qk(q.reshape(bs, nh, d).numpy(), k.reshape(bs, nh_k, d).numpy(), cos.reshape(bs, d).numpy(), sin.reshape(bs, d).numpy())
    assert np.array_equal(oq.view(np.uint16 if dtype == torch.float16 else np.uint32),
                          qr.reshape(bs, nh, d).numpy().view(np.uint16 if dtype == torch.float16 else np.uint32))
    assert np.array_equal(ok.view(np.uint16 if dtype == torch.float16 else np.uint32),
                          kr.reshape(bs, nh_k, d).numpy().view(np.uint16 if dtype == torch.float16 else np.uint32))
