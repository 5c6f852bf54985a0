"""GPU: BASELINE config 0 end to end — a tiny random-init HF Llama (2 layers, 8 heads, head_dim 128) with its attention
replaced by the PQ cache (million_b200.hf_llama, the hook of modeling_llama.py:455-663):
  * synthetic-token perplexity with the PQ path in prefill (distort_recent=True, pq_utils.py:243-260) within 0.5 % of the
    oracle's PQ path (north_star);
  * prefill + decode over the PQ codes (window flushes included) tracks the oracle's cache step by step."""
import numpy as np
import pytest
import torch

from oracle import pq_oracle as O

pytestmark = pytest.mark.gpu


def tiny_llama(seed=0):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    cfg = LlamaConfig(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=8,
                      head_dim=128, vocab_size=512, max_position_embeddings=8192, rms_norm_eps=1e-5, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).eval()


def caches(bs=1):
    from million_b200.pq_utils import DynamicPQCache, Singleton
    Singleton.clear_instance()
    kw = dict(bs=bs, nh=8, num_key_value_heads=8, M=64, layer_num=2, d=128)
    g = torch.Generator().manual_seed(7)
    kc, vc = torch.randn(64, 256, 2, generator=g).half(), torch.randn(64, 256, 2, generator=g).half()
    gpu = DynamicPQCache(scalar_t=torch.float16, **kw)
    gpu.set_cent(kc.cuda(), vc.cuda())
    ora = O.DynamicPQCacheOracle(**kw)
    ora.set_cent(kc.numpy(), vc.numpy())
    return gpu, ora


@torch.no_grad()
def test_synthetic_token_perplexity_matches_oracle_pq_path():
    from million_b200.hf_llama import patched_llama
    model = tiny_llama()
    gpu_cache, ora_cache = caches()
    ids = torch.randint(1, 512, (1, 384), generator=torch.Generator().manual_seed(1))
    m16 = tiny_llama().half().cuda()
    with patched_llama(m16, gpu_cache, distort_recent=True):
        nll_gpu = m16(input_ids=ids.cuda(), labels=ids.cuda(), use_cache=False).loss.float().item()
    with patched_llama(model, ora_cache, distort_recent=True):
        nll_ref = model(input_ids=ids, labels=ids, use_cache=False).loss.item()
    ppl_gpu, ppl_ref = float(np.exp(nll_gpu)), float(np.exp(nll_ref))
    print(f"synthetic-token PPL: GPU PQ path {ppl_gpu:.3f}, oracle PQ path {ppl_ref:.3f}")
    assert abs(ppl_gpu / ppl_ref - 1) < 5e-3
    # the PQ path really is in the loop: the codes of the prefill are in the cache.  The two models run in fp16 (GPU) and fp32
    # (CPU), so their K activations differ by fp16 rounding and ~2 % of the codes sit on the other side of a cell boundary;
    # bit-exactness of the encoder on IDENTICAL inputs is tests/test_gpu_parity.py's job.
    assert gpu_cache.key_cache[0].shape == (1, 8, 384, 64)
    same = (gpu_cache.key_cache[0].cpu().numpy() == ora_cache.key_cache[0]).mean()
    assert same > 0.95


@torch.no_grad()
def test_prefill_then_decode_tracks_oracle_cache():
    from million_b200.hf_llama import decode_step, patched_llama
    m32 = tiny_llama()
    m16 = tiny_llama().half().cuda()
    gpu_cache, ora_cache = caches()
    T0, steps = 100, 140                      # the 128-token window fills once -> one flush
    ids = torch.randint(1, 512, (1, T0 + steps), generator=torch.Generator().manual_seed(2))
    with patched_llama(m16, gpu_cache), patched_llama(m32, ora_cache):
        lg = m16(input_ids=ids[:, :T0].cuda(), use_cache=False).logits[:, -1].float().cpu()
        lr = m32(input_ids=ids[:, :T0], use_cache=False).logits[:, -1]
        assert (lg - lr).abs().max() < 5e-2
        agree = 0
        for s in range(steps):
            tok = ids[:, T0 + s:T0 + s + 1]
            lg = decode_step(m16, tok.cuda(), T0 + s)[:, -1].float().cpu()
            lr = decode_step(m32, tok, T0 + s)[:, -1]
            assert (lg - lr).abs().max() < 8e-2, f"step {s}"
            agree += int(lg.argmax() == lr.argmax())
        assert agree >= steps - 3
    assert gpu_cache.seen_tokens == ora_cache.seen_tokens == [T0 + steps] * 2
    assert gpu_cache.key_cache[0].shape[2] == T0 + 128 and gpu_cache.residualed_tokens[0] == steps - 128


@torch.no_grad()
@pytest.mark.parametrize("paged", [False, True])
def test_whole_model_graph_decoder_matches_eager_decode(paged):
    """GraphDecoder: embeddings -> layers (projections, RoPE, PQ attention, MLP) -> LM head of one decode step replayed as ONE
    CUDA graph; the logits must be those of the eager decode_step path, across a window flush (re-capture)."""
    from million_b200.hf_llama import GraphDecoder, decode_step, patched_llama
    from million_b200.paged_pq_utils import PagedPQCache
    from million_b200.pq_utils import DynamicPQCache, Singleton
    m16 = tiny_llama().half().cuda()
    T0, steps = 128, 150
    ids = torch.randint(1, 512, (1, T0 + steps), generator=torch.Generator().manual_seed(3)).cuda()
    g = torch.Generator().manual_seed(7)
    kc, vc = torch.randn(64, 256, 2, generator=g).half().cuda(), torch.randn(64, 256, 2, generator=g).half().cuda()
    runs = []
    for mode in ("eager", "graph"):
        Singleton.clear_instance()
        cls = PagedPQCache if paged else DynamicPQCache
        cache = cls(bs=1, nh=8, num_key_value_heads=8, M=64, layer_num=2, d=128, scalar_t=torch.float16)
        cache.set_cent(kc, vc)
        out = []
        with patched_llama(m16, cache):
            m16(input_ids=ids[:, :T0], use_cache=False)
            dec = GraphDecoder(m16, cache) if mode == "graph" else None
            for s in range(steps):
                tok = ids[:, T0 + s:T0 + s + 1]
                lg = dec.step(tok, T0 + s) if dec else decode_step(m16, tok, T0 + s)
                out.append(lg[:, -1].float().clone())
        runs.append(torch.stack(out))
        if dec:
            assert 2 <= dec.captures <= 5
        assert cache.seen_tokens == [T0 + steps] * 2
    torch.cuda.synchronize()
    # same kernels, same inputs; cuBLAS may pick different GEMM algorithms inside and outside a graph, hence a tolerance
    assert (runs[0] - runs[1]).abs().max() < 2e-2
    assert (runs[0].argmax(-1) == runs[1].argmax(-1)).float().mean() > 0.98
