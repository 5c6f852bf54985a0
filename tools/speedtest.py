"""TTFT / TPOT harness — the counterpart of the reference's scripts/benchmarks/speedtest.py:36-116 for the B200 path
(SURVEY 8(f)4).  A random-init HuggingFace Llama with Llama-3.1-8B layer shapes (hidden 4096, 32 q / 8 kv heads, head_dim 128;
the layer count is a parameter: there are no checkpoints in this image and 32 random layers tell no more than a few) is run
with synthetic token ids (speedtest.py:32-34), greedy decoding, through

    * baseline : HF's own attention + DynamicCache (fp16 KV), `model.generate`-style loop
    * pq       : million_b200.hf_llama.patched_llama + DynamicPQCache (4-bit PQ, window 128)
    * pq_paged : the same with PagedPQCache
    * pq_graph : DynamicPQCache + the whole decode step (all layers, LM head) as ONE CUDA graph (hf_llama.GraphDecoder)

and reports, like the reference (`time_to_first_token`, `time_per_output_token`, milliseconds), the first-token latency
(prefill + quantisation of the whole prompt) and the mean latency of the following tokens, measured with CUDA events.

    python tools/speedtest.py [--layers 4] [--prefill 1024 4096 16384] [--decode 64] [--bs 1]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def build_model(layers, device):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers, num_attention_heads=32,
                      num_key_value_heads=8, head_dim=128, max_position_embeddings=131072, rope_theta=500000.0, attn_implementation="sdpa")
    torch.manual_seed(42)                      # scripts/modeldb/configs/default.json:19
    model = LlamaForCausalLM(cfg).to(device=device, dtype=torch.float16).eval()
    return model, cfg


@torch.no_grad()
def run_baseline(model, ids, n_decode):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_decode + 1)]
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    out = model(input_ids=ids, use_cache=True)
    past = out.past_key_values
    tok = out.logits[:, -1:].argmax(-1)
    ev[0].record()
    for i in range(n_decode):
        out = model(input_ids=tok, past_key_values=past, use_cache=True)
        past = out.past_key_values
        tok = out.logits[:, -1:].argmax(-1)
        ev[i + 1].record()
    torch.cuda.synchronize()
    return t0.elapsed_time(ev[0]), [ev[i].elapsed_time(ev[i + 1]) for i in range(n_decode)]


@torch.no_grad()
def run_pq_graph(model, cache, ids, n_decode):
    """PQ cache + the whole decode step as one CUDA graph (million_b200.hf_llama.GraphDecoder)."""
    from million_b200.hf_llama import GraphDecoder, patched_llama
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_decode + 1)]
    T = ids.shape[1]
    with patched_llama(model, cache):
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t0.record()
        tok = model(input_ids=ids, use_cache=False).logits[:, -1:].argmax(-1)
        ev[0].record()
        dec = GraphDecoder(model, cache)
        for i in range(n_decode):
            tok = dec.step(tok, T + i)[:, -1:].argmax(-1)
            ev[i + 1].record()
        torch.cuda.synchronize()
    return t0.elapsed_time(ev[0]), [ev[i].elapsed_time(ev[i + 1]) for i in range(n_decode)]


@torch.no_grad()
def run_pq(model, cache, ids, n_decode):
    from million_b200.hf_llama import decode_step, patched_llama
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_decode + 1)]
    T = ids.shape[1]
    with patched_llama(model, cache):
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t0.record()
        tok = model(input_ids=ids, use_cache=False).logits[:, -1:].argmax(-1)
        ev[0].record()
        for i in range(n_decode):
            tok = decode_step(model, tok, T + i)[:, -1:].argmax(-1)
            ev[i + 1].record()
        torch.cuda.synchronize()
    return t0.elapsed_time(ev[0]), [ev[i].elapsed_time(ev[i + 1]) for i in range(n_decode)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--prefill", type=int, nargs="+", default=[1024, 4096, 16384])
    ap.add_argument("--decode", type=int, default=64)
    ap.add_argument("--bs", type=int, default=1)
    ap.add_argument("--niter", type=int, default=2)
    a = ap.parse_args()
    from million_b200.paged_pq_utils import PagedPQCache
    from million_b200.pq_utils import DynamicPQCache, Singleton
    dev = torch.device("cuda", 0)
    model, cfg = build_model(a.layers, dev)
    g = torch.Generator(device=dev); g.manual_seed(42)
    kcent = torch.randn(64, 256, 2, device=dev, generator=g).half()      # main_pq.py:252-255: synthetic centroids
    vcent = torch.randn(64, 256, 2, device=dev, generator=g).half()
    results = []
    for T in a.prefill:
        ids = torch.randint(0, cfg.vocab_size, (a.bs, T), device=dev, generator=g)
        row = {"prefill_length": T, "decoding_length": a.decode, "bs": a.bs, "layers": a.layers}
        for name in ("baseline", "pq", "pq_paged", "pq_graph"):
            ttft, tpot = [], []
            for it in range(a.niter + 1):                # first iteration = warm-up (speedtest.py:92)
                if name == "baseline":
                    f, d = run_baseline(model, ids, a.decode)
                else:
                    Singleton.clear_instance()
                    cls = PagedPQCache if name == "pq_paged" else DynamicPQCache
                    cache = cls(bs=a.bs, nh=32, num_key_value_heads=8, M=64, layer_num=a.layers, d=128, scalar_t=torch.float16, device=dev)
                    cache.set_cent(kcent, vcent)
                    f, d = (run_pq_graph if name == "pq_graph" else run_pq)(model, cache, ids, a.decode)
                    del cache
                if it:
                    ttft.append(f); tpot.append(sorted(d[1:])[len(d[1:]) // 2])      # median: the graph variant's first step includes the capture
            row[name] = {"time_to_first_token": sum(ttft) / len(ttft), "time_per_output_token": sum(tpot) / len(tpot)}
        results.append(row)
        print(json.dumps(row), flush=True)
    print(json.dumps({"speedtest": results}))


if __name__ == "__main__":
    main()
