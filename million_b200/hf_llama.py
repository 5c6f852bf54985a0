"""The hook the reference patches into HuggingFace Llama attention (scripts/modeldb/models/modeling_llama.py:455-554
`attn_forward_custom_kernel`, :556-663 `attn_forward_paged_kernel`), for transformers >= 5 (`LlamaAttention`; the
`LlamaSdpaAttention` class the reference patches no longer exists).

    with patched_llama(model, cache, distort_recent=False):
        logits = model(input_ids).logits                      # q_len > 1 -> cache.prefill(...)
        logits = decode_step(model, next_token, position)     # q_len == 1 -> cache.decoding(...) / decoding_with_pages(...)

`cache` is a DynamicPQCache / PagedPQCache (or anything with the same prefill/decoding methods).  The KV state lives in
the PQ cache, not in HF's `past_key_values` (the reference does the same and patches prepare_inputs_for_generation,
modeling_llama.py:143-169; here `decode_step` passes explicit position ids instead).
"""
import contextlib
import types

import torch


def _as_tensor(x, like):
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(x).to(device=like.device, dtype=like.dtype)


def make_forward(cache, distort_recent=False):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        q = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        k = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        v = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        if q.size(2) > 1:                                              # modeling_llama.py:540-544
            attn = cache.prefill(q.contiguous(), k.contiguous(), v.contiguous(), self.layer_idx, distort_recent)
        else:                                                          # modeling_llama.py:545-547 / 650-654
            dec = getattr(cache, "decoding_with_pages", None) or cache.decoding
            attn = dec(q.contiguous(), k.contiguous(), v.contiguous(), self.layer_idx)
        attn = _as_tensor(attn, hidden_states)
        attn = attn.transpose(1, 2).reshape(*input_shape, -1).contiguous()
        return self.o_proj(attn), None

    return forward


@contextlib.contextmanager
def patched_llama(model, cache, distort_recent=False):
    """Swap the forward of every LlamaAttention module of `model` for the PQ-cache one; restored on exit."""
    from transformers.models.llama.modeling_llama import LlamaAttention
    fwd = make_forward(cache, distort_recent)
    mods = [m for m in model.modules() if isinstance(m, LlamaAttention)]
    saved = [m.forward for m in mods]
    try:
        for m in mods:
            m.forward = types.MethodType(fwd, m)
        yield model
    finally:
        for m, f in zip(mods, saved):
            m.forward = f


@torch.no_grad()
def decode_step(model, token_ids, position):
    """One decode step (token_ids (bs, 1)) at absolute position `position`; the PQ cache supplies the past."""
    pos = torch.full_like(token_ids, position)
    return model(input_ids=token_ids, position_ids=pos, use_cache=False).logits


class GraphDecoder:
    """Greedy decoding of a patched HF Llama with the WHOLE decode step — embeddings, every layer's projections / RoPE / PQ
    attention / MLP, the LM head — replayed as one CUDA graph.  The PQ cache supplies the past; its policy runs on
    the host between replays (million_b200.pq_utils._GraphStep), the attention launches read the window length from a device
    counter, the token and its position live in static device tensors.

        dec = GraphDecoder(model, cache)
        logits = model(prompt_ids).logits            # under patched_llama: prefill
        tok = logits[:, -1:].argmax(-1)
        for i in range(n):
            tok = dec.step(tok, T + i)[:, -1:].argmax(-1)
    """

    def __init__(self, model, cache, bs=None):
        from .pq_utils import _GraphStep
        self.model, self.cache = model, cache
        dev = next(model.parameters()).device
        bs = bs or cache.bs
        self.tok = torch.zeros(bs, 1, dtype=torch.long, device=dev)
        self.pos = torch.zeros(bs, 1, dtype=torch.long, device=dev)
        self.logits = None
        outer = self

        class _Step(_GraphStep):
            def _run(self_inner):
                with _patched_for_graph(outer.model, self_inner):
                    outer.logits = outer.model(input_ids=outer.tok, position_ids=outer.pos, use_cache=False).logits
                self_inner._advance()

        self._gs = _Step(cache, dev, cache.scalar_t)

    @property
    def captures(self):
        return self._gs.captures

    @torch.no_grad()
    def step(self, token_ids, position):
        self.tok.copy_(token_ids)
        self.pos.fill_(position)
        self._gs._step()
        return self.logits


@contextlib.contextmanager
def _patched_for_graph(model, gs):
    """Attention forward for graph capture: no host-side cache state is touched inside (that is _GraphStep's job)."""
    from transformers.models.llama.modeling_llama import LlamaAttention, apply_rotary_pos_emb

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        q = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        k = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        v = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out = torch.empty_like(q)
        gs._launch_layer(self.layer_idx, q, k, v, out)
        attn = out.transpose(1, 2).reshape(*input_shape, -1).contiguous()
        return self.o_proj(attn), None

    mods = [m for m in model.modules() if isinstance(m, LlamaAttention)]
    saved = [m.forward for m in mods]
    try:
        for m in mods:
            m.forward = types.MethodType(forward, m)
        yield model
    finally:
        for m, f in zip(mods, saved):
            m.forward = f

