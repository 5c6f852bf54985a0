"""The hook the reference patches into HuggingFace Llama attention (scripts/modeldb/models/modeling_llama.py:455-554
`attn_forward_custom_kernel`, :556-663 `attn_forward_paged_kernel`), for transformers >= 5 (`LlamaAttention`; the
`LlamaSdpaAttention` class the reference patches no longer exists).

    with patched_llama(model, cache, distort_recent=False):
        logits = model(input_ids).logits                      # q_len > 1 -> cache.prefill(...)
        logits = decode_step(model, next_token, position)     # q_len == 1 -> cache.decoding(...) / decoding_with_pages(...)

`cache` is a DynamicPQCache / PagedPQCache (or anything with the same prefill/decoding methods).  The KV state lives in
the PQ cache, not in HF's `past_key_values` (the reference does the same and patches prepare_inputs_for_generation,
modeling_llama.py:143-169; here `decode_step` passes explicit position ids instead).  SURVEY 8(f)1, decode side: the rotary
embedding of the new token's q and k is one launch of the library (`million_rope_qk`), and its k goes straight into the attention
launch, which appends it to the window (`k_new`): per layer the producer side is projections -> 1 launch -> attention.
"""
import contextlib
import types

import torch


def _as_tensor(x, like):
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(x).to(device=like.device, dtype=like.dtype)


def _rope(q, k, cos, sin, fused):
    """apply_rotary_pos_emb of the reference's forward (modeling_llama.py:500-512).  For a decode token (q_len = 1) on the GPU the
    library's one-launch kernel (million_rope_qk, bit-identical to the torch expression) replaces its ~10 elementwise launches."""
    if fused and q.size(2) == 1 and q.is_cuda and q.dtype == k.dtype == cos.dtype == sin.dtype and cos.numel() == q.size(0) * q.size(3):
        from . import ops
        return ops.rope_qk(q, k, cos, sin)
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    return apply_rotary_pos_emb(q, k, cos, sin)


def make_forward(cache, distort_recent=False, fused_rope=True):

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        q = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        k = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        v = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = _rope(q, k, cos, sin, fused_rope)
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
def patched_llama(model, cache, distort_recent=False, fused_rope=True):
    """Swap the forward of every LlamaAttention module of `model` for the PQ-cache one; restored on exit."""
    from transformers.models.llama.modeling_llama import LlamaAttention
    fwd = make_forward(cache, distort_recent, fused_rope)
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

    def __init__(self, model, cache, bs=None, fused_rope=True):
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
                with _patched_for_graph(outer.model, self_inner, fused_rope):
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
def _patched_for_graph(model, gs, fused_rope=True):
    """Attention forward for graph capture: no host-side cache state is touched inside (that is _GraphStep's job)."""
    from transformers.models.llama.modeling_llama import LlamaAttention

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        q = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        k = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        v = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = _rope(q, k, cos, sin, fused_rope)
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

