"""ORACLE (test infrastructure, never on the product path): fp32 CPU restatement of the T5 arithmetic
that the reference's ``T5ForDocIDGeneration`` runs through HuggingFace ``T5Stack``.

Follows, for the decode path of ``t5_pretrainer/modeling/t5_generative_retriever.py``:
  * ``get_decoder_inputs_embeds``  :194-214  (start embed, per-position codebook gathers)
  * ``forward`` decode branch      :396-433  (decoder stack, optional d^-1/2 scale-up, LM head)
  * ``get_lm_logits``              :250-262  (per-position output table, shared or not)
and the HF 4.17 ``T5Stack`` block arithmetic the reference calls at :358-366 / :403-416, which is a
third-party dependency (transformers==4.17.0, reference requirements.txt:1) not vendored in the
reference tree; its published algorithm is restated here and pinned against
the installed transformers ``T5Stack`` in tests/test_oracle_t5.py.

Two decoders are provided: ``decoder_full_prefix`` re-runs the whole prefix every step exactly like the
reference does (its KV cache is never consumed, SURVEY.md §0 finding 4), and ``CachedDecoder`` is the
mathematically identical KV-cached form used to check the CUDA path at sizes the full-prefix form
cannot reach in seconds.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from ripor_b200.synthetic import T5Dims

W = Dict[str, torch.Tensor]


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    # T5LayerNorm: no mean subtraction, no bias, variance in fp32.
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    return weight * (x * torch.rsqrt(var + eps))


def relative_bucket(rel: torch.Tensor, bidirectional: bool, num_buckets: int, max_distance: int) -> torch.Tensor:
    """HF ``T5Attention._relative_position_bucket``; ``rel`` = key_pos - query_pos (int64)."""
    buckets = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (rel > 0).to(torch.long) * num_buckets
        rel = rel.abs()
    else:
        rel = -torch.minimum(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (
        torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.minimum(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rel, large)


def position_bias(rel_weight: torch.Tensor, q_len: int, k_len: int, bidirectional: bool, dims: T5Dims,
                  q_offset: int = 0) -> torch.Tensor:
    """[H, q_len, k_len] bias from the block-0 ``relative_attention_bias`` table [num_buckets, H]."""
    ctx = torch.arange(q_offset, q_offset + q_len, dtype=torch.long)[:, None]
    mem = torch.arange(k_len, dtype=torch.long)[None, :]
    b = relative_bucket(mem - ctx, bidirectional, dims.num_buckets, dims.max_distance)
    return rel_weight[b].permute(2, 0, 1).contiguous()


def _heads(x: torch.Tensor, H: int, dk: int) -> torch.Tensor:
    B, T, _ = x.shape
    return x.view(B, T, H, dk).transpose(1, 2)


def _attend(q, k, v, bias) -> torch.Tensor:
    # T5 does NOT scale by 1/sqrt(dk); softmax in fp32.
    scores = torch.matmul(q, k.transpose(-1, -2))
    if bias is not None:
        scores = scores + bias
    p = torch.softmax(scores.float(), dim=-1)
    o = torch.matmul(p, v)
    B, H, T, dk = o.shape
    return o.transpose(1, 2).reshape(B, T, H * dk)


def _lin(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """nn.Linear without bias: x @ W^T in fp32."""
    return x @ weight.t()


def encoder_forward(w: W, dims: T5Dims, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                    linear=_lin) -> torch.Tensor:
    """HF T5 encoder stack: [B,S] ids -> [B,S,d] (final layer norm applied).

    ``linear`` lets tools/precision_probe.py swap the projections for emulated tensor-core arithmetic."""
    H, dk = dims.num_heads, dims.d_kv
    x = w["shared.weight"][input_ids]
    B, S, _ = x.shape
    bias = position_bias(w["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], S, S, True, dims)
    neg = torch.finfo(torch.float32).min
    bias = bias[None] + ((1.0 - attention_mask[:, None, None, :].float()) * neg)
    for i in range(dims.num_layers):
        p = f"encoder.block.{i}.layer."
        h = rmsnorm(x, w[p + "0.layer_norm.weight"], dims.eps)
        q = _heads(linear(h, w[p + "0.SelfAttention.q.weight"]), H, dk)
        k = _heads(linear(h, w[p + "0.SelfAttention.k.weight"]), H, dk)
        v = _heads(linear(h, w[p + "0.SelfAttention.v.weight"]), H, dk)
        x = x + linear(_attend(q, k, v, bias), w[p + "0.SelfAttention.o.weight"])
        h = rmsnorm(x, w[p + "1.layer_norm.weight"], dims.eps)
        x = x + linear(torch.relu(linear(h, w[p + "1.DenseReluDense.wi.weight"])), w[p + "1.DenseReluDense.wo.weight"])
    return rmsnorm(x, w["encoder.final_layer_norm.weight"], dims.eps)


def decoder_input_embeds(w: W, dims: T5Dims, dec_ids: torch.Tensor) -> torch.Tensor:
    """reference get_decoder_inputs_embeds (:194-214): pos 0 = start embed, pos i>=1 = table[i-1][ids[:, i]]."""
    R, T = dec_ids.shape
    parts = [w["start_token_embed"].expand(R, 1, -1)]
    for i in range(1, T):
        parts.append(w[f"list_decoder_embeds.{i - 1}.weight"][dec_ids[:, i]].unsqueeze(1))
    return torch.cat(parts, dim=1)


def output_table(w: W, dims: T5Dims, t: int) -> torch.Tensor:
    """reference get_lm_logits (:254-260): output table of position t."""
    if dims.shared_output_input_embeds:
        return w[f"list_decoder_embeds.{t}.weight"]
    return w[f"list_output_embeds.{t}.weight"]


def decoder_full_prefix(w: W, dims: T5Dims, dec_ids: torch.Tensor, enc: torch.Tensor,
                        enc_mask: torch.Tensor) -> torch.Tensor:
    """Decoder stack over the whole prefix [R,T] with per-row encoder states [R,S,d]; returns [R,T,d]
    after the final layer norm and the optional scale-up (reference :403-428)."""
    H, dk = dims.num_heads, dims.d_kv
    x = decoder_input_embeds(w, dims, dec_ids)
    R, T, _ = x.shape
    neg = torch.finfo(torch.float32).min
    causal = torch.tril(torch.ones(T, T)).bool()
    self_bias = position_bias(w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"],
                              T, T, False, dims)
    self_bias = (self_bias + torch.where(causal, 0.0, neg)[None])[None]
    cross_bias = ((1.0 - enc_mask[:, None, None, :].float()) * neg)
    for i in range(dims.num_decoder_layers):
        p = f"decoder.block.{i}.layer."
        h = rmsnorm(x, w[p + "0.layer_norm.weight"], dims.eps)
        q = _heads(h @ w[p + "0.SelfAttention.q.weight"].t(), H, dk)
        k = _heads(h @ w[p + "0.SelfAttention.k.weight"].t(), H, dk)
        v = _heads(h @ w[p + "0.SelfAttention.v.weight"].t(), H, dk)
        x = x + _attend(q, k, v, self_bias) @ w[p + "0.SelfAttention.o.weight"].t()
        h = rmsnorm(x, w[p + "1.layer_norm.weight"], dims.eps)
        q = _heads(h @ w[p + "1.EncDecAttention.q.weight"].t(), H, dk)
        k = _heads(enc @ w[p + "1.EncDecAttention.k.weight"].t(), H, dk)
        v = _heads(enc @ w[p + "1.EncDecAttention.v.weight"].t(), H, dk)
        x = x + _attend(q, k, v, cross_bias) @ w[p + "1.EncDecAttention.o.weight"].t()
        h = rmsnorm(x, w[p + "2.layer_norm.weight"], dims.eps)
        x = x + torch.relu(h @ w[p + "2.DenseReluDense.wi.weight"].t()) @ w[p + "2.DenseReluDense.wo.weight"].t()
    x = rmsnorm(x, w["decoder.final_layer_norm.weight"], dims.eps)
    if dims.scaleup_output_hidden:
        x = x * (dims.d_model ** -0.5)
    return x


def lm_logits_list(w: W, dims: T5Dims, hidden: torch.Tensor) -> List[torch.Tensor]:
    """reference get_lm_logits: one [R,V] tensor per position (only the last one is consumed)."""
    return [hidden[:, i, :] @ output_table(w, dims, i).t() for i in range(hidden.shape[1])]


class CachedDecoder:
    """KV-cached decoder step, identical mathematics to ``decoder_full_prefix`` (causal attention).

    Cross K/V are projected once per query; self K/V are kept per row and reordered by ``reorder``.
    """

    def __init__(self, w: W, dims: T5Dims, enc: torch.Tensor, enc_mask: torch.Tensor, nb: int, linear=_lin):
        self.w, self.dims, self.nb, self.lin = w, dims, nb, linear
        H, dk = dims.num_heads, dims.d_kv
        self.B = enc.shape[0]
        neg = torch.finfo(torch.float32).min
        self.cross_bias = ((1.0 - enc_mask[:, None, None, :].float()) * neg)           # [B,1,1,S]
        self.ck, self.cv = [], []
        for i in range(dims.num_decoder_layers):
            p = f"decoder.block.{i}.layer.1.EncDecAttention."
            self.ck.append(_heads(linear(enc, w[p + "k.weight"]), H, dk))                   # [B,H,S,dk]
            self.cv.append(_heads(linear(enc, w[p + "v.weight"]), H, dk))
        self.sk: List[Optional[torch.Tensor]] = [None] * dims.num_decoder_layers          # [R,H,t,dk]
        self.sv: List[Optional[torch.Tensor]] = [None] * dims.num_decoder_layers
        self.rel = w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
        self.t = 0

    def reorder(self, beam_idx: torch.Tensor) -> None:
        for i in range(self.dims.num_decoder_layers):
            self.sk[i] = self.sk[i].index_select(0, beam_idx)
            self.sv[i] = self.sv[i].index_select(0, beam_idx)

    def step(self, last_tokens: Optional[torch.Tensor]) -> torch.Tensor:
        """Position ``self.t`` for all R = B*nb rows; ``last_tokens`` [R] (None at t=0). Returns logits [R,V]."""
        w, dims, t, linear = self.w, self.dims, self.t, self.lin
        H, dk, R = dims.num_heads, dims.d_kv, self.B * self.nb
        if t == 0:
            x = w["start_token_embed"].expand(R, 1, -1)
        else:
            x = w[f"list_decoder_embeds.{t - 1}.weight"][last_tokens].unsqueeze(1)
        bias = position_bias(self.rel, 1, t + 1, False, dims, q_offset=t)[None]          # [1,H,1,t+1]
        for i in range(dims.num_decoder_layers):
            p = f"decoder.block.{i}.layer."
            h = rmsnorm(x, w[p + "0.layer_norm.weight"], dims.eps)
            q = _heads(linear(h, w[p + "0.SelfAttention.q.weight"]), H, dk)
            k = _heads(linear(h, w[p + "0.SelfAttention.k.weight"]), H, dk)
            v = _heads(linear(h, w[p + "0.SelfAttention.v.weight"]), H, dk)
            self.sk[i] = k if t == 0 else torch.cat([self.sk[i], k], dim=2)
            self.sv[i] = v if t == 0 else torch.cat([self.sv[i], v], dim=2)
            x = x + linear(_attend(q, self.sk[i], self.sv[i], bias), w[p + "0.SelfAttention.o.weight"])
            h = rmsnorm(x, w[p + "1.layer_norm.weight"], dims.eps)
            q = _heads(linear(h, w[p + "1.EncDecAttention.q.weight"]), H, dk).view(self.B, self.nb, H, 1, dk)
            sc = torch.einsum("bnhqd,bhsd->bnhqs", q, self.ck[i]) + self.cross_bias[:, None]
            pr = torch.softmax(sc.float(), dim=-1)
            o = torch.einsum("bnhqs,bhsd->bnhqd", pr, self.cv[i]).reshape(R, H, 1, dk)
            x = x + linear(o.transpose(1, 2).reshape(R, 1, H * dk), w[p + "1.EncDecAttention.o.weight"])
            h = rmsnorm(x, w[p + "2.layer_norm.weight"], dims.eps)
            x = x + linear(torch.relu(linear(h, w[p + "2.DenseReluDense.wi.weight"])), w[p + "2.DenseReluDense.wo.weight"])
        x = rmsnorm(x, w["decoder.final_layer_norm.weight"], dims.eps)
        if dims.scaleup_output_hidden:
            x = x * (dims.d_model ** -0.5)
        self.t += 1
        return linear(x[:, 0, :], output_table(w, dims, t))
