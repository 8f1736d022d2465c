"""Seeded synthetic inputs for the constrained-beam-search retrieval path.

There are no T5 checkpoints, MS MARCO files or tokenizer models on the build or GPU boxes, so every
test and bench line runs on synthetic inputs of the reference's shapes:

* model weights keyed exactly like an HF ``save_pretrained`` state dict of the reference's
  ``T5ForDocIDGeneration`` (reference ``t5_pretrainer/modeling/t5_generative_retriever.py:85-112``),
* RQ code matrices ``codes[N, L]`` standing in for ``docid_to_smtid.json``
  (reference ``t5_pretrainer/evaluate.py:400-446``),
* padded query token batches like ``CollectionDataWithDocIDLoader`` emits
  (reference ``t5_pretrainer/dataset/dataloader.py:62-79``).

Everything here is numpy / CPU torch only; the same seeds give the same tensors on every box that
runs the same image, which is what lets the CPU oracle and the CUDA path see identical inputs.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import numpy as np
import torch

MODEL_SEED = 1234
CODEBOOK_SEED = 1235
TRIE_SEED = 2024
QUERY_SEED = 77


@dataclasses.dataclass
class T5Dims:
    """Dimensions of the T5 stack behind ``T5forDocIDConfig`` (reference t5_generative_retriever.py:45-67)."""

    d_model: int = 768
    num_heads: int = 12
    d_kv: int = 64
    d_ff: int = 3072
    num_layers: int = 12            # encoder blocks
    num_decoder_layers: int = 12
    vocab_size: int = 32128         # encoder token vocabulary (shared.weight)
    num_buckets: int = 32           # relative_attention_num_buckets
    max_distance: int = 128         # relative_attention_max_distance
    eps: float = 1e-6               # layer_norm_epsilon
    decoder_vocab_size: int = 256   # V: codebook size, uniform across positions (evaluate.py:433-436)
    docid_len: int = 32             # L = len(decoder_vocab_sizes)
    shared_output_input_embeds: bool = False
    scaleup_output_hidden: bool = False

    @property
    def inner(self) -> int:
        return self.num_heads * self.d_kv

    @staticmethod
    def t5_base(**kw) -> "T5Dims":
        return T5Dims(**kw)

    @staticmethod
    def t5_large(**kw) -> "T5Dims":
        return T5Dims(d_model=1024, num_heads=16, d_ff=4096, num_layers=24, num_decoder_layers=24, **kw)

    @staticmethod
    def tiny(**kw) -> "T5Dims":
        base = dict(d_model=128, num_heads=2, d_ff=256, num_layers=2, num_decoder_layers=2,
                    vocab_size=512, decoder_vocab_size=16, docid_len=4)
        base.update(kw)
        return T5Dims(**base)


def make_weights(dims: T5Dims, seed: int = MODEL_SEED, codebook_seed: int = CODEBOOK_SEED,
                 logit_std: float = 3.0, start_token_embed: Optional[np.ndarray] = None
                 ) -> Dict[str, torch.Tensor]:
    """Random fp32 weights under the reference's state-dict keys.

    Scales follow the T5 initialiser (q: (d*dk)^-1/2, k/v/wi: d^-1/2, o: inner^-1/2, wo: dff^-1/2),
    widened x2 on q so attention is not uniform; layer norms are 1 + 0.1*N(0,1). The output codebooks
    are N(0, (logit_std/sqrt(d))^2) so logits after the final RMSNorm are ~N(0, logit_std^2).
    """
    g = torch.Generator().manual_seed(seed)
    d, inner, dff = dims.d_model, dims.inner, dims.d_ff

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    w: Dict[str, torch.Tensor] = {}
    w["shared.weight"] = rn(dims.vocab_size, d)
    for side, nl in (("encoder", dims.num_layers), ("decoder", dims.num_decoder_layers)):
        for i in range(nl):
            p = f"{side}.block.{i}.layer."
            w[p + "0.SelfAttention.q.weight"] = rn(inner, d, std=2.0 * (d * dims.d_kv) ** -0.5)
            w[p + "0.SelfAttention.k.weight"] = rn(inner, d, std=d ** -0.5)
            w[p + "0.SelfAttention.v.weight"] = rn(inner, d, std=d ** -0.5)
            w[p + "0.SelfAttention.o.weight"] = rn(d, inner, std=inner ** -0.5)
            w[p + "0.layer_norm.weight"] = 1.0 + rn(d, std=0.1)
            if i == 0:
                w[p + "0.SelfAttention.relative_attention_bias.weight"] = rn(dims.num_buckets, dims.num_heads,
                                                                             std=1.0)
            ff = 1
            if side == "decoder":
                w[p + "1.EncDecAttention.q.weight"] = rn(inner, d, std=2.0 * (d * dims.d_kv) ** -0.5)
                w[p + "1.EncDecAttention.k.weight"] = rn(inner, d, std=d ** -0.5)
                w[p + "1.EncDecAttention.v.weight"] = rn(inner, d, std=d ** -0.5)
                w[p + "1.EncDecAttention.o.weight"] = rn(d, inner, std=inner ** -0.5)
                w[p + "1.layer_norm.weight"] = 1.0 + rn(d, std=0.1)
                ff = 2
            w[p + f"{ff}.DenseReluDense.wi.weight"] = rn(dff, d, std=d ** -0.5)
            w[p + f"{ff}.DenseReluDense.wo.weight"] = rn(d, dff, std=dff ** -0.5)
            w[p + f"{ff}.layer_norm.weight"] = 1.0 + rn(d, std=0.1)
        w[f"{side}.final_layer_norm.weight"] = 1.0 + rn(d, std=0.1)

    gc = torch.Generator().manual_seed(codebook_seed)
    V, L = dims.decoder_vocab_size, dims.docid_len
    out_std = logit_std / math.sqrt(d)
    for t in range(L):
        w[f"list_decoder_embeds.{t}.weight"] = torch.randn(V, d, generator=gc) * (
            out_std if dims.shared_output_input_embeds else 1.0)
    if not dims.shared_output_input_embeds:
        for t in range(L):
            w[f"list_output_embeds.{t}.weight"] = torch.randn(V, d, generator=gc) * out_std
    if start_token_embed is not None:
        st = torch.as_tensor(np.asarray(start_token_embed), dtype=torch.float32).reshape(1, 1, d)
    else:
        st = torch.randn(1, 1, d, generator=gc)
    w["start_token_embed"] = st
    return w


def make_codes(n_docs: int, L: int, V: int, seed: int = TRIE_SEED, skew: bool = False,
               dup_frac: float = 0.01) -> np.ndarray:
    """RQ code matrix ``codes[n_docs, L]`` (uint8 for V<=256, else uint16); row i is docid str(i).

    ``skew`` draws each level from Zipf(1.1) over a seeded permutation so branching continues deeper
    than with uniform codes; ``dup_frac`` of the rows are overwritten with copies of other rows so some
    smtids own several docids (reference evaluate.py:439-446 handles those with a list per smtid).
    """
    rng = np.random.default_rng(seed)
    dt = np.uint8 if V <= 256 else np.uint16
    if not skew:
        codes = rng.integers(0, V, size=(n_docs, L), dtype=np.int64).astype(dt)
    else:
        ranks = np.arange(1, V + 1, dtype=np.float64) ** -1.1
        p = ranks / ranks.sum()
        codes = np.empty((n_docs, L), dtype=dt)
        for t in range(L):
            perm = rng.permutation(V)
            codes[:, t] = perm[rng.choice(V, size=n_docs, p=p)].astype(dt)
    n_dup = int(n_docs * dup_frac)
    if n_dup > 0 and n_docs > 1:
        dst = rng.choice(n_docs, size=n_dup, replace=False)
        src = rng.integers(0, n_docs, size=n_dup)
        codes[dst] = codes[src]
    return codes


def codes_to_docid_to_smtid(codes: np.ndarray) -> Dict[str, list]:
    """The ``docid_to_smtid.json`` dict the reference loads: ``{docid: [-1, c1..cL]}`` (evaluate.py:400-401)."""
    return {str(i): [-1] + [int(x) for x in row] for i, row in enumerate(codes)}


def make_queries(batch: int, S: int = 32, vocab_size: int = 32128, seed: int = QUERY_SEED,
                 min_len: int = 8):
    """Token batch shaped like the reference's collate output: ids ~U[2, vocab-28), eos=1 last, 0-padded."""
    rng = np.random.default_rng(seed)
    hi = max(3, vocab_size - 28)
    lens = rng.integers(min(min_len, S), S + 1, size=batch)
    ids = np.zeros((batch, S), dtype=np.int64)
    mask = np.zeros((batch, S), dtype=np.int64)
    for b in range(batch):
        n = int(lens[b])
        ids[b, : n - 1] = rng.integers(2, hi, size=n - 1)
        ids[b, n - 1] = 1
        mask[b, :n] = 1
    if batch > 0:
        # force one full-length row so the padded width is S for every seed (padding="longest")
        ids[0, : S - 1] = rng.integers(2, hi, size=S - 1)
        ids[0, S - 1] = 1
        mask[0, :] = 1
    return torch.from_numpy(ids), torch.from_numpy(mask)
