"""GPU parity at the sizes bench.py measures: the CUDA path against the KV-cached CPU oracle with
an INDEPENDENT mask (oracle/range_mask.py: binary search over the sorted codes, no code or data shared with the product's
trie). BASELINE configs[1] exactly (8,841,823 docs, 32 x 256, t5-base, beam 10, batch 256; 64 queries checked), the same
on the Zipf(1.1)-skewed 1 %-duplicated trie, configs[2] (beam 100) and configs[3] (t5-large) at their real model
dimensions and L = 32, and the reference's shipped launch (batch 1, topk 1000) at t5-base dimensions.
DocID lists bit-exact, scores within 1e-3 (BASELINE.json north_star)."""
import functools

import numpy as np
import pytest
import torch

from oracle import beam as ob, t5_math
from oracle.range_mask import SortedCodesMask
from ripor_b200 import synthetic as syn
from ripor_b200.modeling import T5SeqAQEncoder
from ripor_b200.trie import DocidTrie
from tests import helpers
from tests.test_gpu_parity import _engine_search

pytestmark = pytest.mark.gpu
N_DOCS = 8841823


@functools.lru_cache(maxsize=1)
def _t5base():
    dims = syn.T5Dims.t5_base(docid_len=32)
    return dims, syn.make_weights(dims)


@functools.lru_cache(maxsize=1)
def _codes(kind):
    return syn.make_codes(N_DOCS, 32, 256, skew=(kind == "zipf"))


def _oracle(w, dims, codes, V, ids, mask, nb, L, trace=None):
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        dec = t5_math.CachedDecoder(w, dims, enc, mask, nb)

        def step(dec_ids, bi):
            if bi is not None:
                dec.reorder(bi)
            return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])
        return ob.beam_search_oracle(step, SortedCodesMask(codes, V), ids.shape[0], nb, L, trace=trace)


@pytest.mark.parametrize("kind", ["uniform", "zipf"])
def test_config2_exact_size_64_queries(kind):
    """BASELINE configs[1]: batch 256 runs on the GPU (the benchmarked shape, so the same kernels, tiles and forced-tail
    partition as bench.py); the first 64 queries are checked against the oracle."""
    B, nb, L, V, PQ = 256, 10, 32, 256, 64
    dims, w = _t5base()
    codes = _codes(kind)
    ids, mask = syn.make_queries(B, S=32)
    ref_seq, ref_sc = _oracle(w, dims, codes, V, ids[:PQ], mask[:PQ], nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision="auto")
    assert out.precision == "fp16x3"
    seqs = out.sequences.view(B, nb, L + 1)[:PQ].reshape(PQ * nb, L + 1)
    scores = out.sequences_scores.view(B, nb)[:PQ].reshape(-1)
    assert helpers.compare_ranked(seqs, scores, ref_seq, ref_sc, nb, atol=1e-3) == 0
    hist = out.frozen_at_step
    assert sum(hist) == B                         # at 8.8 M documents every query is frozen well before step 32
    if kind == "zipf":
        assert sum(1 for n in hist if n > 0) >= 3, hist      # ... at different steps on the skewed trie
    # size-independent properties over ALL 256 queries (the oracle covers the first 64):
    # every ranked DocID is a full code of the collection (exactly one leaf) and re-spells through the host trie walk,
    leaf = out.leaf_ranges.cpu()
    assert torch.all(leaf[:, 1] - leaf[:, 0] == 1)
    trie = DocidTrie.from_codes(codes, V)
    all_seqs = out.sequences.cpu()
    for r in range(0, B * nb, 97):
        assert trie.find_leaf(all_seqs[r, 1:].tolist()) == int(leaf[r, 0])
    # scores are finite and non-increasing within a query, the nb DocIDs of a query are distinct, column 0 is the
    # decoder start id,
    sc = out.sequences_scores.view(B, nb).cpu()
    assert torch.isfinite(sc).all() and torch.all(sc[:, :-1] >= sc[:, 1:])
    assert torch.all(all_seqs[:, 0] == 0)
    per_q = all_seqs.view(B, nb, L + 1)
    assert all(len({tuple(row) for row in per_q[b].tolist()}) == nb for b in range(B))
    # and the search is idempotent: the same batch again gives the same bits (no state leaks between calls)
    again = _engine_search(model, trie, ids, mask, nb, L, precision="auto")
    assert torch.equal(again.sequences, out.sequences) and torch.equal(again.sequences_scores, out.sequences_scores)
    # the device-side leaf expansion returns the rows whose codes spell the ranked DocIDs (1 % of the rows are duplicates)
    docs, counts = trie.expand_ranges(out.leaf_ranges, 8)
    docs, counts = docs.cpu().numpy(), counts.cpu().numpy()
    assert counts.min() >= 1
    for r in range(0, B * nb, 211):
        rows = docs[r, : counts[r]]
        assert np.all(codes[rows] == all_seqs[r, 1:].numpy())


def test_config3_beam100_t5base_dims():
    """BASELINE configs[2] at real dimensions: t5-base, L = 32, beam 100 (CTA beam kernel with nb*V = 25,600
    candidates, forced tail with 100 beams per query), 8.8 M-doc trie; 4 queries."""
    B, nb, L, V = 4, 100, 32, 256
    dims, w = _t5base()
    codes = _codes("uniform")
    ids, mask = syn.make_queries(B, S=32, seed=901)
    ref_seq, ref_sc = _oracle(w, dims, codes, V, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision="auto")
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0
    assert sum(out.frozen_at_step) == B


def test_config4_t5large_dims_L32():
    """BASELINE configs[3] at real dimensions: t5-large (d = 1024, 16 heads, d_ff = 4096, 24 + 24 layers), L = 32,
    beam 10, 8.8 M-doc trie; 4 queries."""
    B, nb, L, V = 4, 10, 32, 256
    dims = syn.T5Dims.t5_large(docid_len=L)
    w = syn.make_weights(dims)
    codes = _codes("uniform")
    ids, mask = syn.make_queries(B, S=32, seed=902)
    ref_seq, ref_sc = _oracle(w, dims, codes, V, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision="auto")
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0


def test_shipped_launch_topk1000_batch1_t5base_dims():
    """full_scripts/full_evaluate_t5seq_aq_encoder.sh:191-199 (--batch_size=1 --topk=1000) at t5-base dimensions over
    the 8.8 M-doc trie: radix-select beam kernel, 1000-beam forced tail. DocID length 8 keeps the CPU oracle (1000 rows
    per step) inside a minute; the L = 32 shape is exercised at tiny dimensions in test_gpu_round2.py."""
    B, nb, L, V = 1, 1000, 8, 256
    dims = syn.T5Dims.t5_base(docid_len=L)
    w = syn.make_weights(dims)
    codes = _codes("uniform")[:, :L]
    ids, mask = syn.make_queries(B, S=32, seed=903)
    trace = []
    ref_seq, ref_sc = _oracle(w, dims, np.ascontiguousarray(codes), V, ids, mask, nb, L, trace=trace)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(np.ascontiguousarray(codes), V), ids, mask, nb, L, precision="auto")
    # 1000 beams cut a dense candidate list: the gap at the cut can be below fp32 noise, which is a near-tie and is
    # reported apart from real mismatches (hard part 1)
    real, near = helpers.compare_ranked_near_tie(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, trace)
    assert real == 0, (real, near)


def test_config5_16x1024_t5base_dims():
    """BASELINE configs[4] at real dimensions: t5-base with 16 codebooks of 1024 codes (full_16_1024_scripts), beam 10,
    8.8 M-doc trie with 2-byte codes; 8 queries."""
    B, nb, L, V = 8, 10, 16, 1024
    dims = syn.T5Dims.t5_base(docid_len=L, decoder_vocab_size=V)
    w = syn.make_weights(dims)
    codes = syn.make_codes(N_DOCS, L, V)
    ids, mask = syn.make_queries(B, S=32, seed=904)
    ref_seq, ref_sc = _oracle(w, dims, codes, V, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision="auto")
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0
    assert sum(out.frozen_at_step) == B


def test_shipped_training_data_launch_prefix_search():
    """full_scripts/full_evaluate_t5seq_aq_encoder.sh:128-139 (t5seq_aq_get_qid_to_smtid_rankdata: --batch_size=4
    --topk=100 --max_new_token=8) at t5-base dimensions: a PREFIX search (8 of the trie's 32 positions), so the ranked
    rows are leaf ranges that may hold several documents."""
    B, nb, L, V = 4, 100, 8, 256
    dims, w = _t5base()
    codes = _codes("uniform")
    ids, mask = syn.make_queries(B, S=32, seed=905)
    ref_seq, ref_sc = _oracle(w, dims, codes, V, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    trie = DocidTrie.from_codes(codes, V)
    out = _engine_search(model, trie, ids, mask, nb, L, precision="auto")
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0
    # every ranked prefix exists in the collection: its leaf range is not empty and spells the prefix
    leaf = out.leaf_ranges.cpu()
    assert torch.all(leaf[:, 1] > leaf[:, 0])
    r = 7
    rows = trie.rows_for_range(int(leaf[r, 0]), int(leaf[r, 1]))
    assert len(rows) >= 1 and np.all(codes[rows][:, :L] == out.sequences[r, 1:].cpu().numpy())


@pytest.mark.parametrize("kind", ["uniform", "zipf"])
def test_t5base_variants_log_softmax_long_source_fewer_returns(kind):
    """The remaining switches of the path together at t5-base dimensions over the 8.8 M-doc tries: log-softmax scores
    (apply_log_softmax_for_scores, evaluate.py:123-126), shared input/output codebook tables and d^-1/2 scaling
    (t5_generative_retriever.py:254-258,427-428), a padded source length beyond one 32-key attention chunk (S = 48: the
    ragged tail's multi-chunk cross-attention), and num_return_sequences < num_beams."""
    B, nb, keep, L, V = 12, 10, 3, 32, 256
    dims = syn.T5Dims.t5_base(docid_len=L, shared_output_input_embeds=True, scaleup_output_hidden=True)
    w = syn.make_weights(dims)
    codes = _codes(kind)
    ids, mask = syn.make_queries(B, S=48, seed=906)
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        dec = t5_math.CachedDecoder(w, dims, enc, mask, nb)

        def step(dec_ids, bi):
            if bi is not None:
                dec.reorder(bi)
            return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])
        ref_seq, ref_sc = ob.beam_search_oracle(step, SortedCodesMask(codes, V), B, nb, L, num_return_sequences=keep,
                                                apply_log_softmax_for_scores=True)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, log_softmax=True, keep=keep,
                         precision="auto")
    assert out.sequences.shape == (B * keep, L + 1)
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, keep, atol=1e-3) == 0
    assert sum(out.frozen_at_step) == B
