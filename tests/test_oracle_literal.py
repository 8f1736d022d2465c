"""Pin oracle/beam.py to the reference's own code (literal beam loop + literal mask class)."""
import numpy as np
import pytest
import torch

from oracle import beam as ob
from oracle import ref_literal, t5_math
from ripor_b200 import synthetic as syn

pytestmark = pytest.mark.reference


def _trie(n, L, V, seed, skew=False):
    codes = syn.make_codes(n, L, V, seed=seed, skew=skew, dup_frac=0.05)
    d2s = syn.codes_to_docid_to_smtid(codes)
    return codes, d2s, ob.build_list_smtid_to_nextids(d2s)


def test_hand_trie_masks_match_literal():
    d2s = {"0": [-1, 0, 1, 2], "1": [-1, 0, 1, 3], "2": [-1, 0, 2, 0], "3": [-1, 3, 3, 3], "4": [-1, 3, 0, 1],
           "5": [-1, 3, 3, 3]}
    lst = ob.build_list_smtid_to_nextids(d2s)
    lit = ref_literal.literal_processor(lst, 4)
    mine = ob.TrieMaskOracle(lst, 4)
    for ids in ([[0], [0]], [[0, 0], [0, 3], [0, 1]], [[0, 0, 1], [0, 3, 3], [0, 2, 2], [0, 3, 0]]):
        t = torch.tensor(ids)
        a, b = lit(t, None), mine(t, None)
        assert a.dtype == torch.float64 and b.dtype == torch.float64
        assert torch.equal(a, b)
    # unknown prefix rows are all-zero (generation.py:656-661,675)
    assert mine(torch.tensor([[0, 1]]), None).sum() == 0
    assert mine(torch.tensor([[0, 0], [0, 3]]), None).tolist() == [[0, 1, 1, 0], [1, 0, 0, 1]]


@pytest.mark.parametrize("seed,skew", [(1, False), (2, True)])
def test_random_trie_masks_match_literal(seed, skew):
    codes, d2s, lst = _trie(3000, 5, 16, seed, skew)
    lit = ref_literal.literal_processor(lst, 16)
    mine = ob.TrieMaskOracle(lst, 16)
    rng = np.random.default_rng(seed)
    for T in range(1, 6):
        rows = codes[rng.integers(0, len(codes), 64)][:, : T - 1].astype(np.int64)
        rnd = rng.integers(0, 16, size=(32, T - 1))
        ids = np.concatenate([rows, rnd], 0)
        ids = np.concatenate([np.zeros((len(ids), 1), np.int64), ids], 1)
        t = torch.from_numpy(ids)
        assert torch.equal(lit(t, None), mine(t, None))


@pytest.mark.parametrize("log_softmax", [False, True])
@pytest.mark.parametrize("n_docs,nb", [(200, 4), (6, 5)])
def test_beam_loop_matches_literal_table_model(log_softmax, n_docs, nb):
    """Scorer + loop + mask on a model whose logits are a hash of the prefix (no T5 arithmetic).
    n_docs=6 < nb exercises the -1e9 survivors."""
    L, V, B = 4, 8, 3
    codes, d2s, lst = _trie(n_docs, L, V, seed=5)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(B * nb, 97, V, generator=g)

    def fn(ids):
        h = (ids * torch.arange(1, ids.shape[1] + 1)).sum(1) % 97
        rows = torch.arange(ids.shape[0]) // nb * nb          # same logits for all beams of a query at t=0
        return table[rows if ids.shape[1] == 1 else torch.arange(ids.shape[0]), h]

    lit = ref_literal.literal_beam_search(ref_literal.LogitsTableModelAdapter(fn),
                                          ref_literal.literal_processor(lst, V), B, nb, L,
                                          apply_log_softmax_for_scores=log_softmax)
    seqs, scores = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, V), B, nb, L,
                                         apply_log_softmax_for_scores=log_softmax)
    assert lit.sequences_scores.dtype == torch.float32 and scores.dtype == torch.float32
    assert lit.scores[0].dtype == torch.float64          # SURVEY.md §0 finding 5
    assert torch.equal(lit.sequences, seqs)
    assert torch.equal(lit.sequences_scores, scores)


def test_full_prefix_t5_matches_literal_and_cached():
    """End to end at tiny T5 dims: literal loop + full-prefix adapter == restated loop + full prefix;
    the KV-cached decoder gives the same DocIDs and scores within 1e-5."""
    dims = syn.T5Dims.tiny()
    w = syn.make_weights(dims)
    B, nb, L, V = 3, 4, dims.docid_len, dims.decoder_vocab_size
    codes, d2s, lst = _trie(300, L, V, seed=11)
    ids, mask = syn.make_queries(B, S=12, vocab_size=dims.vocab_size)
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        lit = ref_literal.literal_beam_search(ref_literal.FullPrefixModelAdapter(w, dims),
                                              ref_literal.literal_processor(lst, V), B, nb, L,
                                              encoder_states=enc, attention_mask=mask)
        idx = torch.arange(B).repeat_interleave(nb)
        enc_r, mask_r = enc[idx], mask[idx]

        def full(dec_ids, bi):
            h = t5_math.decoder_full_prefix(w, dims, dec_ids, enc_r, mask_r)
            return h[:, -1, :] @ t5_math.output_table(w, dims, dec_ids.shape[1] - 1).t()

        seqs, scores = ob.beam_search_oracle(full, ob.TrieMaskOracle(lst, V), B, nb, L)
        assert torch.equal(lit.sequences, seqs)
        assert torch.equal(lit.sequences_scores, scores)

        dec = t5_math.CachedDecoder(w, dims, enc, mask, nb)

        def cached(dec_ids, bi):
            if bi is not None:
                dec.reorder(bi)
            return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])

        seqs_c, scores_c = ob.beam_search_oracle(cached, ob.TrieMaskOracle(lst, V), B, nb, L)
        assert torch.equal(seqs_c, seqs)
        assert torch.allclose(scores_c, scores, atol=1e-5, rtol=0)
