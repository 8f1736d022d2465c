"""GPU parity tests of the round-2 machinery: the radix-select beam kernel (wide beams), the per-query forced tail
(ragged pass, exact score replay incl. ties), engines reused across shapes, the teacher-forced forward / rerank score,
and the device-side leaf -> document expansion. Everything goes through the C ABI and is compared with the CPU oracle
(oracle/) on the same seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import beam as ob, t5_math
from ripor_b200 import _lib, synthetic as syn
from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search
from ripor_b200.modeling import T5SeqAQEncoder
from ripor_b200.trie import DocidTrie
from tests import helpers
from tests.test_gpu_parity import _copy_from_device, _engine_search, _run_beam_kernels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sp():
    return _lib.stream_ptr()


def _table_logits(B, nb, V, seed=0, scale=3.0, integer=False):
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(B * nb, 97, V, generator=g) * scale
    if integer:
        table = torch.round(table)                    # exact ties between candidates

    def fn(ids):
        h = (ids * torch.arange(1, ids.shape[1] + 1)).sum(1) % 97
        rows = torch.arange(ids.shape[0]) // nb * nb if ids.shape[1] == 1 else torch.arange(ids.shape[0])
        return table[rows, h]
    return fn


# ------------------------------------------------------------------------------------------------
# wide beams: radix-select kernel (bit-exact on the float64 side)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_softmax", [False, True])
@pytest.mark.parametrize("n_docs,nb,V,L,B", [(200, 4, 8, 4, 3), (5000, 10, 256, 6, 5), (3000, 100, 256, 4, 2),
                                             (4000, 24, 1024, 4, 2)])
def test_select_kernel_matches_oracle_loop(monkeypatch, log_softmax, n_docs, nb, V, L, B):
    """The same cases as the arg-max kernels, forced through the radix-select formulation."""
    monkeypatch.setenv("RB200_BEAM", "select")
    codes = syn.make_codes(n_docs, L, V, seed=5, dup_frac=0.05)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    tr = DocidTrie.from_codes(codes, V).upload(0)
    fn = _table_logits(B, nb, V)
    ref_seq, ref_sc = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, V), B, nb, L,
                                            apply_log_softmax_for_scores=log_softmax)
    seqs, scores, leaf = _run_beam_kernels(tr, B, nb, L, V, fn, log_softmax)
    valid = ref_sc > -1e6
    if log_softmax:
        assert helpers.compare_ranked(seqs, scores, ref_seq, ref_sc, nb, atol=2e-6) == 0
    else:
        assert torch.equal(scores, ref_sc)
        assert torch.equal(seqs[valid], ref_seq[valid])


@pytest.mark.parametrize("log_softmax", [False, True])
def test_cta_argmax_kernel_still_exact_beyond_its_default_range(monkeypatch, log_softmax):
    """Above 8192 candidates per query the dispatcher prefers the radix select; the arg-max CTA kernel (RB200_BEAM=cta)
    must stay bit-exact there too (it remains the fallback up to 131 072 candidates)."""
    monkeypatch.setenv("RB200_BEAM", "cta")
    n_docs, nb, V, L, B = 3000, 100, 256, 4, 2
    codes = syn.make_codes(n_docs, L, V, seed=5, dup_frac=0.05)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    tr = DocidTrie.from_codes(codes, V).upload(0)
    fn = _table_logits(B, nb, V)
    ref_seq, ref_sc = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, V), B, nb, L,
                                            apply_log_softmax_for_scores=log_softmax)
    seqs, scores, leaf = _run_beam_kernels(tr, B, nb, L, V, fn, log_softmax)
    valid = ref_sc > -1e6
    if log_softmax:
        assert helpers.compare_ranked(seqs, scores, ref_seq, ref_sc, nb, atol=2e-6) == 0
    else:
        assert torch.equal(scores, ref_sc) and torch.equal(seqs[valid], ref_seq[valid])


def _adversarial_logits(kind, B, nb, V, seed):
    """Logit tables that stress the fp32 pre-filter of the warp beam kernel: its thresholds must never drop a candidate
    the float64 ranking would keep."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(B * nb, 97, V, generator=g)
    if kind == "ties":            # a handful of distinct values: most candidates tie exactly at the cut
        t = torch.round(t * 1.5)
    elif kind == "huge":          # |logit| ~ 1e9: penalised and allowed candidates interleave
        t = t * 7e8
    elif kind == "tiny":          # values far below the float64 spacing of the accumulated beam scores
        t = t * 1e-30
    elif kind == "mixed_scale":   # every row on its own scale, 1e-6 .. 1e6 (below the penalty: the oracle loop applies)
        t = t * (10.0 ** torch.randint(-6, 7, (B * nb, 97, 1), generator=g).float())
    elif kind == "nonfinite":     # NaN ranks last, infinities are ordinary extreme values
        r = torch.rand(B * nb, 97, V, generator=g)
        t = torch.where(r < 0.02, torch.full_like(t, float("nan")), t)
        t = torch.where((r >= 0.02) & (r < 0.03), torch.full_like(t, float("inf")), t)
        t = torch.where((r >= 0.03) & (r < 0.05), torch.full_like(t, float("-inf")), t)
    elif kind == "constant":      # everything ties
        t = torch.zeros_like(t) + 0.25

    def fn(ids):
        h = (ids * torch.arange(1, ids.shape[1] + 1)).sum(1) % 97
        rows = torch.arange(ids.shape[0]) // nb * nb if ids.shape[1] == 1 else torch.arange(ids.shape[0])
        return t[rows, h]
    return fn


@pytest.mark.parametrize("log_softmax", [False, True])
@pytest.mark.parametrize("kind", ["ties", "huge", "tiny", "mixed_scale", "nonfinite", "constant"])
@pytest.mark.parametrize("n_docs,nb,V,L,B", [(5000, 10, 256, 6, 7), (40, 16, 256, 5, 5), (3000, 7, 1024, 4, 3),
                                             (2000, 10, 64, 5, 4)])
def test_warp_kernel_prefilter_is_exact_on_adversarial_logits(monkeypatch, log_softmax, kind, n_docs, nb, V, L, B):
    """The warp kernel rules most candidates out with an fp32 comparison against per-beam thresholds; the CTA arg-max
    kernel evaluates every candidate in float64. Same inputs -> the same beams, bit for bit (sequences, float64-exact
    scores, leaf ranges), also with ties at the cut, |logits| around the 1e9 penalty, NaN / inf logits and small tries
    whose beams carry the penalty."""
    codes = syn.make_codes(n_docs, L, V, seed=9, dup_frac=0.05)
    tr = DocidTrie.from_codes(codes, V).upload(0)
    fn = _adversarial_logits(kind, B, nb, V, seed=21)
    got = _run_beam_kernels(tr, B, nb, L, V, fn, log_softmax)                # default dispatch: the warp kernel
    monkeypatch.setenv("RB200_BEAM", "cta")
    want = _run_beam_kernels(tr, B, nb, L, V, fn, log_softmax)
    assert torch.equal(got[0], want[0])
    assert torch.equal(got[1].view(torch.int32), want[1].view(torch.int32))  # bit pattern: NaN-safe equality
    assert torch.equal(got[2], want[2])
    if kind == "mixed_scale" and not log_softmax:                            # finite, untied logits: also the oracle loop
        lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
        ref_seq, ref_sc = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, V), B, nb, L)
        assert torch.equal(got[1], ref_sc)


@pytest.mark.parametrize("nb,V,n_docs,L", [(1000, 256, 60000, 4), (600, 256, 300, 3), (160, 1024, 50000, 3)])
def test_beam_1000_matches_oracle_loop(nb, V, n_docs, L):
    """topk = 1000 is the reference's shipped evaluation setting (full_evaluate_t5seq_aq_encoder.sh:191-199): more
    beams than the arg-max kernel can rank (nb*V > 131072) and, with nb > V, the step-0 regime in which the beams that
    start at -1e9 (generation.py:418-420) offer hundreds of exactly tied candidates."""
    B = 2
    codes = syn.make_codes(n_docs, L, V, seed=11, dup_frac=0.02)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    tr = DocidTrie.from_codes(codes, V).upload(0)
    fn = _table_logits(B, nb, V, seed=3)
    ref_seq, ref_sc = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, V), B, nb, L)
    seqs, scores, leaf = _run_beam_kernels(tr, B, nb, L, V, fn, False)
    valid = ref_sc > -1e6
    assert torch.equal(valid, scores > -1e6)
    assert torch.equal(scores[valid], ref_sc[valid])
    assert torch.equal(seqs[valid], ref_seq[valid])
    # rows that carry the -1e9 penalty are tied duplicates (torch.topk leaves their order unspecified): same multiset
    for b in range(B):
        sl = slice(b * nb, (b + 1) * nb)
        inv = ~valid[sl]
        assert torch.equal(torch.sort(scores[sl][inv]).values, torch.sort(ref_sc[sl][inv]).values)


# ------------------------------------------------------------------------------------------------
# forced tail, beam arithmetic only: T replayed steps == T real steps, bit for bit, with exact ties
# ------------------------------------------------------------------------------------------------
def _view(lib, h, what, n, dtype):
    p = C.c_void_p()
    _lib.check(lib.rb200_beam_view(h, what, C.byref(p)))
    return _copy_from_device(p.value, n * np.dtype(dtype).itemsize, dtype)


@pytest.mark.parametrize("rerank", ["count", "sort"])
@pytest.mark.parametrize("integer_logits", [False, True])
@pytest.mark.parametrize("log_softmax", [False, True])
@pytest.mark.parametrize("nb,V,n_docs,L,B", [(4, 8, 60, 10, 5), (10, 256, 4000, 12, 4), (40, 16, 3000, 9, 3),
                                             (150, 256, 60000, 8, 2)])
def test_forced_tail_replay_equals_real_steps(monkeypatch, rerank, integer_logits, log_softmax, nb, V, n_docs, L, B):
    """rb200_beam_forced_tail (the engine's tail score replay) leaves the beam state exactly where the same number of
    rb200_beam_step calls would: float64 scores, token history, leaves AND beam order. Integer-valued logits make
    candidates tie exactly at every step, so the per-step re-ranking and its tie rule (lower flat index first) decide
    the final order. Both re-ranking formulations of the replay kernel: rank counting (narrow
    beams) and the sorting network (wide beams; the default above 64 beams)."""
    monkeypatch.setenv("RB200_TAIL_SORT", "1" if rerank == "sort" else "0")
    lib = _lib.lib()
    codes = np.unique(syn.make_codes(n_docs, L, V, seed=7, dup_frac=0.0), axis=0)
    tr = DocidTrie.from_codes(codes, V).upload(0)
    R = B * nb
    g = torch.Generator().manual_seed(5)
    mk = (lambda *s: torch.round(torch.randn(*s, generator=g) * 2)) if integer_logits else \
        (lambda *s: torch.randn(*s, generator=g) * 3)
    hs = [C.c_void_p(), C.c_void_p()]
    for h in hs:
        _lib.check(lib.rb200_beam_create(0, B, nb, L, V, C.byref(h)))
    try:
        for h in hs:
            _lib.check(lib.rb200_beam_reset(h, tr.handle, B, _sp()))
        t0 = None
        for t in range(L - 2):
            lg = mk(B if t == 0 else R, V).to(DEV)
            for h in hs:
                _lib.check(lib.rb200_beam_step(h, tr.handle, lg.data_ptr(), 1 if t == 0 else nb, int(log_softmax),
                                               None, None, 0, _sp()))
            torch.cuda.synchronize()
            st = _view(lib, hs[0], 5, R * 4, np.int32).reshape(R, 4)
            if np.all(st[:, 1] - st[:, 0] == 1):
                t0 = t + 1
                break
        assert t0 is not None, "the test trie never forces every beam: pick other sizes"
        T = L - t0
        tail = mk(T, R, V)                                   # logits per LINEAGE (beam order at step t0)
        # A: real steps; the row of slot j must carry the logits of the lineage that sits in slot j
        lineage = torch.arange(R)
        for j in range(T):
            lg = tail[j][lineage].contiguous().to(DEV)
            _lib.check(lib.rb200_beam_step(hs[0], tr.handle, lg.data_ptr(), nb, int(log_softmax), None, None, 0, _sp()))
            torch.cuda.synchronize()
            parent = torch.from_numpy(_view(lib, hs[0], 1, R, np.int32).astype(np.int64))
            lineage = lineage[torch.arange(R) // nb * nb + parent]
        # B: one replay
        tail_d = tail.contiguous().to(DEV)
        _lib.check(lib.rb200_beam_forced_tail(hs[1], tr.handle, T, tail_d.data_ptr(), int(log_softmax), _sp()))
        torch.cuda.synchronize()
        assert lib.rb200_beam_current_step(hs[0]) == lib.rb200_beam_current_step(hs[1]) == L
        sc = [_view(lib, h, 0, R, np.float64) for h in hs]
        hist = [_view(lib, h, 3, R * L, np.int32) for h in hs]
        st = [_view(lib, h, 5, R * 4, np.int32).reshape(R, 4)[:, :2] for h in hs]
        assert np.array_equal(sc[0], sc[1]), np.abs(sc[0] - sc[1]).max()
        assert np.array_equal(hist[0], hist[1])
        assert np.array_equal(st[0], st[1])
        if integer_logits and not log_softmax:
            assert len(np.unique(sc[0])) < R                 # the case really holds exactly tied final scores
        outs = []
        for h in hs:
            seqs = torch.empty((R, L + 1), dtype=torch.int64, device=DEV)
            scores = torch.empty((R,), dtype=torch.float32, device=DEV)
            leaf = torch.empty((R, 2), dtype=torch.int32, device=DEV)
            _lib.check(lib.rb200_beam_finalize(h, tr.handle, nb, 1.0, seqs.data_ptr(), scores.data_ptr(),
                                               leaf.data_ptr(), _sp()))
            torch.cuda.synchronize()
            outs.append((seqs.cpu(), scores.cpu(), leaf.cpu()))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
    finally:
        for h in hs:
            lib.rb200_beam_free(h)


# ------------------------------------------------------------------------------------------------
# encoder self-attention on the tensor-core attention kernel (S = 32: relative position bias on the score fragments)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("enc_mma", ["1", "0"])
def test_encoder_states_at_source_length_32(monkeypatch, enc_mma):
    """With 32 source positions the fp16x3 engine runs the encoder's bidirectional self-attention through the
    tensor-core kernel of the forced tail (relative position bias added to the score fragments, padded keys masked);
    RB200_ENC_MMA=0 keeps it on the FFMA kernel. Both: encoder states == the oracle's T5 encoder, ragged masks."""
    monkeypatch.setenv("RB200_ENC_MMA", enc_mma)
    L, nb, B, S = 6, 4, 7, 32
    dims = syn.T5Dims.tiny(docid_len=L)
    w = syn.make_weights(dims)
    codes = syn.make_codes(500, L, dims.decoder_vocab_size)
    ids, mask = syn.make_queries(B, S=S, vocab_size=dims.vocab_size, min_len=3)
    assert mask.sum(1).min() < S                          # at least one padded sequence
    model = T5SeqAQEncoder.from_weights(dims, w)
    trie = DocidTrie.from_codes(codes, dims.decoder_vocab_size)
    out = _engine_search(model, trie, ids, mask, nb, L, precision="fp16x3")
    eng = model.base_model.get_engine(B, nb, S, "fp16x3")
    p = C.c_void_p()
    _lib.check(_lib.lib().rb200_engine_encoder_states(eng.h, C.byref(p)))
    enc = _copy_from_device(p.value, B * S * dims.d_model * 4, np.float32).reshape(B, S, -1)
    ref = t5_math.encoder_forward(w, dims, ids.long(), mask.long()).numpy()
    m = mask.numpy().astype(bool)
    assert np.abs(enc[m] - ref[m]).max() < 2e-4
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0


# ------------------------------------------------------------------------------------------------
# engine: queries freeze at different steps (skewed trie), shapes change between calls
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "tf32x3"])
@pytest.mark.parametrize("log_softmax", [False, True])
def test_per_query_forced_tail_on_skewed_trie(monkeypatch, precision, log_softmax):
    """Zipf-skewed codes: beams of different queries reach single leaves at different steps, so queries
    are frozen one by one, the step loop goes on with a compacted batch and the tail pass is ragged. Result ==
    step-by-step loop == oracle."""
    L, nb, B = 14, 5, 9
    dims = syn.T5Dims.tiny(docid_len=L)
    w = syn.make_weights(dims)
    V = dims.decoder_vocab_size
    codes = syn.make_codes(6000, L, V, skew=True, dup_frac=0.01)
    ids, mask = syn.make_queries(B, S=20, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L, log_softmax=log_softmax)
    trie = DocidTrie.from_codes(codes, V)
    outs = {}
    for tail in ("1", "0"):
        monkeypatch.setenv("RB200_TAIL", tail)
        model = T5SeqAQEncoder.from_weights(dims, w)
        outs[tail] = _engine_search(model, trie, ids, mask, nb, L, log_softmax, precision=precision)
        assert helpers.compare_ranked(outs[tail].sequences, outs[tail].sequences_scores, ref_seq, ref_sc, nb,
                                      atol=1e-3) == 0
    hist = outs["1"].frozen_at_step
    assert sum(hist) >= 2 and sum(1 for n in hist if n > 0) >= 2, hist      # frozen at two or more different steps
    assert sum(outs["0"].frozen_at_step) == 0
    assert torch.equal(outs["1"].sequences, outs["0"].sequences)
    assert torch.allclose(outs["1"].sequences_scores, outs["0"].sequences_scores, atol=2e-5, rtol=0)


def test_small_trie_with_dead_beams_never_freezes_wrongly():
    """Fewer documents than beams: -1e9 survivors stay in the beam. Such a query must keep stepping."""
    L, nb, B = 6, 8, 3
    dims = syn.T5Dims.tiny(docid_len=L)
    w = syn.make_weights(dims)
    V = dims.decoder_vocab_size
    codes = syn.make_codes(5, L, V, seed=3, dup_frac=0.0)
    ids, mask = syn.make_queries(B, S=12, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision="fp32")
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0
    assert sum(out.frozen_at_step) == 0


def test_one_engine_serves_changing_shapes():
    """The CLI pads every batch to its own longest row and the last batch is smaller; topk can change between tasks.
    One engine (one set of packed weights) must serve all of it: fewer beams than capacity, shorter / longer sources,
    bigger batches (workspaces re-allocated, weights kept)."""
    L, V = 8, 256
    dims = syn.T5Dims.tiny(docid_len=L, decoder_vocab_size=V)
    w = syn.make_weights(dims)
    codes = syn.make_codes(3000, L, V)
    trie = DocidTrie.from_codes(codes, V)
    model = T5SeqAQEncoder.from_weights(dims, w)
    engines = set()
    for B, nb, S in [(5, 6, 20), (2, 3, 11), (5, 6, 33), (9, 6, 20), (1, 40, 70), (4, 6, 16)]:
        ids, mask = syn.make_queries(B, S=S, vocab_size=dims.vocab_size, seed=100 + B + S)
        ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
        out = _engine_search(model, trie, ids, mask, nb, L, precision="tf32x3")
        assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0, (B, nb, S)
        engines.add(id(model.base_model._engines["tf32x3"]))
    assert len(engines) == 1
    assert model.base_model._engines["tf32x3"].resizes <= 4


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
def test_topk_1000_batch_1_engine(precision):
    """The reference's shipped evaluation launch: --topk 1000 --batch_size 1 (full_evaluate_t5seq_aq_encoder.sh:191-199)."""
    L, nb, B, V = 8, 1000, 1, 256
    dims = syn.T5Dims.tiny(docid_len=L, decoder_vocab_size=V)
    w = syn.make_weights(dims)
    codes = syn.make_codes(200000, L, V)
    ids, mask = syn.make_queries(B, S=16, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision=precision)
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0


# ------------------------------------------------------------------------------------------------
# teacher-forced forward / rerank_forward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("shared,scaleup", [(False, False), (True, True)])
def test_forward_and_rerank_match_full_prefix_oracle(precision, shared, scaleup):
    """T5ForDocIDGeneration.forward / T5SeqAQEncoder.rerank_forward (reference t5_generative_retriever.py:295-450,
    794-798) against the oracle's full-prefix decoder (the reference's own way of running the model)."""
    L, B = 7, 5
    dims = syn.T5Dims.tiny(docid_len=L, shared_output_input_embeds=shared, scaleup_output_hidden=scaleup)
    w = syn.make_weights(dims)
    V = dims.decoder_vocab_size
    ids, mask = syn.make_queries(B, S=18, vocab_size=dims.vocab_size)
    rng = np.random.default_rng(4)
    doc = torch.from_numpy(rng.integers(0, V, size=(B, L)).astype(np.int64))
    dec = torch.cat([torch.zeros((B, 1), dtype=torch.int64), doc[:, :-1]], dim=1)
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        hid = t5_math.decoder_full_prefix(w, dims, dec, enc, mask)                 # [B, L, d]
        logits = t5_math.lm_logits_list(w, dims, hid)
    model = T5SeqAQEncoder.from_weights(dims, w).to(DEV)
    model.base_model.precision = precision
    model.base_model.config.decoding = True
    out = model.base_model(input_ids=ids.to(DEV), attention_mask=mask.to(DEV), decoder_input_ids=dec.to(DEV))
    assert out.decoder_last_hidden_state.shape == hid.shape and len(out.logits) == L
    assert (out.decoder_last_hidden_state.cpu() - hid).abs().max() < 5e-4
    m = mask.bool()
    assert (out.encoder_last_hidden_state.cpu()[m] - enc[m]).abs().max() < 5e-4
    for p in range(L):
        assert (out.logits[p].cpu() - logits[p]).abs().max() < 1e-3, p
    model.base_model.config.decoding = False
    assert model.base_model(input_ids=ids.to(DEV), attention_mask=mask.to(DEV), decoder_input_ids=dec.to(DEV)).logits is None
    # rerank_forward: sum over positions of <hidden, output embedding of the doc's code>
    tab = [w[("list_decoder_embeds" if shared else "list_output_embeds") + f".{i}.weight"] for i in range(L)]
    ref_score = sum((hid[:, i] * tab[i][doc[:, i]]).sum(-1) for i in range(L))
    got = model.rerank_forward(tokenized_query={"input_ids": ids, "attention_mask": mask, "decoder_input_ids": dec},
                               doc_encoding=doc)
    assert torch.allclose(got.cpu(), ref_score, atol=1e-3, rtol=1e-5)
    # n candidates per query share the encoder pass
    cand = torch.from_numpy(rng.integers(0, V, size=(B, 3, L)).astype(np.int64))
    cand[:, 0] = doc
    sc = model.score_docids(ids, mask, cand).cpu()
    assert torch.allclose(sc[:, 0], ref_score, atol=1e-3, rtol=1e-5)
    # a DocID found by the beam search scores (L+1) * its sequences_score (finalize divides by L+1)
    codes = syn.make_codes(2000, L, V)
    trie = DocidTrie.from_codes(codes, V)
    res = _engine_search(model, trie, ids, mask, 4, L, precision=precision)
    best = res.sequences.view(B, 4, L + 1)[:, :1, 1:]
    sc = model.score_docids(ids, mask, best.cpu()).cpu().view(-1)
    assert torch.allclose(sc, res.sequences_scores.view(B, 4)[:, 0].cpu() * (L + 1), atol=2e-3, rtol=1e-5)


# ------------------------------------------------------------------------------------------------
# device-side leaf -> documents
# ------------------------------------------------------------------------------------------------
def test_leaf_expand_matches_host_mapping():
    """rb200_trie_leaf_expand == the host's smtid -> docids mapping (evaluate.py:439-446), incl. multi-document leaves,
    prefix ranges holding several leaves, rows that are not in the trie and rows that overflow the output width."""
    L, V, n = 6, 16, 4000
    codes = syn.make_codes(n, L, V, seed=9, dup_frac=0.2)
    tr = DocidTrie.from_codes(codes, V).upload(0)
    U = tr.n_unique
    rng = np.random.default_rng(0)
    lo = rng.integers(0, U, size=300)
    width = np.concatenate([np.ones(200, np.int64), rng.integers(2, 6, size=60), rng.integers(40, 90, size=20),
                            np.zeros(20, np.int64)])
    hi = np.minimum(lo + width, U)
    ranges = torch.from_numpy(np.stack([lo, hi], 1).astype(np.int32)).to(DEV)
    k = 32
    docs, counts = tr.expand_ranges(ranges, k)
    docs, counts = docs.cpu().numpy(), counts.cpu().numpy()
    n_over = 0
    for i in range(len(lo)):
        want = tr.rows_for_range(int(lo[i]), int(hi[i]))
        assert counts[i] == len(want)
        if len(want) <= k:
            assert np.array_equal(docs[i, : len(want)], want) and np.all(docs[i, len(want):] == -1)
        else:
            n_over += 1
            assert np.all(docs[i] == -1)
    assert n_over > 0
