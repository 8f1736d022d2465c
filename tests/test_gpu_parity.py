"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs
and against the golden vectors made from the reference's literal loop."""
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

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sp():
    return _lib.stream_ptr()


# ------------------------------------------------------------------------------------------------
# trie mask on device (integer work: bit-exact)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,L,V,skew", [(6, 3, 4, False), (3000, 5, 16, True), (200000, 8, 256, False),
                                        (5000, 4, 1024, True)])
def test_device_mask_equals_oracle(n, L, V, skew):
    codes = syn.make_codes(n, L, V, seed=5, skew=skew, dup_frac=0.05)
    tr = DocidTrie.from_codes(codes, V).upload(0)
    rng = np.random.default_rng(1)
    for T in range(1, L + 1):
        rows = codes[rng.integers(0, n, 300)][:, : T - 1].astype(np.int64)
        rnd = rng.integers(0, V, size=(100, T - 1))
        ids = np.concatenate([rows, rnd], 0)
        ids = torch.from_numpy(np.concatenate([np.zeros((len(ids), 1), np.int64), ids], 1))
        host = tr.mask(ids)                       # host walk, already pinned to the oracle on CPU
        dev = tr.mask(ids.to(DEV))
        assert dev.dtype == torch.float64 and torch.equal(dev.cpu(), host)


# ------------------------------------------------------------------------------------------------
# beam step + finalize alone, fed with the oracle's logits (float64 beam arithmetic: bit-exact)
# ------------------------------------------------------------------------------------------------
def _run_beam_kernels(tr, B, nb, L, V, logits_fn, log_softmax, keep=None):
    keep = keep or nb
    lib = _lib.lib()
    h = C.c_void_p()
    _lib.check(lib.rb200_beam_create(0, B, nb, L, V, C.byref(h)))
    try:
        _lib.check(lib.rb200_beam_reset(h, tr.handle, B, _sp()))
        ids = torch.zeros((B * nb, 1), dtype=torch.long)
        for t in range(L):
            lg = logits_fn(ids)                                   # [B*nb, V] fp32 (CPU)
            if t == 0:
                lg = lg.view(B, nb, V)[:, 0].contiguous()
            lg_d = lg.to(DEV)
            _lib.check(lib.rb200_beam_step(h, tr.handle, lg_d.data_ptr(), 1 if t == 0 else nb, int(log_softmax),
                                           None, None, 0, _sp()))
            p = C.c_void_p()
            _lib.check(lib.rb200_beam_view(h, 3, C.byref(p)))
            torch.cuda.synchronize()
            hist_np = _copy_from_device(p.value, B * nb * L * 4, np.int32).reshape(B * nb, L)
            ids = torch.cat([torch.zeros((B * nb, 1), dtype=torch.long),
                             torch.from_numpy(hist_np[:, : t + 1].astype(np.int64))], dim=1)
        seqs = torch.empty((B * keep, L + 1), dtype=torch.int64, device=DEV)
        scores = torch.empty((B * keep,), dtype=torch.float32, device=DEV)
        leaf = torch.empty((B * keep, 2), dtype=torch.int32, device=DEV)
        _lib.check(lib.rb200_beam_finalize(h, tr.handle, keep, 1.0, seqs.data_ptr(), scores.data_ptr(), leaf.data_ptr(),
                                           _sp()))
        torch.cuda.synchronize()
        return seqs.cpu(), scores.cpu(), leaf.cpu()
    finally:
        lib.rb200_beam_free(h)


def _copy_from_device(ptr, nbytes, dtype):
    """Bytes behind a raw device pointer returned by the C ABI, as a numpy array (zero-copy view -> .cpu())."""
    class _Holder:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Holder(), device=DEV).cpu().numpy().view(dtype)


@pytest.mark.parametrize("log_softmax", [False, True])
@pytest.mark.parametrize("n_docs,nb,V,L,B", [(200, 4, 8, 4, 3), (6, 5, 8, 4, 3), (5000, 10, 256, 6, 5),
                                             (3000, 100, 256, 4, 2), (4000, 10, 1024, 4, 2)])
def test_beam_kernels_match_oracle_loop(log_softmax, n_docs, nb, V, L, B):
    codes = syn.make_codes(n_docs, L, V, seed=5, dup_frac=0.05)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    tr = DocidTrie.from_codes(codes, V).upload(0)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(B * nb, 97, V, generator=g) * 3

    def fn(ids):
        h = (ids * torch.arange(1, ids.shape[1] + 1)).sum(1) % 97
        rows = torch.arange(ids.shape[0]) // nb * nb if ids.shape[1] == 1 else torch.arange(ids.shape[0])
        return table[rows, h]

    ref_seq, ref_sc = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, V), B, nb, L,
                                            apply_log_softmax_for_scores=log_softmax)
    seqs, scores, leaf = _run_beam_kernels(tr, B, nb, L, V, fn, log_softmax)
    valid = ref_sc > -1e6
    if log_softmax:
        assert helpers.compare_ranked(seqs, scores, ref_seq, ref_sc, nb, atol=2e-6) == 0
    else:
        assert torch.equal(scores, ref_sc)                           # float64 path: bit-exact
        assert torch.equal(seqs[valid], ref_seq[valid])
    # leaf ranges: valid rows are exactly one leaf that spells the row's code
    for r in torch.nonzero(valid).view(-1).tolist():
        lo, hi = leaf[r].tolist()
        assert hi - lo == 1 and tr.find_leaf(seqs[r, 1:].tolist()) == lo
    assert torch.all((leaf[~valid, 1] - leaf[~valid, 0]) == 0)


def test_finalize_num_return_and_errors():
    codes = syn.make_codes(500, 3, 8, seed=2)
    tr = DocidTrie.from_codes(codes, 8).upload(0)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    g = torch.Generator().manual_seed(1)
    table = torch.randn(8, 31, 8, generator=g)
    fn = lambda ids: table[torch.arange(ids.shape[0]) // 4 * 4 if ids.shape[1] == 1 else torch.arange(ids.shape[0]),
                           (ids * torch.arange(1, ids.shape[1] + 1)).sum(1) % 31]
    ref_seq, ref_sc = ob.beam_search_oracle(lambda ids, bi: fn(ids), ob.TrieMaskOracle(lst, 8), 2, 4, 3,
                                            num_return_sequences=2)
    seqs, scores, _ = _run_beam_kernels(tr, 2, 4, 3, 8, fn, False, keep=2)
    assert torch.equal(seqs, ref_seq) and torch.equal(scores, ref_sc)
    lib = _lib.lib()
    h = C.c_void_p()
    _lib.check(lib.rb200_beam_create(0, 2, 4, 3, 8, C.byref(h)))
    with pytest.raises(ValueError):
        _lib.check(lib.rb200_beam_step(h, tr.handle, 1, 4, 0, None, None, 0, None))     # before reset
    lib.rb200_beam_free(h)


# ------------------------------------------------------------------------------------------------
# GEMM family vs float64 matmul
# ------------------------------------------------------------------------------------------------
GEMM_TOL = {"fp32": 2e-6, "tf32x3": 3e-5, "fp16x3": 3e-5, "bf16x3": 6e-5, "tf32": 2e-3, "bf16": 1.5e-2}


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "fp16x3", "bf16x3", "tf32", "bf16"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 256, 128), (2560, 2304, 768), (77, 16, 128), (40, 768, 3072),
                                   (300, 3072, 768), (1, 256, 768),
                                   (400, 2304, 768), (1000, 768, 768),   # 64-wide tiles: one round of tiles, small batches
                                   (40000, 512, 64)])    # many tiles per SM pair: the 256-wide tile path of the tail
def test_gemm_modes_against_float64(mode, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) * K ** -0.5
    ref = (A.double() @ W.double().t())
    Ad, Wd = A.to(DEV), W.to(DEV)
    Cd = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(_lib.lib().rb200_gemm(_lib.PRECISIONS[mode], Ad.data_ptr(), Wd.data_ptr(), Cd.data_ptr(), M, N, K,
                                     0, 0, _sp()))
    torch.cuda.synchronize()
    err = (Cd.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < GEMM_TOL[mode], (mode, M, N, K, err)
    # residual and relu epilogues
    C0 = torch.randn(M, N, generator=g)
    Cd = C0.to(DEV)
    _lib.check(_lib.lib().rb200_gemm(_lib.PRECISIONS[mode], Ad.data_ptr(), Wd.data_ptr(), Cd.data_ptr(), M, N, K,
                                     1, 0, _sp()))
    err = (Cd.cpu().double() - (ref + C0.double())).abs().max().item() / ref.abs().max().item()
    assert err < GEMM_TOL[mode], (mode, "residual", err)
    Cd = torch.zeros((M, N), device=DEV)
    _lib.check(_lib.lib().rb200_gemm(_lib.PRECISIONS[mode], Ad.data_ptr(), Wd.data_ptr(), Cd.data_ptr(), M, N, K,
                                     0, 1, _sp()))
    tol = GEMM_TOL[mode] if mode in ("fp32", "tf32x3", "fp16x3", "bf16x3") else 1.5e-2
    err = (Cd.cpu().double() - ref.clamp(min=0)).abs().max().item() / ref.abs().max().item()
    assert err < max(tol, 1e-5), (mode, "relu", err)


# ------------------------------------------------------------------------------------------------
# engine: encoder, single decoder steps, end-to-end search
# ------------------------------------------------------------------------------------------------
def _engine_search(model, trie, ids, mask, nb, L, log_softmax=False, keep=None, precision=None, host=False):
    proc = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
    model.to(DEV)
    if not host:
        ids, mask = ids.to(DEV), mask.to(DEV)
    out = generate_for_constrained_prefix_beam_search(
        model.base_model, proc, input_ids=ids.long(), attention_mask=mask.long(), max_new_tokens=L, output_scores=True,
        return_dict=True, return_dict_in_generate=True, num_beams=nb, num_return_sequences=keep or nb,
        apply_log_softmax_for_scores=log_softmax, precision=precision)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("name", ["tiny_plain", "tiny_logsoftmax", "tiny_shared_scaleup", "tiny_few_docs", "c1_t5base"])
def test_search_matches_golden_from_literal_reference(name, precision):
    c, dims, w, codes, ids, mask, g = helpers.load_golden(name)
    model = T5SeqAQEncoder.from_weights(dims, w)
    trie = DocidTrie.from_codes(codes, c["V"])
    out = _engine_search(model, trie, ids, mask, c["nb"], c["L"], c["log_softmax"], precision=precision)
    # encoder states
    eng = model.base_model.get_engine(c["B"], c["nb"], c["S"], precision)
    p = C.c_void_p()
    _lib.check(_lib.lib().rb200_engine_encoder_states(eng.h, C.byref(p)))
    enc = _copy_from_device(p.value, c["B"] * c["S"] * dims.d_model * 4, np.float32).reshape(c["B"], c["S"], -1)
    m = g["attention_mask"].astype(bool)
    assert np.abs(enc[m] - g["encoder_states"][m]).max() < 2e-4
    ref_seq, ref_sc = torch.from_numpy(g["sequences"]), torch.from_numpy(g["sequences_scores"])
    assert out.sequences.dtype == torch.int64 and out.sequences_scores.dtype == torch.float32
    assert out.sequences.shape == ref_seq.shape and torch.all(out.sequences[:, 0] == 0)
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, c["nb"], atol=1e-3) == 0
    # host mapping (evaluate.py:116-128) gives the same run dict
    import json
    d2s = syn.codes_to_docid_to_smtid(codes)
    run = ob.rankdata_for_batch(list(range(c["B"])), out.sequences.cpu(), out.sequences_scores.cpu(),
                                ob.build_smtid_to_docids(d2s, c["L"]), c["nb"], c["L"], c["log_softmax"])
    gold = json.loads(str(g["run_json"]))
    for q in range(c["B"]):
        assert list(run[q].keys()) == list(gold[str(q)].keys())
        for k in run[q]:
            assert abs(run[q][k] - gold[str(q)][k]) < 1e-3 * c["L"]


@pytest.mark.parametrize("precision,fold", [("fp32", "1"), ("tf32x3", "1"), ("fp16x3", "1"), ("bf16x3", "1"),
                                            ("tf32x3", "0"), ("fp16x3", "0")])
def test_t5base_search_matches_cached_oracle(monkeypatch, precision, fold):
    """t5-base, 100k-doc trie, L=32, beam=10: DocIDs exact, scores within 1e-3 (north-star tolerances). fold = 1 is
    the default engine (layer norms inside the GEMM epilogues), fold = 0 the separate RMSNorm launches."""
    monkeypatch.setenv("RB200_FOLD", fold)
    L, nb, B, V = 32, 10, 6, 256
    dims = syn.T5Dims.t5_base(docid_len=L)
    w = syn.make_weights(dims)
    codes = syn.make_codes(100000, L, V)
    ids, mask = syn.make_queries(B, S=32)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    trie = DocidTrie.from_codes(codes, V)
    out = _engine_search(model, trie, ids, mask, nb, L, precision=precision, host=True)
    assert not out.sequences.is_cuda
    n_mismatch = helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3)
    assert n_mismatch == 0, f"{n_mismatch} of {B} queries differ"
    assert out.gpu_launches > 100


def test_fast_modes_run_and_stay_close():
    """tf32 / bf16 single-pass modes are throughput modes: not parity-safe, but scores must stay close."""
    L, nb, B, V = 8, 5, 4, 256
    dims = syn.T5Dims.t5_base(docid_len=L)
    w = syn.make_weights(dims)
    codes = syn.make_codes(1000, L, V)
    ids, mask = syn.make_queries(B, S=32)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    trie = DocidTrie.from_codes(codes, V)
    for precision, tol in (("tf32", 2e-2), ("bf16", 1e-1)):
        out = _engine_search(model, trie, ids, mask, nb, L, precision=precision)
        top = out.sequences_scores.view(B, nb)[:, 0].cpu()
        assert torch.allclose(top, ref_sc.view(B, nb)[:, 0], atol=tol), precision


@pytest.mark.parametrize("fold", ["0", "1"])
def test_fp16x3_overflow_is_loud(monkeypatch, fold):
    """fp16x3 cannot represent |x| > 65504: the engine must return NaN scores (or refuse the weights), never silently
    wrong DocIDs. fold = 0: separate RMSNorm launches, the layer-norm gain acts on the activations; fold = 1 (the
    default): the gain is folded into the packed wi matrix, which then no longer fits the fp16 planes."""
    monkeypatch.setenv("RB200_FOLD", fold)
    dims = syn.T5Dims.tiny()
    w = syn.make_weights(dims)
    w = dict(w)
    # weights that still fit the fp16 planes (|w| * 2^8 < 65504) but whose activations do not: the layer norm gain
    # puts the normalised input at ~150 and wi (std ~22) lifts the FFN hidden state to ~1e5 > 65504
    w["decoder.block.0.layer.2.layer_norm.weight"] = w["decoder.block.0.layer.2.layer_norm.weight"] * 150
    w["decoder.block.0.layer.2.DenseReluDense.wi.weight"] = w["decoder.block.0.layer.2.DenseReluDense.wi.weight"] * 250
    codes = syn.make_codes(400, dims.docid_len, dims.decoder_vocab_size)
    ids, mask = syn.make_queries(2, S=12, vocab_size=dims.vocab_size)
    model = T5SeqAQEncoder.from_weights(dims, w)
    trie = DocidTrie.from_codes(codes, dims.decoder_vocab_size)
    if fold == "0":
        out = _engine_search(model, trie, ids, mask, 4, dims.docid_len, precision="fp16x3")
        assert torch.isnan(out.sequences_scores).all()
    else:
        with pytest.raises(ValueError, match="fp16 range"):      # gain x wi does not fit the fp16 planes: refused
            _engine_search(model, trie, ids, mask, 4, dims.docid_len, precision="fp16x3")
    out = _engine_search(model, trie, ids, mask, 4, dims.docid_len, precision="tf32x3")
    assert not torch.isnan(out.sequences_scores).any()
    # precision="auto" starts in fp16x3, sees the poisoned scores (or the refusal) and redoes the batch in tf32x3
    auto = _engine_search(model, trie, ids, mask, 4, dims.docid_len, precision="auto")
    assert auto.precision == "tf32x3" and model.base_model.fp16_ok is False
    assert torch.equal(auto.sequences, out.sequences) and torch.equal(auto.sequences_scores, out.sequences_scores)
    # a weight that does not fit fp16 after the 2^8 pre-scale is refused when the engine is built
    w["decoder.block.0.layer.2.DenseReluDense.wi.weight"] = w["decoder.block.0.layer.2.DenseReluDense.wi.weight"] * 3e4
    with pytest.raises(ValueError, match="fp16 range"):
        _engine_search(T5SeqAQEncoder.from_weights(dims, w), trie, ids, mask, 4, dims.docid_len, precision="fp16x3")


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("L,S,nb", [(40, 45, 12), (70, 100, 3)])
def test_long_docid_and_source_lengths(precision, L, S, nb):
    """More than 32 DocID positions / source tokens and more than 10 beams: the multi-chunk attention paths."""
    dims = syn.T5Dims.tiny(docid_len=L)
    w = syn.make_weights(dims)
    V, B = dims.decoder_vocab_size, 3
    codes = syn.make_codes(300, L, V)
    ids, mask = syn.make_queries(B, S=S, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision=precision)
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0


@pytest.mark.parametrize("precision", ["tf32x3", "fp16x3"])
@pytest.mark.parametrize("nb,V,L,n_docs,B", [(100, 256, 6, 3000, 2),      # BASELINE configs[2]: beam 100
                                             (10, 1024, 16, 5000, 3)])    # BASELINE configs[4]: 16 x 1024 codebooks
def test_wide_beam_and_wide_codebook_configs(precision, nb, V, L, n_docs, B):
    dims = syn.T5Dims.tiny(docid_len=L, decoder_vocab_size=V)
    w = syn.make_weights(dims)
    codes = syn.make_codes(n_docs, L, V)
    ids, mask = syn.make_queries(B, S=20, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision=precision)
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0


def test_t5large_dims_match_cached_oracle():
    """BASELINE configs[3]: t5-large (d=1024, 16 heads, d_ff=4096, 24+24 layers) at a size the CPU oracle finishes."""
    L, nb, B, V = 4, 4, 2, 256
    dims = syn.T5Dims.t5_large(docid_len=L)
    w = syn.make_weights(dims)
    codes = syn.make_codes(20000, L, V)
    ids, mask = syn.make_queries(B, S=16)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w)
    out = _engine_search(model, DocidTrie.from_codes(codes, V), ids, mask, nb, L, precision="fp16x3")
    assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("log_softmax", [False, True])
def test_forced_tail_equals_step_by_step(monkeypatch, precision, log_softmax):
    """Once every beam sits on a single trie leaf the engine runs the remaining positions as one teacher-forced pass;
    the result must be what the step-by-step loop gives (and what the oracle gives)."""
    L, nb, B = 12, 6, 5
    dims = syn.T5Dims.tiny(docid_len=L)
    w = syn.make_weights(dims)
    V = dims.decoder_vocab_size
    codes = syn.make_codes(3000, L, V)
    ids, mask = syn.make_queries(B, S=20, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L, log_softmax=log_softmax)
    trie = DocidTrie.from_codes(codes, V)
    outs = {}
    for tail in ("1", "0"):
        monkeypatch.setenv("RB200_TAIL", tail)                      # read when the engine is created
        model = T5SeqAQEncoder.from_weights(dims, w)
        outs[tail] = _engine_search(model, trie, ids, mask, nb, L, log_softmax, precision=precision)
        assert helpers.compare_ranked(outs[tail].sequences, outs[tail].sequences_scores, ref_seq, ref_sc, nb,
                                      atol=1e-3) == 0
    assert 1 <= outs["1"].forced_tail_from < L - 1 and outs["0"].forced_tail_from == -1
    assert torch.equal(outs["1"].sequences, outs["0"].sequences)
    assert torch.allclose(outs["1"].sequences_scores, outs["0"].sequences_scores, atol=2e-5, rtol=0)
