"""CPU suite: the C-ABI library loads and exports what the header declares; the host-side trie matches
the oracle; the restated oracle reproduces the golden vectors made from the reference's literal loop."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import beam as ob, t5_math
from ripor_b200 import _lib, synthetic as syn
from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search
from ripor_b200.trie import DocidTrie
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "riporb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()                      # binds every symbol; AttributeError if one is missing
    assert b"sm_100a" in lib.rb200_version()
    for name in declared:
        assert hasattr(lib, name)


@pytest.mark.parametrize("n,L,V,skew", [(6, 3, 4, False), (3000, 5, 16, True), (50000, 6, 256, False),
                                        (2000, 4, 1024, False), (40, 8, 256, False)])
def test_trie_matches_reference_dicts(n, L, V, skew):
    codes = syn.make_codes(n, L, V, seed=3, skew=skew, dup_frac=0.05)
    d2s = syn.codes_to_docid_to_smtid(codes)
    lst = ob.build_list_smtid_to_nextids(d2s)
    tr = DocidTrie.from_docid_to_smtid(d2s, V)
    assert tr.level_counts() == [len(d) for d in lst]         # evaluate.py:425-426 "effective smtid" figures
    orc = ob.TrieMaskOracle(lst, V)
    rng = np.random.default_rng(0)
    for T in range(1, L + 1):
        rows = codes[rng.integers(0, n, 200)][:, : T - 1].astype(np.int64)
        rnd = rng.integers(0, V, size=(50, T - 1))
        ids = np.concatenate([rows, rnd], 0)
        ids = torch.from_numpy(np.concatenate([np.zeros((len(ids), 1), np.int64), ids], 1))
        assert torch.equal(tr.mask(ids), orc(ids, None)), (n, L, V, T)
        proc = PrefixConstrainLogitProcessorFastSparse.from_trie(tr)
        assert torch.equal(proc(ids, None), orc(ids, None))
    s2d = ob.build_smtid_to_docids(d2s, L)
    for sid, docs in list(s2d.items())[:300]:
        leaf = tr.find_leaf([int(x) for x in sid.split("_")])
        assert tr.docids_for_range(leaf, leaf + 1) == docs     # json order inside an smtid (evaluate.py:439-446)
    assert tr.find_leaf([V - 1] * L) in (-1, tr.find_leaf([V - 1] * L))
    tr2 = DocidTrie.from_list_smtid_to_nextids(lst, V)         # the reference's pickle format
    assert tr2.level_counts() == tr.level_counts() and tr2.n_unique == tr.n_unique


def test_trie_cache_roundtrip_and_errors(tmp_path):
    codes = syn.make_codes(5000, 6, 256, seed=9)
    tr = DocidTrie.from_codes(codes, 256)
    p = str(tmp_path / "trie.bin")
    tr.save(p)
    tr2 = DocidTrie.load(p)
    assert tr2.level_counts() == tr.level_counts()
    ids = torch.zeros((4, 3), dtype=torch.int64)
    ids[:, 1:] = torch.from_numpy(codes[:4, :2].astype(np.int64))
    assert torch.equal(tr.mask(ids), tr2.mask(ids))
    with pytest.raises(ValueError):
        DocidTrie.from_codes(np.full((3, 2), 300), 256)
    with pytest.raises(_lib.RB200Error):
        DocidTrie.load(str(tmp_path / "missing.bin"))
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"not a trie")
    with pytest.raises(_lib.RB200Error):
        DocidTrie.load(str(bad))
    with pytest.raises(ValueError):
        tr.mask(torch.zeros((1, 8), dtype=torch.int64))         # prefix longer than L


def test_generate_argument_errors_match_reference():
    tr = DocidTrie.from_codes(syn.make_codes(100, 4, 16), 16)
    proc = PrefixConstrainLogitProcessorFastSparse.from_trie(tr)
    ids = torch.zeros((2, 5), dtype=torch.long)
    with pytest.raises(ValueError, match="num_return_sequences"):     # generation.py:216-217
        generate_for_constrained_prefix_beam_search(None, proc, input_ids=ids, attention_mask=ids, max_new_tokens=4,
                                                    num_beams=2, num_return_sequences=3)
    with pytest.raises(ValueError):
        PrefixConstrainLogitProcessorFastSparse([], 16)


@pytest.mark.parametrize("name", ["tiny_plain", "tiny_logsoftmax", "tiny_shared_scaleup", "tiny_few_docs", "c1_t5base"])
def test_restated_oracle_reproduces_golden(name):
    """Golden vectors come from the reference's literal loop (oracle/make_golden.py); the self-contained
    restatement must give the same DocIDs and scores without /root/reference."""
    c, dims, w, codes, ids, mask, g = helpers.load_golden(name)
    seqs, scores, _, enc = helpers.oracle_cached_search(w, dims, codes, ids, mask, c["nb"], c["L"], c["log_softmax"])
    assert np.allclose(enc.numpy(), g["encoder_states"], atol=1e-5)
    ref_seq, ref_sc = torch.from_numpy(g["sequences"]), torch.from_numpy(g["sequences_scores"])
    n_mismatch = helpers.compare_ranked(seqs, scores, ref_seq, ref_sc, c["nb"], atol=1e-5)
    assert n_mismatch == 0
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    assert [len(d) for d in lst] == g["level_counts"].tolist()


def test_no_cpu_fallback_exists():
    from ripor_b200.modeling import T5SeqAQEncoder
    dims = syn.T5Dims.tiny()
    model = T5SeqAQEncoder.from_weights(dims, syn.make_weights(dims))
    with pytest.raises(_lib.RB200Error):
        model.to("cpu")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.RB200Error):
            model.base_model.get_engine(1, 2, 8)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument checks come before any CUDA call: they must answer (not crash) on a box without a GPU."""
    import ctypes as C
    L = _lib.lib()
    us = C.c_double()
    assert L.rb200_gemm_bench(9, 128, 128, 64, 0, 1, 0, C.byref(us), None) != 0      # unknown precision
    assert L.rb200_gemm_bench(5, 128, 128, 64, 7, 1, 0, C.byref(us), None) != 0      # unknown epilogue
    assert L.rb200_gemm_bench(5, 0, 128, 64, 0, 1, 0, C.byref(us), None) != 0        # empty problem
    assert b"precision" in L.rb200_last_error() or b"bad argument" in L.rb200_last_error()
    assert L.rb200_engine_last_tail_step(None) == -1


def test_round2_entry_points_validate_arguments_without_a_gpu():
    """The entry points added in round 2 answer bad arguments with a status and a message, never a crash."""
    import ctypes as C
    import numpy as np
    L = _lib.lib()
    codes = syn.make_codes(50, 4, 8, seed=1)
    tr = DocidTrie.from_codes(codes, 8)
    ranges = np.zeros((2, 2), np.int32)
    docs, counts = np.zeros((2, 4), np.int64), np.zeros(2, np.int32)
    # leaf expansion needs an uploaded trie and a sane width
    assert L.rb200_trie_leaf_expand(tr.handle, ranges.ctypes.data, 2, 4, docs.ctypes.data, counts.ctypes.data, None) == -4
    assert b"not uploaded" in L.rb200_last_error()
    assert L.rb200_trie_leaf_expand(tr.handle, ranges.ctypes.data, 2, 0, docs.ctypes.data, counts.ctypes.data, None) == -1
    assert L.rb200_trie_leaf_expand(None, None, 0, 4, None, None, None) == -1
    # engine entry points with null handles
    assert L.rb200_engine_resize(None, 1, 1, 1) == -1
    assert L.rb200_engine_forward(None, None, None, 1, 1, 1, None, 1, None, None, None, None) == -1
    assert L.rb200_engine_last_freeze_histogram(None, None, 0, None) == -1
    nx = C.c_void_p()
    assert L.rb200_engine_next_input(None, C.byref(nx)) == -1
    assert L.rb200_beam_forced_tail(None, None, 1, None, 0, None) == -1
    assert L.rb200_beam_reset_beams(None, None, 1, 1, None) == -1
    # readers
    tab = C.c_void_p()
    assert L.rb200_docid_json_open(b"/nonexistent/docid_to_smtid.json", 0, C.byref(tab)) == -3
    assert L.rb200_docid_json_open(None, 0, C.byref(tab)) == -1
    assert L.rb200_unpack_codes(None, 1, 1, 1, 8, None) == -1
    out = np.zeros(4, np.int32)
    packed = np.zeros(4, np.uint8)
    assert L.rb200_unpack_codes(packed.ctypes.data, 1, 4, 4, 30, out.ctypes.data) == -1       # bits > 24
    # a beam wider than the select kernel's shared-memory budget is refused at creation, with a message
    h = C.c_void_p()
    assert L.rb200_beam_create(0, 1, 4096, 4, 256, C.byref(h)) == -1
    assert b"exceeds the beam kernels" in L.rb200_last_error()
