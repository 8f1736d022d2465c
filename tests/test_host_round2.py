"""CPU tests of the round-2 host pieces: the streaming docid_to_smtid.json reader, faiss bitstring unpacking, the
tagged trie cache, the independent sorted-codes mask oracle, the three-way allowed-token check against the reference's
legacy Trie, the packed collectives (gloo, world size 2), the prefetching query feed and MRR@10."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

from oracle import beam as ob, formats as ofmt, ref_literal
from oracle.range_mask import SortedCodesMask
from ripor_b200 import _lib, evaluate as ev, synthetic as syn
from ripor_b200.trie import DocidTrie, source_tag

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------
# on-disk inputs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("style", ["compact", "spaced", "indented"])
def test_json_reader_equals_json_load(tmp_path, style):
    codes = syn.make_codes(3000, 8, 1024, seed=4, dup_frac=0.05)
    d = {("d%d" % i if i % 11 else 'we"ird\\id\n%d' % i): [-1] + [int(x) for x in r] for i, r in enumerate(codes)}
    p = str(tmp_path / "docid_to_smtid.json")
    with open(p, "w") as f:
        if style == "compact":
            json.dump(d, f, separators=(",", ":"))          # what ujson.dump writes
        elif style == "spaced":
            json.dump(d, f)
        else:
            json.dump(d, f, indent=2)
    ref_ids, ref_codes = ofmt.load_docid_to_smtid(p)
    got_codes, got_ids = DocidTrie.read_json_codes(p)
    assert np.array_equal(got_codes, ref_codes) and list(got_ids) == ref_ids
    assert got_ids[-1] == ref_ids[-1] and len(got_ids) == len(ref_ids)
    cut_codes, _ = DocidTrie.read_json_codes(p, 5)           # evaluate.py:443 smtids[1:1+max_new_token_for_docid]
    assert np.array_equal(cut_codes, ofmt.load_docid_to_smtid(p, 5)[1])
    t = DocidTrie.from_json(p, 1024)
    t0 = DocidTrie.from_codes(codes, 1024, ref_ids)
    assert (t.n_docs, t.n_unique, t.L) == (t0.n_docs, t0.n_unique, t0.L)
    assert t.level_counts() == t0.level_counts()
    leaf = t.find_leaf(codes[7].tolist())
    assert t.docids_for_range(leaf, leaf + 1) == t0.docids_for_range(leaf, leaf + 1)


@pytest.mark.parametrize("text,needle", [('{"a": [0, 1]}', "must start with -1"), ('{"a": [-1, 1], "b": [-1, 1, 2]}', "different lengths"),
                                         ('{"a": [-1, 1], "b": [-1, x]}', "expected an integer"), ('{"a": [-1, 1]', "end of file"),
                                         ('[1, 2]', "expected '{'"), ('{"a": [-1, -5]}', "negative code"), ('{}', "no documents")])
def test_json_reader_rejects_malformed_input(tmp_path, text, needle):
    p = tmp_path / "bad.json"
    p.write_text(text)
    with pytest.raises(_lib.RB200Error, match=needle):
        DocidTrie.read_json_codes(str(p))
    with pytest.raises(_lib.RB200Error, match="cannot open"):
        DocidTrie.read_json_codes(str(tmp_path / "missing.json"))


def test_json_reader_refuses_codes_beyond_the_vocabulary(tmp_path):
    p = tmp_path / "d.json"
    p.write_text('{"a": [-1, 3, 300]}')
    with pytest.raises(ValueError, match="decoder_vocab_size"):
        DocidTrie.from_json(str(p), 256)


def test_bitstring_unpack_matches_oracle_and_hand_vectors():
    # hand vector: M=3 codes of 11 bits, LSB first: 5 (0b00000000101), 2047, 1024
    bits = [1, 0, 1] + [0] * 8 + [1] * 11 + [0] * 10 + [1]
    by = np.packbits(np.array(bits + [0] * (40 - len(bits)), np.uint8), bitorder="little")[None, :]
    assert ofmt.unpack_bitstring_codes(by, 3, 11).tolist() == [[5, 2047, 1024]]
    rng = np.random.default_rng(0)
    for M, nbits in ((8, 11), (16, 10), (32, 8), (4, 12), (24, 5), (7, 13)):
        code_size = (M * nbits + 7) // 8 + (M % 2)            # faiss rounds code_size up; spare bytes are ignored
        packed = rng.integers(0, 256, size=(257, code_size), dtype=np.uint8)
        want = ofmt.unpack_bitstring_codes(packed, M, nbits)
        got = np.zeros((257, M), np.int32)
        _lib.check(_lib.lib().rb200_unpack_codes(packed.ctypes.data, 257, code_size, M, nbits, got.ctypes.data))
        assert np.array_equal(got, want), (M, nbits)
        assert got.max() < (1 << nbits)
    with pytest.raises(ValueError, match="cannot hold"):
        _lib.check(_lib.lib().rb200_unpack_codes(packed.ctypes.data, 1, 2, 8, 11, got.ctypes.data))


def test_trie_cache_is_tagged_with_its_source(tmp_path):
    codes = syn.make_codes(500, 6, 16, seed=1)
    d = syn.codes_to_docid_to_smtid(codes)
    src = tmp_path / "docid_to_smtid.json"
    src.write_text(json.dumps(d))
    tag = source_tag(str(src))
    t = DocidTrie.from_json(str(src), 16)
    cache = str(tmp_path / "docid_trie.rb200")
    t.save(cache, tag)
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]          # written under a temporary name, renamed
    t2 = DocidTrie.load(cache, t.docids, expect_tag=tag)
    assert t2.level_counts() == t.level_counts()
    # the json is regenerated in place (other codes, same size): the stale cache must not be used
    time.sleep(0.01)
    d2 = syn.codes_to_docid_to_smtid(syn.make_codes(500, 6, 16, seed=2))
    src.write_text(json.dumps(d2))
    assert source_tag(str(src)) != tag
    with pytest.raises(ValueError, match="another docid_to_smtid.json"):
        DocidTrie.load(cache, t.docids, expect_tag=source_tag(str(src)))
    with pytest.raises(ValueError, match="documents"):
        DocidTrie.load(cache, list(t.docids)[:-1], expect_tag=tag)
    # load_docid_trie rebuilds instead of trusting it (cache only under "experiments-full", like the pickle)
    full = tmp_path / "experiments-full"
    full.mkdir()
    src2 = full / "docid_to_smtid.json"
    src2.write_text(json.dumps(d))
    a = ev.load_docid_trie(str(src2), 16)
    assert os.path.exists(full / "docid_trie.rb200")
    b = ev.load_docid_trie(str(src2), 16)                                 # from the cache
    assert b.level_counts() == a.level_counts() and b.docids[3] == a.docids[3]
    src2.write_text(json.dumps(d2))
    c = ev.load_docid_trie(str(src2), 16)                                 # stale cache ignored and replaced
    assert c.level_counts() == DocidTrie.from_docid_to_smtid(d2, 16).level_counts()


# ---------------------------------------------------------------------------------------------------
# masks: independent oracle, and the reference's legacy Trie as a third voice
# ---------------------------------------------------------------------------------------------------
def _prefix_batches(codes, V, rng, n_rows=48, n_junk=16):
    n, L = codes.shape
    for t in range(L):
        rows = codes[rng.integers(0, n, n_rows), :t].astype(np.int64)
        junk = rng.integers(0, V, size=(n_junk, t))
        pre = np.concatenate([rows, junk])
        yield t, torch.from_numpy(np.concatenate([np.zeros((len(pre), 1), np.int64), pre], 1))


@pytest.mark.parametrize("n,L,V,skew", [(300, 5, 8, False), (5000, 6, 256, True), (4000, 4, 1024, False),
                                        (20000, 7, 256, False)])
def test_sorted_codes_mask_equals_dict_oracle_and_host_trie(n, L, V, skew):
    codes = syn.make_codes(n, L, V, seed=3, skew=skew, dup_frac=0.05)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    a, b, c = ob.TrieMaskOracle(lst, V), SortedCodesMask(codes, V), DocidTrie.from_codes(codes, V)
    for t, ids in _prefix_batches(codes, V, np.random.default_rng(0)):
        ma = a(ids)
        assert torch.equal(ma, b(ids)), t
        assert torch.equal(ma, c.mask(ids)), t


@pytest.mark.reference
@pytest.mark.parametrize("seed,skew", [(1, False), (2, True)])
def test_three_way_allowed_tokens_with_legacy_trie(seed, skew):
    """Same allowed-token sets from (1) the reference's sparse processor, (2) the reference's legacy nested-dict Trie
    (utils/generation_utils.py:9-124; -inf mask, eos fallback for an empty set) and (3) the product's flattened trie."""
    legacy = ref_literal.load_reference_legacy_trie()
    V, L = 16, 5
    codes = syn.make_codes(3000, L, V, seed=seed, skew=skew, dup_frac=0.05)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    sparse = ref_literal.literal_processor(lst, V)
    eos = V                                              # one column past the codebook: never a real code
    old = legacy.PrefixConstrainedLogitsProcessorForSmtidTree([[int(x) for x in r] for r in codes], eos)
    mine = DocidTrie.from_codes(codes, V)
    for t, ids in _prefix_batches(codes, V, np.random.default_rng(seed)):
        m1 = sparse(ids, None)
        m3 = mine.mask(ids)
        out = old(ids[:, 1:], torch.zeros((ids.shape[0], V + 1)))        # legacy prefixes carry no start token
        m2 = torch.isfinite(out[:, :V]).double()
        empty = torch.isfinite(out[:, V])                                # eos fallback <=> empty allowed set
        assert torch.equal(m1, m3) and torch.equal(m1, m2), t
        assert torch.equal(empty, m1.sum(1) == 0), t


# ---------------------------------------------------------------------------------------------------
# collectives, feed, metric
# ---------------------------------------------------------------------------------------------------
def test_pack_unpack_run_roundtrip():
    run = {7: {"d3": 1.5, "d1": -2.25}, 9: {}, 11: {"x": 0.1}}
    back = ev.unpack_run(ev.pack_run(run))
    assert back == run and list(back[7]) == ["d3", "d1"]
    assert ev.unpack_run(ev.pack_run({})) == {}


WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from ripor_b200 import evaluate as ev
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
n, nb = 5, 3
idx = ev.distributed_sampler_indices(n, 2, rank)
qids = torch.tensor(idx)
rows = torch.stack([qids * 100 + j for j in range(nb)], 1)
scores = torch.stack([-(qids.float() + 0.25 * j) for j in range(nb)], 1)
gq, gr, gs = ev.gather_ranked_lists(qids, rows, scores)
keep = ev.unique_queries(gq, n)
print("LISTS", json.dumps([gq[keep].tolist(), gr[keep].tolist(), gs[keep].tolist()]))
local = {{str(q): {{f"doc{{q}}_{{rank}}": float(q), "shared": 1.0}} for q in idx}}
merged = ev.gather_runs(local)
if rank == 0:
    print("MERGED", json.dumps(merged, sort_keys=True))
dist.destroy_process_group()
"""


def test_gloo_world2_packed_gathers(tmp_path):
    """The N>1 exchange on CPU: two ranks shard 5 queries like DistributedSampler, all-gather their packed ranked lists
    (every rank gets all of them, in query order, padding dropped) and gather their run dicts to rank 0."""
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, port=29621))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    for out, _ in outs:
        gq, gr, gs = json.loads([l for l in out.splitlines() if l.startswith("LISTS")][0][6:])
        assert gq == [0, 1, 2, 3, 4]
        assert gr == [[q * 100 + j for j in range(3)] for q in range(5)]
        assert gs == [[-(q + 0.25 * j) for j in range(3)] for q in range(5)]
    merged = json.loads([l for l in outs[0][0].splitlines() if l.startswith("MERGED")][0][7:])
    assert sorted(merged) == ["0", "1", "2", "3", "4"]
    assert merged["0"] == {"doc0_0": 0.0, "doc0_1": 0.0, "shared": 1.0}    # padded duplicate collapses by dict update
    assert merged["3"] == {"doc3_1": 3.0, "shared": 1.0}


def test_prefetch_loader_keeps_order_and_surfaces_errors():
    batches = [{"input_ids": torch.full((2, 3), i), "attention_mask": torch.ones(2, 3, dtype=torch.long),
                "id": torch.tensor([2 * i, 2 * i + 1])} for i in range(7)]
    got = list(ev.PrefetchLoader(batches, depth=2))
    assert len(got) == 7 and all(torch.equal(a["input_ids"], b["input_ids"]) for a, b in zip(got, batches))

    def broken():
        yield batches[0]
        raise RuntimeError("tokenizer blew up")

    class It:
        def __iter__(self):
            return broken()

        def __len__(self):
            return 2
    with pytest.raises(RuntimeError, match="tokenizer blew up"):
        list(ev.PrefetchLoader(It(), depth=2))
    # the consumer may stop early without hanging the worker
    it = iter(ev.PrefetchLoader(batches, depth=1))
    next(it)
    del it


def test_mrr_matches_truncate_run_plus_trec_eval_semantics():
    # ties: all documents of one smtid get the same score (evaluate.py:118-128); trec_eval ranks equal scores by
    # docid descending. Queries outside the qrel do not count; queries without a run entry do not count either.
    run = {"q1": {"a": 2.0, "b": 2.0, "c": 1.0}, "q2": {"x": 1.0}, "q3": {"y": 5.0}}
    qrel = {"q1": {"a": 1}, "q2": {"x": 1}, "q4": {"z": 1}}
    # q1: tie between a and b -> b first (descending docid), a at rank 2 -> 0.5; q2 -> 1.0; q3 not in the qrel
    assert ev.mrr_k(run, qrel, 10) == pytest.approx(0.75)
    # truncation happens BEFORE the tie-break: with k=1 the stable sort keeps "a" (inserted first)
    assert ev.mrr_k({"q1": {"a": 2.0, "b": 2.0}}, {"q1": {"b": 1}}, 1) == 0.0
    assert ev.mrr_k({}, qrel, 10) == 0.0


# ------------------------------------------------------------------------------------------------
# the fp32 pre-filter of the warp beam kernel (csrc/beam.cu): the threshold rule, restated in numpy
# ------------------------------------------------------------------------------------------------
def _f32_round_down(v):
    """float64 -> the largest float32 <= v (what __double2float_rd returns)."""
    with np.errstate(over="ignore"):
        f = np.float32(v)
    if np.isnan(f):
        return f
    if np.float64(f) > v:
        f = np.nextafter(f, np.float32(-np.inf))
    return f


def _thresholds(tau, bs):
    with np.errstate(invalid="ignore", over="ignore"):
        dlt = np.float64(tau) - np.float64(bs)
        mag = abs(np.float64(tau)) + abs(np.float64(bs))
        t_ok = _f32_round_down(dlt - mag * 2.0 ** -48)
        t_pen = _f32_round_down((dlt + 1e9) - (mag + 1e9) * 2.0 ** -47)
    return t_ok, t_pen


def _value(x, ok, bs):
    with np.errstate(invalid="ignore", over="ignore"):
        processed = np.float64(x) if ok else np.float64(x) + np.float64(-1e9)
        val = processed + np.float64(bs)
    return val if val == val else -1.7976931348623157e308


def test_beam_prefilter_threshold_rule_never_drops_a_candidate():
    """beam_step_warp_kernel skips a candidate when its fp32 logit x is below a per-(beam, class) threshold. The rule
    must be one-sided: x < threshold implies value(x) < tau in the kernel's float64 arithmetic, for every tau and beam
    score (also across the 1e9 penalty, at exact ties and next to the threshold itself); non-finite thresholds must
    fail the comparison. Exhaustive neighbourhood walk around the threshold for seeded (tau, score) pairs."""
    rng = np.random.default_rng(0)
    scales = [1e-30, 1e-6, 1.0, 37.5, 1e4, 1e9, 3e9, 1e15, 1e30]
    cases = [(0.0, 0.0), (0.0, -1e9), (-1e9, 0.0), (5.0, 5.0), (-2e9, -1e9), (1.0, -3.0)]
    for _ in range(400):
        s1, s2 = rng.choice(scales, 2)
        cases.append((float(rng.standard_normal() * s1), float(rng.standard_normal() * s2)))
    for v in (np.inf, -np.inf, np.nan, 1.7e308, -1.7976931348623157e308):
        cases.append((v, 1.0))
    checked = 0
    for tau, bs in cases:
        t_ok, t_pen = _thresholds(tau, bs)
        for ok, thr in ((True, t_ok), (False, t_pen)):
            if not np.isfinite(thr):
                # -inf / NaN / +inf thresholds: `x < thr` is false for every finite x below them only if thr is -inf or
                # NaN; a +inf threshold may only arise when no finite x can reach tau
                if thr == np.inf:
                    assert _value(np.finfo(np.float32).max, ok, bs) < tau
                continue
            x = np.float32(thr)
            for _step in range(64):                       # the 64 floats just below the threshold are all skipped ...
                x = np.nextafter(x, np.float32(-np.inf))
                assert _value(x, ok, bs) < tau, (tau, bs, ok, float(x))
                checked += 1
            for x in (np.float32(thr) - np.float32(abs(float(thr)) * 0.5 + 1.0), np.float32(-3.0e38)):
                if x < thr:
                    assert _value(x, ok, bs) < tau
            # ... and the margin is small: a logit four margins above tau - score (rounded up to fp32) reaches tau
            if abs(tau) < 1e30 and abs(bs) < 1e30:
                mag = abs(np.float64(tau)) + abs(np.float64(bs))
                up = (np.float64(tau) - np.float64(bs)) + 4 * mag * 2.0 ** -48 if ok else \
                    (np.float64(tau) - np.float64(bs) + 1e9) + 4 * (mag + 1e9) * 2.0 ** -47
                xr = np.float32(up)
                if np.float64(xr) < up:
                    xr = np.nextafter(xr, np.float32(np.inf))
                if np.isfinite(xr):
                    assert not (xr < thr) and _value(xr, ok, bs) >= tau, (tau, bs, ok)
    assert checked > 20000
