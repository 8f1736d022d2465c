"""GPU test of the drop-in CLI: the mirrored evaluate.py tasks write the reference's files with the reference's
contents (oracle = reference algorithm on CPU + reference host mapping, evaluate.py:116-128 / :163-174)."""
import json
import os

import pytest
import torch

from oracle import beam as ob
from ripor_b200 import evaluate as ev, synthetic as syn
from ripor_b200.modeling import T5SeqAQEncoder
from tests import helpers

pytestmark = pytest.mark.gpu


def _tok(text):
    return [3 + (sum(map(ord, w)) % 400) for w in text.split()] + [1]


def _setup(tmp_path, n_docs=500, n_q=7):
    dims = syn.T5Dims.tiny(docid_len=8, decoder_vocab_size=16)
    w = syn.make_weights(dims)
    model_dir = tmp_path / "ckpt"
    T5SeqAQEncoder.from_weights(dims, w).save_pretrained(str(model_dir))
    codes = syn.make_codes(n_docs, 8, 16, dup_frac=0.1)
    d2s = {f"D{i}": [-1] + [int(x) for x in row] for i, row in enumerate(codes)}
    exp = tmp_path / "exp"
    exp.mkdir()
    json.dump(d2s, open(exp / "docid_to_smtid.json", "w"))
    qdir = tmp_path / "toy_queries"
    qdir.mkdir()
    words = ["alpha", "beta", "gamma", "delta", "trie", "beam", "search", "docid"]
    lines = [f"{100 + i}\t" + " ".join(words[(i * 3 + j) % 8] for j in range(3 + i % 4)) for i in range(n_q)]
    (qdir / "raw.tsv").write_text("\n".join(lines) + "\n")
    return dims, w, codes, d2s, model_dir, exp, qdir


def _oracle_outputs(dims, w, codes, qdir, nb, L):
    ds = ev.CollectionDatasetWithDocIDPreLoad(str(qdir), "row_id", add_prefix=True, is_query=True)
    loader = ev.CollectionDataWithDocIDLoader(ds, batch_size=len(ds), tokenizer=_tok)
    batch = next(iter(loader))
    seqs, scores, _, _ = helpers.oracle_cached_search(w, dims, codes, batch["input_ids"], batch["attention_mask"], nb, L)
    return batch["id"].tolist(), seqs, scores


def test_retrieve_docids_cli_matches_reference_mapping(tmp_path):
    dims, w, codes, d2s, model_dir, exp, qdir = _setup(tmp_path)
    nb, L = 5, 8
    out = tmp_path / "out"
    args = ev.get_args(["--task", "t5seq_aq_retrieve_docids", "--pretrained_path", str(model_dir),
                        "--docid_to_smtid_path", str(exp / "docid_to_smtid.json"), "--q_collection_paths",
                        json.dumps([str(qdir) + "/"]), "--batch_size", "3", "--max_new_token_for_docid", str(L),
                        "--topk", str(nb), "--out_dir", str(out), "--local_rank", "0"])
    ev.t5seq_aq_retrieve_docids(args, tokenizer=_tok)
    run = json.load(open(out / "TOY" / "run_0.json"))
    # oracle: batches of 3 pad to the longest row of each batch, exactly like the loader
    ds = ev.CollectionDatasetWithDocIDPreLoad(str(qdir), "row_id", add_prefix=True, is_query=True)
    s2d = ob.build_smtid_to_docids(d2s, L)
    gold = {}
    for batch in ev.CollectionDataWithDocIDLoader(ds, batch_size=3, tokenizer=_tok):
        seqs, scores, _, _ = helpers.oracle_cached_search(w, dims, codes, batch["input_ids"], batch["attention_mask"], nb, L)
        gold.update(ob.rankdata_for_batch(batch["id"].tolist(), seqs, scores, s2d, nb, L))
    assert sorted(run) == sorted(str(q) for q in gold)
    for q, r in gold.items():
        assert list(run[str(q)].keys()) == list(r.keys()), q          # same docids in the same order
        for d in r:
            assert abs(run[str(q)][d] - r[d]) < 1e-3 * L
    # merge task
    args2 = ev.get_args(["--task", "t5seq_aq_retrieve_docids_2", "--out_dir", str(out), "--q_collection_paths",
                         json.dumps([str(qdir) + "/"]), "--num_ranks", "1"])
    ev.t5seq_aq_retrieve_docids_2(args2)
    assert json.load(open(out / "TOY" / "run.json")) == run
    assert not os.path.exists(out / "TOY" / "run_0.json")


def test_smtid_rankdata_cli_prefix_search(tmp_path):
    """t5seq_aq_get_qid_to_smtid_rankdata: beams over DocID prefixes (max_new_token 4 < L 8) on the full trie."""
    dims, w, codes, d2s, model_dir, exp, qdir = _setup(tmp_path, n_docs=300, n_q=4)
    nb, Lp = 6, 4
    out = tmp_path / "rank"
    args = ev.get_args(["--task", "t5seq_aq_get_qid_to_smtid_rankdata", "--pretrained_path", str(model_dir),
                        "--docid_to_smtid_path", str(exp / "docid_to_smtid.json"), "--train_query_dir", str(qdir),
                        "--batch_size", "4", "--max_new_token", str(Lp), "--topk", str(nb), "--out_dir", str(out),
                        "--local_rank", "0"])
    got = ev.t5seq_aq_get_qid_to_smtid_rankdata(args, tokenizer=_tok)
    on_disk = json.load(open(out / "qid_smtid_rankdata_0.json"))
    # oracle with the full-length trie mask but only Lp steps
    ds = ev.CollectionDatasetWithDocIDPreLoad(str(qdir), "row_id", add_prefix=True, is_query=True)
    batch = next(iter(ev.CollectionDataWithDocIDLoader(ds, batch_size=4, tokenizer=_tok)))
    import oracle.t5_math as t5_math
    lst = ob.build_list_smtid_to_nextids(d2s)
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, batch["input_ids"], batch["attention_mask"])
        dec = t5_math.CachedDecoder(w, dims, enc, batch["attention_mask"], nb)

        def step(dec_ids, bi):
            if bi is not None:
                dec.reorder(bi)
            return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])
        seqs, scores = ob.beam_search_oracle(step, ob.TrieMaskOracle(lst, 16), 4, nb, Lp)
    s2d = ob.build_smtid_to_docids(d2s, Lp)
    strs = ob.convert_ptsmtids_to_strsmtid(seqs.view(-1, nb, Lp + 1), Lp)
    sc = scores.view(-1, nb).tolist()
    for qi, qid in enumerate(batch["id"].tolist()):
        assert list(got[qid].keys()) == strs[qi]
        assert list(on_disk[str(qid)].keys()) == strs[qi]
        for smtid, s in zip(strs[qi], sc[qi]):
            ref_docs = s2d.get(smtid, [])
            assert list(got[qid][smtid].keys()) == ref_docs             # json order of the docids of a prefix
            for d in ref_docs:
                assert abs(got[qid][smtid][d] - s * Lp) < 1e-3 * Lp
