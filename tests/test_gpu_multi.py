"""Two-GPU test of the one collective on the path: the CLI's --gather mode (NCCL all_gather of the packed per-rank
runs, rank 0 writes run.json directly) against the single-process run and the file-merge task. Skipped on a box with
fewer than two GPUs (the driver's scaling bench exercises the same collective through bench.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from ripor_b200 import evaluate as ev
from tests.test_gpu_cli import _setup, _tok

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys
sys.path.insert(0, {root!r})
from ripor_b200 import evaluate as ev
from tests.test_gpu_cli import _tok
args = ev.get_args(sys.argv[1:])
ev.t5seq_aq_retrieve_docids(args, tokenizer=_tok)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cli_gather_over_nccl_equals_single_process_run(tmp_path):
    dims, w, codes, d2s, model_dir, exp, qdir = _setup(tmp_path, n_docs=800, n_q=9)      # 9 queries: rank 1 gets a pad
    nb, L = 5, 8
    common = ["--task", "t5seq_aq_retrieve_docids", "--pretrained_path", str(model_dir), "--docid_to_smtid_path",
              str(exp / "docid_to_smtid.json"), "--q_collection_paths", json.dumps([str(qdir) + "/"]), "--batch_size", "2",
              "--max_new_token_for_docid", str(L), "--topk", str(nb)]
    single = tmp_path / "single"
    ev.t5seq_aq_retrieve_docids(ev.get_args(common + ["--out_dir", str(single), "--local_rank", "0"]), tokenizer=_tok)
    ref = json.load(open(single / "TOY" / "run_0.json"))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    for mode, extra in (("gather", ["--gather"]), ("files", [])):
        out = tmp_path / mode
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29641" if mode == "gather" else "29642", str(script)] + common + \
              ["--out_dir", str(out)] + extra
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert res.returncode == 0, res.stderr[-3000:]
        if mode == "files":
            assert sorted(os.listdir(out / "TOY")) == ["run_0.json", "run_1.json"]
            ev.t5seq_aq_retrieve_docids_2(ev.get_args(["--task", "t5seq_aq_retrieve_docids_2", "--out_dir", str(out),
                                                       "--q_collection_paths", json.dumps([str(qdir) + "/"]),
                                                       "--num_ranks", "2"]))
        else:
            assert os.listdir(out / "TOY") == ["run.json"]
        run = json.load(open(out / "TOY" / "run.json"))
        assert sorted(run) == sorted(ref)
        for q in ref:
            assert list(run[q]) == list(ref[q]), (mode, q)                # same documents in the same rank order
            for d in ref[q]:
                assert abs(run[q][d] - ref[q][d]) < 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_trie_handle_serves_two_devices():
    """A single-process multi-GPU caller uploads one trie handle to every device: masks,
    searches and leaf expansions on cuda:1 use cuda:1's copy of the tables."""
    import numpy as np
    from ripor_b200 import synthetic as syn
    from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search
    from ripor_b200.modeling import T5SeqAQEncoder
    from ripor_b200.trie import DocidTrie
    from tests import helpers
    L, nb, B = 8, 5, 4
    dims = syn.T5Dims.tiny(docid_len=L)
    w = syn.make_weights(dims)
    V = dims.decoder_vocab_size
    codes = syn.make_codes(3000, L, V, dup_frac=0.05)
    ids, mask = syn.make_queries(B, S=14, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    trie = DocidTrie.from_codes(codes, V)
    proc = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
    pre = torch.from_numpy(np.concatenate([np.zeros((6, 1), np.int64), codes[:6, :3].astype(np.int64)], 1))
    host_mask = trie.mask(pre)
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        assert torch.equal(trie.mask(pre.to(dev)).cpu(), host_mask)
        model = T5SeqAQEncoder.from_weights(dims, w).to(dev)
        out = generate_for_constrained_prefix_beam_search(
            model.base_model, proc, input_ids=ids.to(dev), attention_mask=mask.to(dev), max_new_tokens=L, num_beams=nb,
            num_return_sequences=nb, output_scores=True, return_dict_in_generate=True, precision="tf32x3")
        assert out.sequences.device == torch.device(dev)
        assert helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3) == 0
        docs, counts = trie.expand_ranges(out.leaf_ranges, 4)
        lo, hi = out.leaf_ranges[0].tolist()
        assert docs[0, : int(counts[0])].tolist() == trie.rows_for_range(lo, hi).tolist()
