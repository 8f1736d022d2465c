"""CPU tests of the host-side mirror of evaluate.py: query feed, sampler sharding, merge, gloo gather."""
import json
import os
import subprocess
import sys

import pytest
import torch

from ripor_b200 import evaluate as ev
from ripor_b200.utils import convert_ptsmtids_to_strsmtid, get_dataset_name

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dataset_name_and_smtid_strings():
    assert get_dataset_name("/x/msmarco/TREC_DL_2019/queries_2019/") == "TREC_DL_2019"
    assert get_dataset_name("/x/msmarco/dev_queries/") == "MSMARCO"
    assert get_dataset_name("/x/msmarco/train_queries/") == "MSMARCO_TRAIN"
    assert get_dataset_name("/tmp/toy/q") == "TOY"
    assert get_dataset_name("/tmp/zzz") == "other_dataset"
    t = torch.tensor([[[0, 5, 7], [0, 1, 2]]])
    assert convert_ptsmtids_to_strsmtid(t, 2) == [["5_7", "1_2"]]
    with pytest.raises(AssertionError):
        convert_ptsmtids_to_strsmtid(t, 3)


@pytest.mark.parametrize("n,world", [(10, 4), (7, 2), (3, 8), (16, 4), (1, 2)])
def test_sampler_indices_equal_torch_distributed_sampler(n, world):
    from torch.utils.data.distributed import DistributedSampler
    ds = list(range(n))
    for r in range(world):
        ref = list(DistributedSampler(ds, num_replicas=world, rank=r, shuffle=False))
        assert ev.distributed_sampler_indices(n, world, r) == ref


def test_query_feed_matches_reference_format(tmp_path):
    d = tmp_path / "toy_queries"
    d.mkdir()
    (d / "raw.tsv").write_text("11\twhat is a trie\n12\tbeam search\twith tab\n\n13\tlast one\n")
    ds = ev.CollectionDatasetWithDocIDPreLoad(str(d), "row_id", add_prefix=True, is_query=True)
    assert len(ds) == 3 and ds[0] == ("11", "query: what is a trie", [-1])
    assert ds[1][1] == "query: beam search with tab"
    tok = lambda text: [3 + (hash(w) % 50) for w in text.split()] + [1]
    loader = ev.CollectionDataWithDocIDLoader(ds, batch_size=2, tokenizer=tok, sampler=[0, 1, 3])
    batches = list(loader)
    assert len(batches) == 2 and batches[0]["input_ids"].shape == batches[0]["attention_mask"].shape
    assert batches[0]["id"].tolist() == [11, 12] and batches[1]["id"].tolist() == [13]
    assert batches[0]["attention_mask"].sum(1).tolist() == [6, 6]


def test_merge_rank_runs_and_merge_task(tmp_path):
    out = tmp_path / "out" / "TOY"
    out.mkdir(parents=True)
    json.dump({"1": {"a": 1.0}, "2": {"b": 2.0}}, open(out / "run_0.json", "w"))
    json.dump({"2": {"c": 3.0}, "3": {"d": 4.0}}, open(out / "run_1.json", "w"))
    args = ev.get_args(["--task", "t5seq_aq_retrieve_docids_2", "--out_dir", str(tmp_path / "out"),
                        "--q_collection_paths", json.dumps([str(tmp_path / "toy") + "/"]), "--num_ranks", "2"])
    ev.t5seq_aq_retrieve_docids_2(args)
    merged = json.load(open(out / "run.json"))
    assert merged == {"1": {"a": 1.0}, "2": {"b": 2.0, "c": 3.0}, "3": {"d": 4.0}}
    assert sorted(os.listdir(out)) == ["run.json"]
    assert ev.mrr_k({"q": {"x": 3.0, "y": 2.0}}, {"q": {"y": 1}}) == 0.5
    with pytest.raises(ValueError):
        ev.main(["--task", "index"])


WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch.distributed as dist
from ripor_b200 import evaluate as ev
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
idx = ev.distributed_sampler_indices(5, 2, rank)
local = {{str(q): {{f"doc{{q}}_{{rank}}": float(q)}} for q in idx}}
merged = ev.gather_runs(local)
if rank == 0:
    print("MERGED", json.dumps(merged, sort_keys=True))
dist.destroy_process_group()
"""


def test_gloo_world2_gather_runs(tmp_path):
    """N>1 path on CPU: two ranks shard 5 queries like DistributedSampler and gather their runs to rank 0."""
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, port=29613))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0].splitlines() if l.startswith("MERGED")][0]
    merged = json.loads(line[len("MERGED "):])
    assert sorted(merged) == ["0", "1", "2", "3", "4"]
    assert merged["0"] == {"doc0_0": 0.0, "doc0_1": 0.0}        # padded duplicate of query 0 collapses by dict update
    assert merged["3"] == {"doc3_1": 3.0}
