"""CLI entry point for the retrieval tasks, same flags and files as the reference's evaluate.py.

    python -m ripor_b200.evaluate --task=t5seq_aq_retrieve_docids --pretrained_path P --docid_to_smtid_path J \
        --q_collection_paths '["dir/"]' --batch_size B --max_new_token_for_docid L --topk nb --out_dir O \
        [--apply_log_softmax_for_scores] [--local_rank r]
    python -m ripor_b200.evaluate --task=t5seq_aq_retrieve_docids_2 --out_dir O --q_collection_paths '["dir/"]'

Mirrors reference t5_pretrainer/evaluate.py: ``t5seq_aq_retrieve_docids`` (:396-487), ``constrained_decode_doc``
(:87-132), ``t5seq_aq_retrieve_docids_2`` (:489-526), the query feed ``CollectionDatasetWithDocIDPreLoad`` /
``CollectionDataWithDocIDLoader`` (dataset/dataset.py:266-332, dataset/dataloader.py:62-79) and
``DistributedSampler(shuffle=False)`` sharding (:463-468). Output files are the reference's:
``<out_dir>/<dataset>/run_{local_rank}.json`` = {qid: {docid: score}} and the merged ``run.json``.
When a process group is initialised the per-rank runs can also be gathered over NCCL/gloo
(``gather_runs``) instead of going through the shared file system.
"""
from __future__ import annotations

import argparse
import json
import os
import queue
import threading
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search
from .modeling import T5SeqAQEncoder
from . import _lib
from .trie import DocidTrie, source_tag
from .utils import convert_ptsmtids_to_strsmtid, get_dataset_name

QUERY_PREFIX = "query: "          # reference dataset/dataset.py:15


# ---------------------------------------------------------------------------------------------------
# query feed
# ---------------------------------------------------------------------------------------------------
class CollectionDatasetWithDocIDPreLoad:
    """raw.tsv reader with the reference's row-id access and "query: " prefix (dataset.py:266-332)."""

    def __init__(self, data_dir: str, id_style: str = "row_id", tid_to_smtid_path=None, add_prefix=False,
                 is_query=False):
        assert id_style in ("content_id", "row_id"), "provide valid id_style"
        self.data_dir, self.id_style = data_dir, id_style
        self.data_dict, self.line_dict = {}, {}
        with open(os.path.join(data_dir, "raw.tsv")) as reader:
            for i, line in enumerate(reader):
                if len(line) > 1:
                    id_, *data = line.split("\t")
                    data = " ".join(" ".join(data).splitlines())
                    if id_style == "row_id":
                        self.data_dict[i] = data
                        self.line_dict[i] = id_.strip()
                    else:
                        self.data_dict[id_] = data.strip()
        self.nb_ex = len(self.data_dict)
        self.add_prefix, self.is_query = add_prefix, is_query

    def __len__(self):
        return self.nb_ex

    def __getitem__(self, idx):
        text = self.data_dict[idx]
        if self.add_prefix and self.is_query:
            text = QUERY_PREFIX + text
        return self.line_dict[idx], text, [-1]


def distributed_sampler_indices(n: int, world_size: int, rank: int) -> List[int]:
    """torch DistributedSampler(shuffle=False, drop_last=False): pad by wrapping, then stride by rank."""
    if world_size <= 1:
        return list(range(n))
    total = -(-n // world_size) * world_size
    idx = list(range(n))
    pad = total - n
    if pad > 0:
        idx += (idx * (pad // max(len(idx), 1) + 1))[:pad]
    return idx[rank:total:world_size]


class CollectionDataWithDocIDLoader:
    """Batches of {"input_ids", "attention_mask", "id"}: tokenise, pad to the longest row, truncate to
    max_length (dataloader.py:62-79). ``tokenizer`` is an HF tokenizer (AutoTokenizer.from_pretrained(path)) or any
    callable text -> list[int]; batches are placed in pinned host memory like the reference's pin_memory=True."""

    def __init__(self, dataset, tokenizer_type=None, max_length=256, batch_size=1, sampler: Optional[Sequence[int]] = None,
                 tokenizer: Optional[Callable] = None, **_):
        self.dataset, self.max_length, self.batch_size = dataset, max_length, batch_size
        self.indices = list(sampler) if sampler is not None else list(range(len(dataset)))
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(tokenizer_type)
        self.tokenizer = tokenizer

    def __len__(self):
        return -(-len(self.indices) // self.batch_size)

    def _encode(self, texts: List[str]):
        if hasattr(self.tokenizer, "pad_token_id") or hasattr(self.tokenizer, "batch_encode_plus"):
            enc = self.tokenizer(texts, add_special_tokens=True, padding="longest", truncation="longest_first",
                                 max_length=self.max_length, return_attention_mask=True)
            return torch.tensor(enc["input_ids"]), torch.tensor(enc["attention_mask"])
        rows = [list(self.tokenizer(t))[: self.max_length] for t in texts]
        S = max(len(r) for r in rows)
        ids = torch.zeros((len(rows), S), dtype=torch.long)
        mask = torch.zeros((len(rows), S), dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return ids, mask

    def __iter__(self):
        for s in range(0, len(self.indices), self.batch_size):
            items = [self.dataset[i] for i in self.indices[s: s + self.batch_size]]
            id_, texts, _ = zip(*items)
            ids, mask = self._encode(list(texts))
            pin = torch.cuda.is_available()
            yield {"input_ids": ids.pin_memory() if pin else ids, "attention_mask": mask.pin_memory() if pin else mask,
                   "id": torch.tensor([int(i) for i in id_], dtype=torch.long)}


class PrefetchLoader:
    """Double-buffered query feed: a worker thread tokenises and pins the next batches while the GPU searches the
    current one (the reference gets the same overlap from DataLoader(num_workers=1, pin_memory=True),
    dataloader.py:19, evaluate.py:467). ``device`` set: the H2D copies are issued from the consumer side with
    non_blocking=True out of the pinned buffers."""

    def __init__(self, loader, depth: int = 2, device=None):
        self.loader, self.depth, self.device = loader, max(1, depth), device

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def work():
            try:
                for batch in self.loader:
                    while not stop.is_set():
                        try:
                            q.put(batch, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
                q.put(None)
            except BaseException as exc:          # surfaced on the consumer side
                q.put(exc)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                if self.device is not None:
                    item = {k: (v.to(self.device, non_blocking=True) if k != "id" else v) for k, v in item.items()}
                yield item
        finally:
            stop.set()


# ---------------------------------------------------------------------------------------------------
# decode loop (evaluate.py:87-132)
# ---------------------------------------------------------------------------------------------------
LEAF_EXPAND_WIDTH = 8     # documents per ranked DocID expanded on the device; wider rows fall back to the host table


def _expand_on_device(trie: DocidTrie, outputs, topk: int):
    """(doc rows [B, topk, k] int64, counts [B, topk] int32, leaf ranges [B, topk, 2]) as host arrays; the expansion
    itself (evaluate.py:118-128 smtid -> docids) runs on the device when the search results live there."""
    leaf = outputs.leaf_ranges
    if leaf.is_cuda:
        docs, counts = trie.expand_ranges(leaf, LEAF_EXPAND_WIDTH)
        return (docs.view(-1, topk, LEAF_EXPAND_WIDTH).cpu().numpy(), counts.view(-1, topk).cpu().numpy(),
                leaf.view(-1, topk, 2).cpu().numpy())
    return None, None, leaf.view(-1, topk, 2).numpy()


def _docids_of(trie: DocidTrie, docs, counts, leaf, q: int, j: int) -> List[str]:
    lo, hi = int(leaf[q, j, 0]), int(leaf[q, j, 1])
    if hi <= lo:
        return []
    if counts is not None and counts[q, j] <= LEAF_EXPAND_WIDTH:
        return [trie.docid_of_row(r) for r in docs[q, j, : counts[q, j]]]
    return trie.docids_for_range(lo, hi)


def retrieve_batch(model, prefix_constrain_processor, input_ids, attention_mask, max_new_token: int, topk: int,
                   apply_log_softmax_for_scores: bool = False, width: int = LEAF_EXPAND_WIDTH, precision=None):
    """One batch through the whole path, host to host: pinned (or pageable) ``input_ids`` / ``attention_mask`` [B, S]
    are copied to the model's device, searched, the ranked DocIDs expanded to documents on the device
    (evaluate.py:118-128) and the result copied back. Returns host tensors (doc rows int64 [B, topk, width] in json
    order padded with -1, document counts int32 [B, topk], scores fp32 [B, topk], leaf ranges int32 [B, topk, 2]) and
    the generate output. This is the body of ``constrained_decode_doc``'s loop and what bench.py times as ``e2e``."""
    dev = model.device if hasattr(model, "device") else torch.device("cuda", torch.cuda.current_device())
    ids = input_ids.to(dev, non_blocking=True).long()
    mask = attention_mask.to(dev, non_blocking=True).long()
    out = generate_for_constrained_prefix_beam_search(
        model, prefix_constrain_processor, input_ids=ids, attention_mask=mask, max_new_tokens=max_new_token,
        output_scores=True, return_dict=True, return_dict_in_generate=True, num_beams=topk, num_return_sequences=topk,
        apply_log_softmax_for_scores=apply_log_softmax_for_scores, precision=precision)
    docs, counts = prefix_constrain_processor.trie.expand_ranges(out.leaf_ranges, width)
    B = ids.shape[0]
    res = (docs.view(B, topk, width).to("cpu", non_blocking=True), counts.view(B, topk).to("cpu", non_blocking=True),
           out.sequences_scores.view(B, topk).to("cpu", non_blocking=True),
           out.leaf_ranges.view(B, topk, 2).to("cpu", non_blocking=True))
    torch.cuda.current_stream(dev).synchronize()
    return res, out


def constrained_decode_doc(model, dataloader, prefix_constrain_processor, smtid_to_docids, max_new_token, device,
                           out_dir, local_rank, topk=100, apply_log_softmax_for_scores=False, write=True):
    """smtid_to_docids: the reference's dict {"c1_.._cL": [docids]} or None to expand leaves from the trie."""
    qid_to_rankdata: Dict[int, Dict[str, float]] = {}
    trie: DocidTrie = prefix_constrain_processor.trie
    for batch in dataloader:
        outputs = generate_for_constrained_prefix_beam_search(
            model, prefix_constrain_processor, input_ids=batch["input_ids"].long(),
            attention_mask=batch["attention_mask"].long(), max_new_tokens=max_new_token, output_scores=True,
            return_dict=True, return_dict_in_generate=True, num_beams=topk, num_return_sequences=topk,
            apply_log_softmax_for_scores=apply_log_softmax_for_scores)
        batch_qids = batch["id"].cpu().tolist()
        seqs = outputs.sequences.view(-1, topk, max_new_token + 1)
        relevant_scores = outputs.sequences_scores.view(-1, topk).cpu().tolist()
        if smtid_to_docids is not None:
            str_smtids = convert_ptsmtids_to_strsmtid(seqs, max_new_token)
            for qid, ranked_smtids, rel_scores in zip(batch_qids, str_smtids, relevant_scores):
                qid_to_rankdata[qid] = {}
                for smtid, rel_score in zip(ranked_smtids, rel_scores):
                    if smtid not in smtid_to_docids:
                        print(f"smtid: {smtid} not in smtid_to_docid")
                    else:
                        for docid in smtid_to_docids[smtid]:
                            qid_to_rankdata[qid][docid] = rel_score if apply_log_softmax_for_scores \
                                else rel_score * max_new_token
        else:   # same mapping from the trie's leaf table, no Python dict of 8.8M strings
            docs, counts, leaf = _expand_on_device(trie, outputs, topk)
            for q, (qid, rel_scores) in enumerate(zip(batch_qids, relevant_scores)):
                qid_to_rankdata[qid] = {}
                for j, rel_score in enumerate(rel_scores):
                    docids = _docids_of(trie, docs, counts, leaf, q, j)
                    if not docids:
                        print("smtid not in smtid_to_docid")
                        continue
                    for docid in docids:
                        qid_to_rankdata[qid][docid] = rel_score if apply_log_softmax_for_scores \
                            else rel_score * max_new_token
    if write == "gather":           # collective merge (replaces run_{rank}.json + the _2 task's file merge)
        merged = gather_runs(qid_to_rankdata)
        if merged is not None:
            with open(os.path.join(out_dir, "run.json"), "w") as fout:
                json.dump(merged, fout)
        return merged if merged is not None else qid_to_rankdata
    if write:
        with open(os.path.join(out_dir, f"run_{local_rank}.json"), "w") as fout:
            json.dump(qid_to_rankdata, fout)
    return qid_to_rankdata


def build_list_smtid_to_nextids(docid_to_smtids):
    """evaluate.py:411-424 (kept for callers that want the reference's pickle format)."""
    out = [dict() for _ in range(len(next(iter(docid_to_smtids.values()))) - 1)]
    for _, smtids in docid_to_smtids.items():
        for i in range(len(smtids) - 1):
            out[i].setdefault("_".join(str(x) for x in smtids[: i + 1]), set()).add(int(smtids[i + 1]))
    return [{k: list(v) for k, v in d.items()} for d in out]


def ddp_setup():
    """reference evaluate.py:181-182 (init_process_group("nccl")); one process per GPU, the device follows LOCAL_RANK."""
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo")


def load_docid_trie(docid_to_smtid_path: str, V: int, local_rank: int = 0, use_cache: bool = True) -> DocidTrie:
    """docid_to_smtid.json -> DocidTrie (reference evaluate.py:400-446 builds three Python dict structures instead).
    The file goes through the streaming C++ reader; the flattened trie is cached next to it as ``docid_trie.rb200``
    (the counterpart of the reference's list_smtid_to_nextids.pkl, :404-432). The cache header carries a tag of the
    json it was made from and is rebuilt when that does not match; it is written atomically by rank 0 only."""
    cache = os.path.join(os.path.dirname(docid_to_smtid_path), "docid_trie.rb200")
    tag = source_tag(docid_to_smtid_path)
    dist = torch.distributed
    if use_cache and os.path.exists(cache):
        codes, docids = DocidTrie.read_json_codes(docid_to_smtid_path)
        try:
            trie = DocidTrie.load(cache, docids, expect_tag=tag)
            if trie.L == codes.shape[1] and trie.V == V:
                print("read flattened trie from {}".format(cache))
                return trie
        except (ValueError, _lib.RB200Error) as err:
            print(f"ignoring {cache}: {err}")
    trie = DocidTrie.from_json(docid_to_smtid_path, V)
    if local_rank <= 0:
        for i, n in enumerate(trie.level_counts()):
            print(f"{i}-th step has {n:,} effective smtid ")
        if use_cache and "experiments-full" in docid_to_smtid_path:       # same condition as the pickle (:405,428)
            trie.save(cache, tag)
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    return trie


def t5seq_aq_retrieve_docids(args, tokenizer=None):
    ddp_setup()
    model = T5SeqAQEncoder.from_pretrained(args.pretrained_path)
    model.eval()
    print(args.docid_to_smtid_path)
    V = model.config.decoder_vocab_sizes
    if len(set(V)) != 1:
        raise ValueError("not valid decoder_vocab_size")
    trie = load_docid_trie(args.docid_to_smtid_path, V[0], args.local_rank)
    prefix_constrain_processor = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
    max_new_token = args.max_new_token_for_docid
    assert trie.L >= max_new_token, (trie.L, max_new_token)
    if args.local_rank <= 0:
        print("max_new_token: ", max_new_token)
        os.makedirs(args.out_dir, exist_ok=True)
    q_paths = args.q_collection_paths
    if len(q_paths) == 1 and q_paths[0].lstrip().startswith("["):
        q_paths = json.loads(q_paths[0])
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    local_rank = max(args.local_rank, 0)
    for data_dir in q_paths:
        dev_dataset = CollectionDatasetWithDocIDPreLoad(data_dir=data_dir, id_style="row_id", tid_to_smtid_path=None,
                                                        add_prefix=True, is_query=True)
        dev_loader = CollectionDataWithDocIDLoader(dataset=dev_dataset, tokenizer_type=args.pretrained_path,
                                                   max_length=256, batch_size=args.batch_size,
                                                   sampler=distributed_sampler_indices(len(dev_dataset), world, rank),
                                                   tokenizer=tokenizer)
        model.to(local_rank)
        out_dir = os.path.join(args.out_dir, get_dataset_name(data_dir))
        print("out_dir: ", out_dir)
        os.makedirs(out_dir, exist_ok=True)
        model.base_model.config.decoding = True
        # the trie was built on the full codes; shorter DocIDs (max_new_token < L) map to leaf ranges
        feed = PrefetchLoader(dev_loader, depth=2, device=torch.device("cuda", local_rank))
        constrained_decode_doc(model.base_model, feed, prefix_constrain_processor, None, max_new_token,
                               device=local_rank, out_dir=out_dir, local_rank=local_rank, topk=args.topk,
                               apply_log_softmax_for_scores=args.apply_log_softmax_for_scores,
                               write="gather" if getattr(args, "gather", False) and world > 1 else True)


def merge_rank_runs(sub_runs: List[Dict]) -> Dict:
    """evaluate.py:506-515: dict update per qid (duplicates from sampler padding collapse)."""
    merged: Dict = {}
    for sub in sub_runs:
        if len(merged) == 0:
            merged.update(sub)
        else:
            for qid, rankdata in sub.items():
                if qid not in merged:
                    merged[qid] = rankdata
                else:
                    merged[qid].update(rankdata)
    return merged


def t5seq_aq_retrieve_docids_2(args):
    q_paths = args.q_collection_paths
    if len(q_paths) == 1 and q_paths[0].lstrip().startswith("["):
        q_paths = json.loads(q_paths[0])
    for data_dir in q_paths:
        out_dir = os.path.join(args.out_dir, get_dataset_name(data_dir))
        if os.path.exists(os.path.join(out_dir, "run.json")):
            print("old run.json exisit.")
            os.remove(os.path.join(out_dir, "run.json"))
        sub_paths = [p for p in os.listdir(out_dir) if "run" in p]
        expected = args.num_ranks if args.num_ranks else torch.cuda.device_count()
        assert len(sub_paths) == expected, (sub_paths, expected)          # evaluate.py:503-504
        subs = []
        for sub_path in sub_paths:
            with open(os.path.join(out_dir, sub_path)) as fin:
                subs.append(json.load(fin))
        merged = merge_rank_runs(subs)
        print("length of pids and avg rankdata length in qid_to_rankdata: {}, {}".format(
            len(merged), np.mean([len(xs) for xs in merged.values()])))
        with open(os.path.join(out_dir, "run.json"), "w") as fout:
            json.dump(merged, fout)
        for sub_path in sub_paths:
            os.remove(os.path.join(out_dir, sub_path))
    if args.eval_qrel_path:
        evaluate_runs(args, q_paths)


def pack_run(local_run: Dict) -> Dict[str, np.ndarray]:
    """{qid: {docid: score}} -> flat arrays: qids int64 [n], lengths int64 [n], docids as bytes + offsets, scores f32.
    Insertion order is kept (the merged run.json must list a query's documents in rank order)."""
    qids = np.fromiter((int(q) for q in local_run), dtype=np.int64, count=len(local_run))
    lens = np.fromiter((len(v) for v in local_run.values()), dtype=np.int64, count=len(local_run))
    names = [str(d).encode() for v in local_run.values() for d in v]
    scores = np.fromiter((float(x) for v in local_run.values() for x in v.values()), dtype=np.float64, count=len(names))
    off = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in names], out=off[1:])
    return {"qids": qids, "lens": lens, "off": off, "scores": scores,
            "names": np.frombuffer(b"".join(names), dtype=np.uint8).copy()}


def unpack_run(p: Dict[str, np.ndarray]) -> Dict:
    run: Dict = {}
    names, off, scores = p["names"].tobytes(), p["off"], p["scores"]
    e = 0
    for qid, n in zip(p["qids"].tolist(), p["lens"].tolist()):
        d = run.setdefault(qid, {})
        for _ in range(n):
            d[names[off[e]: off[e + 1]].decode()] = float(scores[e])
            e += 1
    return run


def gather_runs(local_run: Dict, group=None) -> Optional[Dict]:
    """Collective replacement of the reference's file merge (evaluate.py:130-132 run_{rank}.json, :489-526 merge):
    every rank contributes its {qid: {docid: score}} as packed tensors over NCCL (gloo on CPU); rank 0 returns the
    merged dict (duplicated queries from DistributedSampler padding collapse, as dict.update does there), the other
    ranks None. One all_gather of the five array lengths and one all_gather_into_tensor of the byte payloads."""
    import torch.distributed as dist
    if not dist.is_initialized():
        return local_run
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    p = pack_run(local_run)
    order = ("qids", "lens", "off", "scores", "names")
    blobs = [np.ascontiguousarray(p[k]).view(np.uint8).reshape(-1) for k in order]
    sizes = torch.tensor([b.size for b in blobs], dtype=torch.int64, device=dev)
    all_sizes = torch.zeros(world * len(order), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.cpu().numpy().reshape(world, len(order))
    cap = int(all_sizes.sum(axis=1).max())
    cap = (cap + 15) // 16 * 16
    payload = torch.zeros(cap, dtype=torch.uint8, device=dev)
    flat = np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)
    payload[: flat.size] = torch.from_numpy(flat).to(dev)
    gathered = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, payload, group=group)
    if rank != 0:
        return None
    gathered = gathered.cpu().numpy().reshape(world, cap)
    dtypes = {"qids": np.int64, "lens": np.int64, "off": np.int64, "scores": np.float64, "names": np.uint8}
    subs = []
    for r in range(world):
        o, parts = 0, {}
        for k, n in zip(order, all_sizes[r].tolist()):
            parts[k] = gathered[r, o: o + n].copy().view(dtypes[k])
            o += n
        subs.append(unpack_run(parts))
    return merge_rank_runs(subs)


def gather_ranked_lists(qids: torch.Tensor, doc_rows: torch.Tensor, scores: torch.Tensor, group=None):
    """The per-batch collective of the data path: all_gather of the packed ranked lists
    ``doc_rows`` int64 [B_loc, m] (+ ``scores`` fp32 [B_loc, n], ``qids`` int64 [B_loc]) over NCCL. Every rank gets
    ([world*B_loc], [world*B_loc, m], [world*B_loc, n]) in global query order; rows of padded (duplicated) queries
    are dropped by the caller with ``unique_queries``. Without a process group the inputs are returned."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return qids, doc_rows, scores
    world = dist.get_world_size(group)
    B, m = doc_rows.shape
    n = scores.shape[1]
    # one packed int64 buffer per rank: [qid | doc rows | score bits]
    packed = torch.empty((B, 1 + m + n), dtype=torch.int64, device=doc_rows.device)
    packed[:, 0] = qids.to(doc_rows.device)
    packed[:, 1: 1 + m] = doc_rows
    packed[:, 1 + m:] = scores.to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    out = torch.empty((world * B, 1 + m + n), dtype=torch.int64, device=doc_rows.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    # DistributedSampler order: global index i*world + rank  <->  row rank*B + i
    out = out.view(world, B, -1).transpose(0, 1).reshape(world * B, -1)
    return out[:, 0], out[:, 1: 1 + m], out[:, 1 + m:].to(torch.int32).view(torch.float32)


def unique_queries(qids: torch.Tensor, n_queries: int) -> torch.Tensor:
    """Row indices that keep the first occurrence of every query (DistributedSampler pads by wrapping around)."""
    return torch.arange(min(n_queries, qids.shape[0]), device=qids.device)


def mrr_k(run: Dict, qrel: Dict, k: int = 10) -> float:
    """MRR@k as reference utils/metrics.py:9-25 computes it, without pytrec_eval (not installed): ``truncate_run``
    keeps the first k documents of a stable sort by score (descending); pytrec_eval's ``recip_rank`` then ranks those
    by score with ties broken by docid in DESCENDING lexicographic order (trec_eval's sort), looks at the queries that
    are in both the run and the qrel, and the mean is taken over those queries only."""
    total, n = 0.0, 0
    for qid, docs in run.items():
        rel = qrel.get(qid)
        if rel is None:
            continue
        top = sorted(docs.items(), key=lambda kv: kv[1], reverse=True)[:k]
        top.sort(key=lambda kv: kv[0], reverse=True)            # docid descending ...
        top.sort(key=lambda kv: kv[1], reverse=True)            # ... within equal scores (stable)
        rr = 0.0
        for i, (docid, _) in enumerate(top):
            if rel.get(docid, 0) > 0:
                rr = 1.0 / (i + 1)
                break
        total += rr
        n += 1
    return total / max(n, 1)


def evaluate_runs(args, q_paths):
    for data_dir in q_paths:
        out_dir = os.path.join(args.out_dir, get_dataset_name(data_dir))
        with open(os.path.join(out_dir, "run.json")) as f:
            run = json.load(f)
        for qrel_path in args.eval_qrel_path:
            if get_dataset_name(qrel_path) != get_dataset_name(data_dir) or not os.path.exists(qrel_path):
                continue
            with open(qrel_path) as f:
                qrel = json.load(f)
            perf = {"mrr_10": mrr_k(run, qrel, 10)}
            print(get_dataset_name(data_dir), perf)
            with open(os.path.join(out_dir, "perf.json"), "w") as f:
                json.dump(perf, f)


def constrained_decode_smtid(model, dataloader, prefix_constrain_processor, smtid_to_docids, max_new_token, device,
                             out_dir, local_rank, topk=100, apply_log_softmax_for_scores=False, write=True):
    """reference evaluate.py:134-178: {qid: {smtid_prefix: {docid: score}}}; prefixes that are not in the trie keep
    an empty dict. ``smtid_to_docids`` may be None: the docids of a prefix are then read from the trie leaves."""
    qid_to_rankdata: Dict[int, Dict[str, Dict[str, float]]] = {}
    trie: DocidTrie = prefix_constrain_processor.trie
    for batch in dataloader:
        outputs = generate_for_constrained_prefix_beam_search(
            model, prefix_constrain_processor, input_ids=batch["input_ids"].long(),
            attention_mask=batch["attention_mask"].long(), max_new_tokens=max_new_token, output_scores=True,
            return_dict=True, return_dict_in_generate=True, num_beams=topk, num_return_sequences=topk,
            apply_log_softmax_for_scores=apply_log_softmax_for_scores)
        batch_qids = batch["id"].cpu().tolist()
        str_smtids = convert_ptsmtids_to_strsmtid(outputs.sequences.view(-1, topk, max_new_token + 1), max_new_token)
        relevant_scores = outputs.sequences_scores.view(-1, topk).cpu().tolist()
        docs, counts, leaf = _expand_on_device(trie, outputs, topk)
        for q, (qid, ranked_smtids, rel_scores) in enumerate(zip(batch_qids, str_smtids, relevant_scores)):
            qid_to_rankdata[qid] = {}
            for j, (smtid, rel_score) in enumerate(zip(ranked_smtids, rel_scores)):
                qid_to_rankdata[qid][smtid] = {}
                if smtid_to_docids is not None:
                    docids = smtid_to_docids.get(smtid, [])
                else:
                    docids = _docids_of(trie, docs, counts, leaf, q, j)
                for docid in docids:
                    qid_to_rankdata[qid][smtid][docid] = rel_score if apply_log_softmax_for_scores \
                        else rel_score * max_new_token
    if write:
        with open(os.path.join(out_dir, f"qid_smtid_rankdata_{local_rank}.json"), "w") as fout:
            json.dump(qid_to_rankdata, fout)
    return qid_to_rankdata


def t5seq_aq_get_qid_to_smtid_rankdata(args, tokenizer=None):
    """reference evaluate.py:528-611: beam search to DocID *prefixes* of length max_new_token in {4,8,16,32} over
    the full-length trie, for the train queries."""
    ddp_setup()
    model = T5SeqAQEncoder.from_pretrained(args.pretrained_path)
    model.eval()
    V = model.config.decoder_vocab_sizes
    if len(set(V)) == 2:
        raise NotImplementedError
    if len(set(V)) != 1:
        raise ValueError("not valid decoder_vocab_size")
    trie = load_docid_trie(args.docid_to_smtid_path, V[0], args.local_rank)     # (asserts smtids[0] == -1)
    prefix_constrain_processor = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
    assert args.max_new_token in [4, 8, 16, 32], args.max_new_token
    if args.local_rank <= 0:
        os.makedirs(args.out_dir, exist_ok=True)
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    local_rank = max(args.local_rank, 0)
    dev_dataset = CollectionDatasetWithDocIDPreLoad(data_dir=args.train_query_dir, id_style="row_id",
                                                    tid_to_smtid_path=None, add_prefix=True, is_query=True)
    dev_loader = CollectionDataWithDocIDLoader(dataset=dev_dataset, tokenizer_type=args.pretrained_path, max_length=256,
                                               batch_size=args.batch_size,
                                               sampler=distributed_sampler_indices(len(dev_dataset), world, rank),
                                               tokenizer=tokenizer)
    model.to(local_rank)
    print("out_dir: ", args.out_dir)
    model.base_model.config.decoding = True
    feed = PrefetchLoader(dev_loader, depth=2, device=torch.device("cuda", local_rank))
    return constrained_decode_smtid(model.base_model, feed, prefix_constrain_processor, None, args.max_new_token,
                                    device=local_rank, out_dir=args.out_dir, local_rank=local_rank, topk=args.topk,
                                    apply_log_softmax_for_scores=args.apply_log_softmax_for_scores)


def merge_rank_smtid_rankdata(subs: List[Dict]) -> Dict:
    """reference evaluate.py:625-637."""
    merged: Dict = {}
    for sub in subs:
        for qid in sub:
            if qid not in merged:
                merged[qid] = sub[qid]
            else:
                for smtid in sub[qid]:
                    if smtid not in merged[qid]:
                        merged[qid][smtid] = sub[qid][smtid]
                    else:
                        for docid, score in sub[qid][smtid].items():
                            merged[qid][smtid][docid] = score
    return merged


def t5seq_aq_get_qid_to_smtid_rankdata_2(args):
    out_dir = args.out_dir
    if os.path.exists(os.path.join(out_dir, "qid_smtid_rankdata.json")):
        print("old run.json exisit.")
        os.remove(os.path.join(out_dir, "qid_smtid_rankdata.json"))
    sub_paths = [p for p in os.listdir(out_dir) if "qid_smtid_rankdata" in p]
    expected = args.num_ranks if args.num_ranks else torch.cuda.device_count()
    assert len(sub_paths) == expected, (sub_paths, expected)
    subs = []
    for sub_path in sub_paths:
        with open(os.path.join(out_dir, sub_path)) as fin:
            subs.append(json.load(fin))
    merged = merge_rank_smtid_rankdata(subs)
    smtid_lengths = [len(v) for v in merged.values()]
    doc_lengths = [len(d) for v in merged.values() for d in v.values()]
    q = [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]
    print("smtid_length per query: ", np.quantile(smtid_lengths, q))
    print("doc_length per smtid: ", np.quantile(doc_lengths, q))
    with open(os.path.join(out_dir, "qid_smtid_rankdata.json"), "w") as fout:
        json.dump(merged, fout)
    for sub_path in sub_paths:
        os.remove(os.path.join(out_dir, sub_path))


def get_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", type=str, default="")
    ap.add_argument("--pretrained_path", type=str, default="")
    ap.add_argument("--out_dir", type=str, default="")
    ap.add_argument("--docid_to_smtid_path", type=str, default=None)
    ap.add_argument("--q_collection_paths", type=str, nargs="+", default=[])
    ap.add_argument("--eval_qrel_path", type=str, nargs="*", default=[])
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--topk", type=int, default=200)
    ap.add_argument("--max_new_token_for_docid", type=int, default=32)
    ap.add_argument("--max_new_token", type=int, default=None)
    ap.add_argument("--train_query_dir", type=str, default=None)
    ap.add_argument("--apply_log_softmax_for_scores", action="store_true")
    ap.add_argument("--local_rank", type=int, default=int(os.environ.get("LOCAL_RANK", -1)))
    ap.add_argument("--gather", action="store_true",
                    help="merge the per-rank runs with one NCCL gather and let rank 0 write run.json directly "
                         "(instead of run_{rank}.json files + the _2 merge task)")
    ap.add_argument("--num_ranks", type=int, default=0, help="run_*.json files expected by the merge task "
                                                              "(default: torch.cuda.device_count() like the reference)")
    return ap.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    if args.task == "t5seq_aq_retrieve_docids":
        t5seq_aq_retrieve_docids(args)
    elif args.task == "t5seq_aq_retrieve_docids_2":
        t5seq_aq_retrieve_docids_2(args)
    elif args.task == "t5seq_aq_get_qid_to_smtid_rankdata":
        t5seq_aq_get_qid_to_smtid_rankdata(args)
    elif args.task == "t5seq_aq_get_qid_to_smtid_rankdata_2":
        t5seq_aq_get_qid_to_smtid_rankdata_2(args)
    else:
        raise ValueError(f"task {args.task!r} is not part of the retrieval path served here "
                         "(t5seq_aq_retrieve_docids[_2], t5seq_aq_get_qid_to_smtid_rankdata[_2])")


if __name__ == "__main__":
    main()
