"""The reference's generation API for the retrieval path, served by the CUDA engine.

Same names, arguments and results as reference ``t5_pretrainer/tasks/generation.py``:
``generate_for_constrained_prefix_beam_search`` (:35-251) and ``PrefixConstrainLogitProcessorFastSparse``
(:603-677). The whole loop (encoder, L decoder steps, trie mask, float64 beam arithmetic, top-k,
finalize) runs inside ``rb200_engine_search*``; nothing is computed in PyTorch.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional

import torch

from . import _lib
from .trie import DocidTrie


class BeamSearchEncoderDecoderOutput:
    """The fields of HF's BeamSearchEncoderDecoderOutput that reference callers read (evaluate.py:116-117)."""

    def __init__(self, sequences, sequences_scores, leaf_ranges=None):
        self.sequences = sequences
        self.sequences_scores = sequences_scores
        self.leaf_ranges = leaf_ranges      # extra: [B*nrs, 2] trie leaf range per row (device-side docid lookup)
        self.scores = None                  # the reference fills these when output_scores=True but no caller reads them
        self.beam_indices = None

    def __getitem__(self, k):
        return getattr(self, k)


class PrefixConstrainLogitProcessorFastSparse:
    """Allowed-next-token mask from the DocID trie; ``__call__(input_ids, scores) -> valid_mask[R, V]`` float64,
    1.0 allowed / 0.0 not, all-zero row for a prefix that is not in the trie (reference generation.py:666-677)."""

    def __init__(self, list_smtid_to_nextids: Optional[List[Dict[str, Iterable[int]]]], vocab_size: int,
                 trie: Optional[DocidTrie] = None):
        self.vocab_size = vocab_size
        if trie is None:
            if not list_smtid_to_nextids:
                raise ValueError("list_smtid_to_nextids is empty")
            trie = DocidTrie.from_list_smtid_to_nextids(list_smtid_to_nextids, vocab_size)
        if trie.V != vocab_size:
            raise ValueError(f"trie was built for vocab_size {trie.V}, processor asked for {vocab_size}")
        self.trie = trie

    @classmethod
    def from_trie(cls, trie: DocidTrie) -> "PrefixConstrainLogitProcessorFastSparse":
        return cls(None, trie.V, trie=trie)

    def __call__(self, input_ids: torch.Tensor, next_token_scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert input_ids.dim() == 2
        return self.trie.mask(input_ids)


def generate_for_constrained_prefix_beam_search(model, valid_smtids, inputs: Optional[torch.Tensor] = None,
                                                max_length: Optional[int] = None,
                                                num_beams: Optional[int] = None,
                                                num_return_sequences: Optional[int] = None,
                                                max_new_tokens: Optional[int] = None,
                                                output_scores: Optional[bool] = None,
                                                return_dict_in_generate: Optional[bool] = None,
                                                apply_log_softmax_for_scores: Optional[bool] = False,
                                                input_ids: Optional[torch.Tensor] = None,
                                                attention_mask: Optional[torch.Tensor] = None,
                                                precision: Optional[str] = None, **model_kwargs):
    """Drop-in for reference generation.py:35. ``valid_smtids`` is the PrefixConstrainLogitProcessorFastSparse.

    input_ids / attention_mask: LongTensor [B, S] on the host (pinned or pageable: copied inside the call)
    or on the model's CUDA device. Returns an object with ``.sequences`` LongTensor [B*nrs, L+1] (column 0
    is the decoder start id 0, rows grouped per query in descending score) and ``.sequences_scores``
    FloatTensor [B*nrs], on the device of ``input_ids``.
    """
    input_ids = input_ids if input_ids is not None else inputs
    if input_ids is None or attention_mask is None:
        raise ValueError("input_ids and attention_mask are required")
    num_beams = num_beams if num_beams is not None else 1
    num_return_sequences = num_return_sequences if num_return_sequences is not None else 1
    if max_new_tokens is None:
        if max_length is None:
            raise ValueError("`max_length` needs to be a stopping_criteria for now.")
        max_new_tokens = max_length - 1                              # generation.py:153-154
    if num_return_sequences > num_beams:
        raise ValueError("`num_return_sequences` has to be smaller or equal to `num_beams`.")   # :216-217
    if num_beams < 2:    # HF 4.17 BeamSearchScorer.__init__ (constructed at generation.py:222) refuses num_beams <= 1
        raise ValueError(f"`num_beams` has to be an integer strictly greater than 1, but is {num_beams}.")
    base = getattr(model, "base_model", model)
    trie: DocidTrie = valid_smtids.trie
    B, S = input_ids.shape
    ids = input_ids.to(torch.int64).contiguous()
    mask = attention_mask.to(torch.int64).contiguous()
    auto = (precision or base.precision) == "auto"
    n = B * num_return_sequences
    L = _lib.lib()
    while True:
        mode = base.resolve_precision(precision)
        try:
            engine = base.get_engine(B, num_beams, S, mode)
        except ValueError as err:                                    # a WEIGHT does not fit the fp16 planes
            if auto and mode == "fp16x3" and "fp16 range" in str(err):
                base.fp16_ok = False
                continue
            raise
        dev_index = engine.key[0]
        trie.upload(dev_index)
        with torch.cuda.device(dev_index):
            stream = _lib.stream_ptr()
            if ids.is_cuda:
                seqs = torch.empty((n, max_new_tokens + 1), dtype=torch.int64, device=ids.device)
                scores = torch.empty((n,), dtype=torch.float32, device=ids.device)
                leaf = torch.empty((n, 2), dtype=torch.int32, device=ids.device)
                _lib.check(L.rb200_engine_search(engine.h, trie.handle, ids.data_ptr(), mask.data_ptr(), B, S,
                                                 num_beams, max_new_tokens, num_return_sequences,
                                                 int(bool(apply_log_softmax_for_scores)), seqs.data_ptr(),
                                                 scores.data_ptr(), leaf.data_ptr(), stream))
            else:
                # pinned staging buffers live with the engine (cudaHostAlloc costs milliseconds); the caller gets
                # its own pageable copies, as the reference's `.cpu()` results are
                key = (n, max_new_tokens + 1)
                if getattr(engine, "host_out_key", None) != key:
                    engine.host_out = (torch.empty((n, max_new_tokens + 1), dtype=torch.int64).pin_memory(),
                                       torch.empty((n,), dtype=torch.float32).pin_memory(),
                                       torch.empty((n, 2), dtype=torch.int32).pin_memory())
                    engine.host_out_key = key
                pseqs, pscores, pleaf = engine.host_out
                _lib.check(L.rb200_engine_search_host(engine.h, trie.handle, ids.data_ptr(), mask.data_ptr(), B, S,
                                                      num_beams, max_new_tokens, num_return_sequences,
                                                      int(bool(apply_log_softmax_for_scores)), pseqs.data_ptr(),
                                                      pscores.data_ptr(), pleaf.data_ptr(), stream))
                seqs, scores, leaf = pseqs.clone(), pscores.clone(), pleaf.clone()
        # fp16x3 poisons EVERY score with NaN when an activation left the fp16 range: in auto mode look at one
        # element (4-byte read) and redo the batch in tf32x3; explicit fp16x3 hands the NaNs to the caller
        if auto and mode == "fp16x3" and bool(torch.isnan(scores[:1]).item()):
            base.fp16_ok = False
            base.drop_engine("fp16x3")           # do not keep two sets of packed weights and workspaces in HBM
            continue
        break
    out = BeamSearchEncoderDecoderOutput(seqs, scores, leaf)
    out.gpu_launches = int(L.rb200_engine_last_launch_count(engine.h))
    out.precision = mode
    out.forced_tail_from = int(L.rb200_engine_last_tail_step(engine.h))   # -1: every query stepped to the end
    hist = (C.c_int32 * 33)()
    rows = C.c_int64()
    _lib.check(L.rb200_engine_last_freeze_histogram(engine.h, hist, 33, C.byref(rows)))
    out.frozen_at_step = list(hist)[: max_new_tokens + 1]                # queries frozen after t steps
    out.forced_tail_rows = int(rows.value)
    if return_dict_in_generate is False:
        return seqs
    return out
