"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's
prefix-constrained beam search and of the host-side smtid -> docid mapping.

Follows
  * ``t5_pretrainer/tasks/generation.py:603-677``  PrefixConstrainLogitProcessorFastSparse (the trie mask)
  * ``t5_pretrainer/tasks/generation.py:381-382,418-420,423-540``  beam_search_for_constrained_prefix
  * HF transformers==4.17.0 ``BeamSearchScorer.process/finalize`` + ``BeamHypotheses.add`` (third-party,
    not vendored in /root/reference; call sites generation.py:222-229,496-503,532-540) restated per
    SURVEY.md Appendix A.3
  * ``t5_pretrainer/evaluate.py:411-424`` (per-level prefix -> next ids dicts), ``:439-446`` (smtid -> docids),
    ``:116-128`` (run dict, score*L), ``t5_pretrainer/utils/utils.py:46-59`` (ids -> "c1_.._cL").

Pinning: the reference holds no tests or golden vectors for this path, so this restatement
is pinned against the reference's own code run in the build container: oracle/ref_literal.py imports the
literal ``beam_search_for_constrained_prefix`` and ``PrefixConstrainLogitProcessorFastSparse`` from
/root/reference and tests/test_oracle_literal.py + the fixtures made by oracle/make_golden.py compare the
two on seeded inputs.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------
# trie inputs (evaluate.py:411-424) and the mask (generation.py:603-677)
# ----------------------------------------------------------------------------------------------
def build_list_smtid_to_nextids(docid_to_smtids: Dict[str, Sequence[int]]) -> List[Dict[str, List[int]]]:
    """evaluate.py:411-424 / aq_preprocess/build_list_smtid_to_nextids.py:23-36."""
    L = len(next(iter(docid_to_smtids.values()))) - 1
    out: List[Dict[str, set]] = [dict() for _ in range(L)]
    for _, smtids in docid_to_smtids.items():
        for i in range(L):
            key = "_".join(str(x) for x in smtids[: i + 1])
            out[i].setdefault(key, set()).add(int(smtids[i + 1]))
    return [{k: list(v) for k, v in d.items()} for d in out]


class TrieMaskOracle:
    """Same observable behaviour as PrefixConstrainLogitProcessorFastSparse.__call__ (generation.py:666-677):
    float64 [R,V] mask, 1.0 on allowed next tokens, all-zero row for an unknown prefix."""

    def __init__(self, list_smtid_to_nextids: List[Dict[str, List[int]]], vocab_size: int):
        self.levels = [{k: np.asarray(sorted(v), dtype=np.int64) for k, v in d.items()}
                       for d in list_smtid_to_nextids]
        self.vocab_size = vocab_size

    def __call__(self, input_ids: torch.Tensor, scores: Optional[torch.Tensor] = None) -> torch.Tensor:
        ids = input_ids.cpu().numpy()
        R, T = ids.shape
        mask = np.zeros((R, self.vocab_size), dtype=np.float64)
        level = self.levels[T - 1]
        for r in range(R):
            key = "-1" if T == 1 else "-1_" + "_".join(str(int(x)) for x in ids[r, 1:])
            nxt = level.get(key)
            if nxt is not None:
                mask[r, nxt] = 1.0
        return torch.from_numpy(mask)


# ----------------------------------------------------------------------------------------------
# HF 4.17 BeamSearchScorer, restated for the only regime the reference reaches (eos/pad = None)
# ----------------------------------------------------------------------------------------------
class BeamSearchScorerOracle:
    """process(): walk the 2*nb sorted candidates and keep the first nb (no candidate is ever an eos).
    finalize(): every beam is added with score / len**length_penalty (python float64), hypotheses are
    sorted ascending (stable) and popped from the end, scores stored as float32."""

    def __init__(self, batch_size: int, num_beams: int, device=None, length_penalty: float = 1.0,
                 do_early_stopping: bool = False, num_beam_hyps_to_keep: int = 1):
        self.num_beams = num_beams
        self.group_size = num_beams
        self.length_penalty = length_penalty
        self.num_beam_hyps_to_keep = num_beam_hyps_to_keep
        self._beam_hyps: List[List[Tuple[float, torch.Tensor]]] = [[] for _ in range(batch_size)]
        self._done = [False] * batch_size

    @property
    def is_done(self) -> bool:
        return all(self._done)

    def process(self, input_ids, next_scores, next_tokens, next_indices, pad_token_id=None, eos_token_id=None):
        assert eos_token_id is None and pad_token_id is None  # generation.py:381-382
        B, nb = len(self._beam_hyps), self.group_size
        nbs = torch.zeros((B, nb), dtype=next_scores.dtype)
        nbt = torch.zeros((B, nb), dtype=next_tokens.dtype)
        nbi = torch.zeros((B, nb), dtype=next_indices.dtype)
        for b in range(B):
            beam_idx = 0
            for tok, sc, idx in zip(next_tokens[b], next_scores[b], next_indices[b]):
                nbs[b, beam_idx] = sc
                nbt[b, beam_idx] = tok
                nbi[b, beam_idx] = b * nb + idx
                beam_idx += 1
                if beam_idx == nb:
                    break
            if beam_idx < nb:
                raise ValueError("fewer than num_beams candidates")
            # BeamHypotheses.is_done: len(beams) < num_beams -> False (no hypothesis is added before finalize)
        return {"next_beam_scores": nbs.view(-1), "next_beam_tokens": nbt.view(-1), "next_beam_indices": nbi.view(-1)}

    def finalize(self, input_ids, final_beam_scores, final_beam_tokens, final_beam_indices, max_length,
                 pad_token_id=None, eos_token_id=None):
        B, nb, keep = len(self._beam_hyps), self.num_beams, self.num_beam_hyps_to_keep
        for b in range(B):
            for j in range(nb):
                r = b * nb + j
                hyp = input_ids[r]
                score = final_beam_scores[r].item() / (hyp.shape[-1] ** self.length_penalty)
                self._beam_hyps[b].append((score, hyp))          # len < num_beams always holds here
        best, best_scores = [], torch.zeros(B * keep, dtype=torch.float32)
        for b in range(B):
            hyps = sorted(self._beam_hyps[b], key=lambda x: x[0])
            for j in range(keep):
                sc, hyp = hyps.pop()
                best.append(hyp)
                best_scores[b * keep + j] = sc
        T = input_ids.shape[-1]
        assert T == max_length
        decoded = torch.stack(best).to(input_ids.dtype)
        return {"sequences": decoded, "sequence_scores": best_scores}


# ----------------------------------------------------------------------------------------------
# the beam loop (generation.py:418-540)
# ----------------------------------------------------------------------------------------------
def beam_search_oracle(step_logits: Callable[[torch.Tensor, Optional[torch.Tensor]], torch.Tensor],
                       mask_fn: Callable[[torch.Tensor, Optional[torch.Tensor]], torch.Tensor],
                       batch_size: int, num_beams: int, max_new_tokens: int,
                       num_return_sequences: Optional[int] = None,
                       apply_log_softmax_for_scores: bool = False,
                       trace: Optional[list] = None):
    """``step_logits(input_ids[R,t+1], beam_idx or None) -> fp32 logits[R,V]`` for the last position.

    ``beam_idx`` is the reorder applied to the rows since the previous call (None at t=0) so a KV-cached
    model can follow the beams; a full-prefix model ignores it. Returns (sequences[B*keep, L+1] int64,
    sequences_scores[B*keep] float32).
    """
    B, nb, L = batch_size, num_beams, max_new_tokens
    keep = num_return_sequences if num_return_sequences is not None else nb
    if keep > nb:
        raise ValueError("`num_return_sequences` has to be smaller or equal to `num_beams`.")  # generation.py:216
    scorer = BeamSearchScorerOracle(B, nb, num_beam_hyps_to_keep=keep)
    input_ids = torch.zeros((B * nb, 1), dtype=torch.long)             # decoder_start_token_id = 0
    beam_scores = torch.zeros((B, nb), dtype=torch.float32)            # :418-420
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    beam_idx = None
    next_tokens = next_indices = None
    for _ in range(L):
        logits = step_logits(input_ids, beam_idx)                      # :448 outputs.logits[-1]
        scores = torch.log_softmax(logits, dim=-1) if apply_log_softmax_for_scores else logits  # :453-458
        valid = mask_fn(input_ids, scores)                             # :461  (float64)
        processed = scores + (1.0 - valid) * (-1e9)                    # :462  -> float64
        nxt = processed + beam_scores[:, None].expand_as(scores)       # :463
        V = nxt.shape[-1]
        nxt = nxt.view(B, nb * V)
        nxt, next_tokens = torch.topk(nxt, 2 * nb, dim=1, largest=True, sorted=True)   # :487-489
        next_indices = torch.div(next_tokens, V, rounding_mode="floor")
        next_tokens = next_tokens % V
        out = scorer.process(input_ids, nxt, next_tokens, next_indices)
        beam_scores = out["next_beam_scores"]
        beam_idx = out["next_beam_indices"]
        input_ids = torch.cat([input_ids[beam_idx, :], out["next_beam_tokens"].unsqueeze(-1)], dim=-1)  # :511
        if trace is not None:
            # cut_gap: distance between the last candidate that stays in the beam and the first one that does not -
            # a gap below the arithmetic noise of an fp32 forward pass is a near-tie, not a parity question
            trace.append({"beam_scores": beam_scores.clone(), "beam_idx": beam_idx.clone(),
                          "tokens": out["next_beam_tokens"].clone(), "processed": processed.clone(),
                          "cut_gap": (nxt[:, nb - 1] - nxt[:, nb]).clone()})
    fin = scorer.finalize(input_ids, beam_scores, next_tokens, next_indices, max_length=L + 1)
    return fin["sequences"], fin["sequence_scores"]


# ----------------------------------------------------------------------------------------------
# host mapping (utils.py:46-59, evaluate.py:439-446, :116-128)
# ----------------------------------------------------------------------------------------------
def convert_ptsmtids_to_strsmtid(input_smtids: torch.Tensor, seq_length: int) -> List[List[str]]:
    assert input_smtids.dim() == 3 and input_smtids.size(2) == seq_length + 1
    return [["_".join(str(x) for x in smt[1:]) for smt in beams] for beams in input_smtids.cpu().tolist()]


def build_smtid_to_docids(docid_to_smtids: Dict[str, Sequence[int]], max_new_token_for_docid: int
                          ) -> Dict[str, List[str]]:
    out: Dict[str, List[str]] = {}
    for docid, smtids in docid_to_smtids.items():
        assert smtids[0] == -1, smtids
        sid = "_".join(str(x) for x in smtids[1: 1 + max_new_token_for_docid])
        out.setdefault(sid, []).append(docid)
    return out


def rankdata_for_batch(qids: Sequence[int], sequences: torch.Tensor, sequences_scores: torch.Tensor,
                       smtid_to_docids: Dict[str, List[str]], topk: int, max_new_token: int,
                       apply_log_softmax_for_scores: bool = False) -> Dict[int, Dict[str, float]]:
    """evaluate.py:115-128 (constrained_decode_doc): {qid: {docid: score or score*L}}; unknown smtids skipped."""
    str_smtids = convert_ptsmtids_to_strsmtid(sequences.view(-1, topk, max_new_token + 1), max_new_token)
    rel = sequences_scores.view(-1, topk).cpu().tolist()
    run: Dict[int, Dict[str, float]] = {}
    for qid, ranked, scs in zip(qids, str_smtids, rel):
        run[qid] = {}
        for smtid, sc in zip(ranked, scs):
            if smtid in smtid_to_docids:
                for docid in smtid_to_docids[smtid]:
                    run[qid][docid] = sc if apply_log_softmax_for_scores else sc * max_new_token
    return run
