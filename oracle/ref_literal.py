"""ORACLE (test infrastructure, never on the product path): run the reference's OWN beam loop and mask
class, loaded by path from /root/reference, so the restatement in oracle/beam.py can be pinned to it.

Only usable where /root/reference is mounted (the build container). Nothing from the reference is
copied: ``t5_pretrainer/tasks/generation.py`` is executed in place through importlib after registering
stand-ins for the HF transformers==4.17.0 modules it imports and that no longer exist in the installed
transformers:

  transformers.generation_beam_search      BeamScorer, BeamSearchScorer -> oracle.beam.BeamSearchScorerOracle
  transformers.generation_utils            8 output classes -> attribute bags
  transformers.generation_logits_process   LogitsProcessorList -> identity list
  transformers.generation_stopping_criteria StoppingCriteriaList (4.17 semantics: python bool, .max_length)
  transformers.generation_beam_constraints Constraint
  transformers.pytorch_utils.torch_int_div floor division
  ujson                                    -> json

What is literal: ``beam_search_for_constrained_prefix`` (generation.py:253-575) and
``PrefixConstrainLogitProcessorFastSparse`` (generation.py:603-677). What is supplied: the scorer
(restated HF 4.17) and a model adapter whose arithmetic is oracle/t5_math.py, driven the way the
reference drives its model (full prefix every step, KV cache produced but never consumed).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
from typing import Optional

import torch

REFERENCE_ROOT = os.environ.get("RIPOR_REFERENCE_ROOT", "/root/reference")
_GEN_PATH = os.path.join(REFERENCE_ROOT, "t5_pretrainer", "tasks", "generation.py")
_module = None


def available() -> bool:
    return os.path.isfile(_GEN_PATH)


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _LogitsProcessorList(list):
    def __call__(self, input_ids, scores, **kw):
        for p in self:
            scores = p(input_ids, scores)
        return scores


class _StoppingCriteriaList(list):
    """HF 4.17: ``__call__`` returns a python bool; ``max_length`` property scans MaxLengthCriteria."""

    def __init__(self, max_length: Optional[int] = None):
        super().__init__([("max_length", max_length)] if max_length is not None else [])
        self._max_length = max_length

    @property
    def max_length(self):
        return self._max_length

    def __call__(self, input_ids, scores, **kw) -> bool:
        return self._max_length is not None and input_ids.shape[-1] >= self._max_length


def _validate_stopping_criteria(sc, max_length):
    return sc


def load_reference_generation():
    """Import /root/reference/t5_pretrainer/tasks/generation.py in place with the HF-4.17 stand-ins."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise FileNotFoundError(f"{_GEN_PATH} not present: the literal reference only exists in the build container")
    import transformers  # noqa: F401  (installed 5.x; provides modeling_outputs / utils.logging used by the file)
    import transformers.pytorch_utils as pu
    from oracle.beam import BeamSearchScorerOracle

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class BeamScorer:  # abstract base in HF; only used as a type annotation
        pass

    mod("transformers.generation_beam_search", BeamScorer=BeamScorer, BeamSearchScorer=BeamSearchScorerOracle)
    names = ["BeamSearchOutput", "BeamSearchEncoderDecoderOutput", "BeamSearchDecoderOnlyOutput",
             "GreedySearchOutput", "GreedySearchDecoderOnlyOutput", "GreedySearchEncoderDecoderOutput",
             "SampleOutput", "BeamSampleOutput"]
    mod("transformers.generation_utils", **{n: type(n, (_Bag,), {}) for n in names})
    mod("transformers.generation_logits_process", LogitsProcessorList=_LogitsProcessorList)
    mod("transformers.generation_stopping_criteria", StoppingCriteriaList=_StoppingCriteriaList,
        validate_stopping_criteria=_validate_stopping_criteria)
    mod("transformers.generation_beam_constraints", Constraint=type("Constraint", (), {}))
    if "ujson" not in sys.modules:
        mod("ujson", load=json.load, loads=json.loads, dump=json.dump, dumps=json.dumps)
    if not hasattr(pu, "torch_int_div"):
        pu.torch_int_div = lambda a, b: torch.div(a, b, rounding_mode="floor")
    spec = importlib.util.spec_from_file_location("_ripor_reference_generation", _GEN_PATH)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    _module = m
    return m


class FullPrefixModelAdapter:
    """Stands where the reference's T5ForDocIDGeneration stands inside its beam loop.

    ``prepare_inputs_for_generation`` mirrors t5_generative_retriever.py:452-479 (the cache arrives under
    the key ``past`` and is therefore ignored: the full prefix is decoded every step); ``__call__``
    mirrors the decode branch of ``forward`` (:396-450) with config.decoding=True.
    """

    def __init__(self, weights, dims):
        from oracle import t5_math
        self._m, self.w, self.dims = t5_math, weights, dims
        self.config = _Bag(output_scores=False, output_attentions=False, output_hidden_states=False,
                           return_dict_in_generate=False, is_encoder_decoder=True)
        self.forward_calls = 0

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None,
                                      encoder_outputs=None, **kwargs):
        if past_key_values is not None:
            input_ids = input_ids[:, -1:]
        return {"decoder_input_ids": input_ids, "past_key_values": past_key_values,
                "encoder_outputs": encoder_outputs, "attention_mask": attention_mask}

    def __call__(self, decoder_input_ids=None, past_key_values=None, encoder_outputs=None, attention_mask=None,
                 return_dict=True, output_attentions=None, output_hidden_states=None, **kw):
        self.forward_calls += 1
        hidden = self._m.decoder_full_prefix(self.w, self.dims, decoder_input_ids, encoder_outputs[0], attention_mask)
        logits = self._m.lm_logits_list(self.w, self.dims, hidden)
        return _Bag(logits=logits, past_key_values=None, decoder_last_hidden_state=hidden)

    def adjust_logits_during_generation(self, logits, **kw):
        return logits

    def _update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False):
        model_kwargs["past"] = outputs.past_key_values          # HF 4.17 key; see SURVEY.md §0 finding 4
        return model_kwargs

    def _reorder_cache(self, past, beam_idx):
        return past


class LogitsTableModelAdapter(FullPrefixModelAdapter):
    """Model whose last-position logits are a pure function of the prefix, for scorer/mask-only checks."""

    def __init__(self, fn):
        self.fn = fn
        self.config = _Bag(output_scores=False, output_attentions=False, output_hidden_states=False,
                           return_dict_in_generate=False, is_encoder_decoder=True)

    def __call__(self, decoder_input_ids=None, **kw):
        return _Bag(logits=[self.fn(decoder_input_ids)], past_key_values=None)


def literal_processor(list_smtid_to_nextids, vocab_size):
    return load_reference_generation().PrefixConstrainLogitProcessorFastSparse(list_smtid_to_nextids, vocab_size)


def literal_beam_search(model, processor, batch_size, num_beams, max_new_tokens, num_return_sequences=None,
                        apply_log_softmax_for_scores=False, encoder_states=None, attention_mask=None):
    """Drive the literal ``beam_search_for_constrained_prefix`` the way generation.py:222-251 does."""
    gen = load_reference_generation()
    keep = num_return_sequences if num_return_sequences is not None else num_beams
    from oracle.beam import BeamSearchScorerOracle
    scorer = BeamSearchScorerOracle(batch_size, num_beams, num_beam_hyps_to_keep=keep)
    input_ids = torch.zeros((batch_size, 1), dtype=torch.long)
    idx = torch.arange(batch_size).view(-1, 1).repeat(1, num_beams).view(-1)       # _expand_inputs_for_generation
    input_ids = input_ids.index_select(0, idx)
    kwargs = {}
    if encoder_states is not None:
        kwargs["encoder_outputs"] = (encoder_states.index_select(0, idx),)
        kwargs["attention_mask"] = attention_mask.index_select(0, idx)
    with torch.no_grad():
        out = gen.beam_search_for_constrained_prefix(
            model, processor, input_ids, scorer,
            logits_processor=_LogitsProcessorList(),
            stopping_criteria=_StoppingCriteriaList(max_length=max_new_tokens + 1),
            pad_token_id=0, eos_token_id=1, output_scores=True, return_dict_in_generate=True,
            synced_gpus=False, apply_log_softmax_for_scores=apply_log_softmax_for_scores, **kwargs)
    return out


_legacy = None


def load_reference_legacy_trie():
    """Import /root/reference/t5_pretrainer/utils/generation_utils.py in place: the nested-dict ``Trie`` (:9-90) and
    ``PrefixConstrainedLogitsProcessorForSmtidTree`` (:92-124, -inf mask, eos fallback) that predate the sparse
    processor. Used as the third voice of the allowed-token cross-check (SURVEY 8c golden item 2)."""
    global _legacy
    if _legacy is not None:
        return _legacy
    path = os.path.join(os.path.dirname(os.path.dirname(_GEN_PATH)), "utils", "generation_utils.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    lp = sys.modules.get("transformers.generation_logits_process")
    if lp is None:
        lp = types.ModuleType("transformers.generation_logits_process")
        sys.modules["transformers.generation_logits_process"] = lp
    if not hasattr(lp, "LogitsProcessor"):
        lp.LogitsProcessor = type("LogitsProcessor", (), {})
    if not hasattr(lp, "LogitsProcessorList"):
        lp.LogitsProcessorList = _LogitsProcessorList
    if "ujson" not in sys.modules:
        m = types.ModuleType("ujson")
        m.__dict__.update(load=json.load, loads=json.loads, dump=json.dump, dumps=json.dumps)
        sys.modules["ujson"] = m
    spec = importlib.util.spec_from_file_location("_ripor_reference_generation_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _legacy = mod
    return mod

