"""Host-side mirror of the reference's model API for the retrieval path.

Same names and call shapes as ``t5_pretrainer/modeling/t5_generative_retriever.py`` (reference):
``T5forDocIDConfig`` (:45-67), ``T5ForDocIDGeneration`` (:70-512) and the inference wrapper
``T5SeqAQEncoder`` (:772-855) with ``from_pretrained(path).base_model`` and ``.config.decoder_vocab_sizes``.
PyTorch only holds the fp32 weights (HF state-dict keys); every FLOP of the path
runs in libriporb200.so. There is no eager/PyTorch forward here on purpose: without the CUDA library the
model cannot run.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Optional

import torch

from . import _lib
from .synthetic import T5Dims


class T5GenRetModelOutput:
    """reference t5_generative_retriever.py:40-43 (a Seq2SeqLMOutput with ``logits`` being a list per position)."""

    def __init__(self, logits=None, past_key_values=None, decoder_last_hidden_state=None,
                 encoder_last_hidden_state=None):
        self.loss = None
        self.logits = logits
        self.past_key_values = past_key_values
        self.decoder_last_hidden_state = decoder_last_hidden_state
        self.decoder_hidden_states = None
        self.decoder_attentions = None
        self.cross_attentions = None
        self.encoder_last_hidden_state = encoder_last_hidden_state
        self.encoder_hidden_states = None
        self.encoder_attentions = None

    def __getitem__(self, k):
        return getattr(self, k)


def _device_view(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    """uint8 tensor aliasing a raw device pointer handed out by the C ABI (valid until the next engine call)."""
    class _Holder:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Holder(), device=device)


def _round_up(x: int, m: int) -> int:
    return -(-x // m) * m


class T5forDocIDConfig:
    """Fields of the reference config that the retrieval path reads (config.json of save_pretrained)."""

    def __init__(self, decoder_vocab_sizes=None, decoding=False, decoder_start_token_path=None,
                 scaleup_output_hidden=False, shared_output_input_embeds=True, d_model=768, d_kv=64, d_ff=3072,
                 num_layers=12, num_decoder_layers=None, num_heads=12, vocab_size=32128,
                 relative_attention_num_buckets=32, relative_attention_max_distance=128,
                 layer_norm_epsilon=1e-6, **kwargs):
        self.decoder_vocab_sizes = list(decoder_vocab_sizes) if decoder_vocab_sizes is not None else [256] * 32
        self.decoding = decoding
        self.decoder_start_token_path = decoder_start_token_path
        self.scaleup_output_hidden = scaleup_output_hidden
        self.shared_output_input_embeds = shared_output_input_embeds
        self.d_model, self.d_kv, self.d_ff = d_model, d_kv, d_ff
        self.num_layers = num_layers
        self.num_decoder_layers = num_decoder_layers if num_decoder_layers is not None else num_layers
        self.num_heads, self.vocab_size = num_heads, vocab_size
        self.relative_attention_num_buckets = relative_attention_num_buckets
        self.relative_attention_max_distance = relative_attention_max_distance
        self.layer_norm_epsilon = layer_norm_epsilon
        self.max_decoder_length = len(self.decoder_vocab_sizes)
        self.tie_word_embeddings = False
        self.is_encoder_decoder = True
        self.decoder_start_token_id = 0
        self.length_penalty = 1.0
        self.extra = kwargs

    @classmethod
    def from_dims(cls, dims: T5Dims) -> "T5forDocIDConfig":
        return cls(decoder_vocab_sizes=[dims.decoder_vocab_size] * dims.docid_len,
                   scaleup_output_hidden=dims.scaleup_output_hidden,
                   shared_output_input_embeds=dims.shared_output_input_embeds, d_model=dims.d_model, d_kv=dims.d_kv,
                   d_ff=dims.d_ff, num_layers=dims.num_layers, num_decoder_layers=dims.num_decoder_layers,
                   num_heads=dims.num_heads, vocab_size=dims.vocab_size,
                   relative_attention_num_buckets=dims.num_buckets,
                   relative_attention_max_distance=dims.max_distance, layer_norm_epsilon=dims.eps)

    @classmethod
    def from_pretrained(cls, path: str) -> "T5forDocIDConfig":
        with open(os.path.join(path, "config.json")) as f:
            return cls(**json.load(f))


class _Engine:
    """One rb200_engine handle plus the shape it was created for."""

    def __init__(self, cfg: T5forDocIDConfig, weights: Dict[str, torch.Tensor], device: int, max_batch: int,
                 max_beams: int, max_src_len: int, precision: str):
        if len(set(cfg.decoder_vocab_sizes)) != 1:
            raise ValueError("not valid decoder_vocab_size")          # reference evaluate.py:433-436
        self.key = (device, max_batch, max_beams, max_src_len, precision)
        self.resizes = 0
        ec = _lib.EngineConfig(cfg.d_model, cfg.num_heads, cfg.d_kv, cfg.d_ff, cfg.num_layers,
                               cfg.num_decoder_layers, cfg.vocab_size, cfg.relative_attention_num_buckets,
                               cfg.relative_attention_max_distance, cfg.layer_norm_epsilon,
                               cfg.decoder_vocab_sizes[0], len(cfg.decoder_vocab_sizes),
                               int(cfg.shared_output_input_embeds), int(cfg.scaleup_output_hidden), max_batch,
                               max_beams, max_src_len, _lib.PRECISIONS[precision], device)
        self.h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.rb200_engine_create(C.byref(ec), C.byref(self.h)))
        stream = _lib.stream_ptr()
        dev = torch.device("cuda", device)
        for name, t in weights.items():
            if name in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "lm_head.weight"):
                continue                                              # ignored-on-load keys (reference :71-78)
            if name.startswith("list_output_embeds.") and cfg.shared_output_input_embeds:
                continue
            td = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            _lib.check(L.rb200_engine_set_weight(self.h, name.encode(), td.data_ptr(), td.numel(), stream))
            torch.cuda.current_stream().synchronize()                 # td may be a temporary copy
        _lib.check(L.rb200_engine_finalize_weights(self.h, stream))

    def __del__(self):
        self.free()

    def free(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().rb200_engine_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def caps(self):
        return self.key[1:4]

    def rows(self, caps=None):
        mb, nb, ms = caps or self.caps
        return mb * nb

    def resize(self, max_batch: int, max_beams: int, max_src_len: int) -> None:
        """New workspace capacities; the packed weights are kept (rb200_engine_resize)."""
        _lib.check(_lib.lib().rb200_engine_resize(self.h, max_batch, max_beams, max_src_len))
        self.key = (self.key[0], max_batch, max_beams, max_src_len, self.key[4])
        self.resizes += 1
        self.host_out_key = None


class T5ForDocIDGeneration:
    """Weights + config of the reference model, runnable only through the CUDA engine."""

    def __init__(self, config: T5forDocIDConfig, state_dict: Dict[str, torch.Tensor]):
        self.config = config
        self._weights = {k: v for k, v in state_dict.items()}
        self._device: Optional[int] = None
        self._engines: Dict[str, _Engine] = {}
        # "auto" = fp16x3 (fp32-grade 3-MMA split on 11-bit fp16 planes, the fastest parity-safe mode) with an
        # automatic re-run in tf32x3 (same mantissa, fp32 exponent range) when a value leaves the fp16 range
        self.precision = os.environ.get("RB200_PRECISION", "auto")
        self.fp16_ok = True

    # -- nn.Module look-alikes the reference callers use -------------------------------------------
    def eval(self):
        return self

    def to(self, device):
        if isinstance(device, int):
            self._device = device
        else:
            d = torch.device(device)
            if d.type != "cuda":
                raise _lib.RB200Error("T5ForDocIDGeneration runs only on CUDA devices (no CPU path exists)")
            self._device = d.index if d.index is not None else torch.cuda.current_device()
        return self

    @property
    def device(self) -> torch.device:
        return torch.device("cuda", self._device if self._device is not None else 0)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self._weights

    def get_engine(self, batch: int, num_beams: int, src_len: int, precision: Optional[str] = None) -> _Engine:
        if not torch.cuda.is_available():
            raise _lib.RB200Error("no CUDA device: the retrieval path has no CPU fallback")
        precision = precision or self.precision
        if precision == "auto":
            raise ValueError("resolve 'auto' with resolve_precision() first")
        dev = self._device if self._device is not None else torch.cuda.current_device()
        # Source lengths come in buckets of 32 positions (the loader pads every batch to its own longest row,
        # dataloader.py:62-79), and an engine serves every batch <= max_batch, num_beams <= max_beams, S <= max_src_len.
        # When a call does not fit, only the workspaces are re-allocated (the packed weights stay): to the union of the
        # old and new shapes if that costs at most 25 % more rows than the larger of the two, else to the new shape.
        src_cap = _round_up(max(src_len, 8), 32)
        e = self._engines.get(precision)
        if e is not None and e.key[0] != dev:
            self.drop_engine(precision)
            e = None
        if e is None:
            with torch.cuda.device(dev):
                e = self._engines[precision] = _Engine(self.config, self._weights, dev, batch, num_beams, src_cap,
                                                       precision)
            return e
        mb, nb, ms = e.caps
        if mb >= batch and nb >= num_beams and ms >= src_len:
            return e
        union = (max(mb, batch), max(nb, num_beams), max(ms, src_cap))
        exact = (batch, num_beams, max(ms, src_cap))
        new = union if union[0] * union[1] <= 1.25 * max(mb * nb, batch * num_beams) else exact
        with torch.cuda.device(dev):
            torch.cuda.synchronize()
            e.resize(*new)
        return e

    def drop_engine(self, precision: str) -> None:
        """Free the engine of one precision mode (e.g. fp16x3 after the automatic fall-back to tf32x3)."""
        e = self._engines.pop(precision, None)
        if e is not None:
            torch.cuda.synchronize()
            e.free()

    # -- the reference's model forward (t5_generative_retriever.py:295-450), teacher-forced on the engine -----------
    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, encoder_outputs=None,
                return_dict=True, precision: Optional[str] = None, **unused):
        """``input_ids`` / ``attention_mask`` [bz, S]; ``decoder_input_ids`` [bz, T] whose column 0 is the decoder start
        token (0 or -1, reference :204) and whose columns 1.. are DocID codes. Returns ``T5GenRetModelOutput`` with
        ``decoder_last_hidden_state`` [bz, T, d_model], ``encoder_last_hidden_state`` [bz, S, d_model] and, when
        ``config.decoding`` is set, ``logits`` = list of T tensors [bz, V] (position i against the i-th output table,
        :250-262). ``past_key_values`` is None: the engine keeps its own KV cache (the reference never consumed
        its cache either, SURVEY.md finding 4)."""
        if input_ids is None or attention_mask is None:
            raise NotImplementedError("the engine runs its own encoder: pass input_ids and attention_mask "
                                      "(encoder_outputs alone is not supported)")
        if decoder_input_ids is None:
            raise ValueError("decoder_input_ids is required (the retrieval path always decodes)")
        dec = decoder_input_ids.to(torch.int64)
        assert dec.dim() == 2 and dec.shape[0] == input_ids.shape[0]
        assert int(dec[0, 0]) in (-1, 0), dec[0, 0]                      # reference :204
        T = dec.shape[1]
        assert T <= len(self.config.decoder_vocab_sizes), "seq_length <= num_decoder_embeds (reference :201)"
        tokens = torch.zeros((dec.shape[0], T), dtype=torch.int32)
        tokens[:, : T - 1] = dec[:, 1:].to(torch.int32).cpu()          # token p feeds position p + 1
        out = self._forced(input_ids, attention_mask, tokens, 1, want_logits=bool(self.config.decoding),
                           want_hidden=True, want_scores=False, precision=precision)
        logits = None
        if out["logits"] is not None:
            logits = [out["logits"][p] for p in range(T)]
        res = T5GenRetModelOutput(logits=logits, past_key_values=None,
                                  decoder_last_hidden_state=out["hidden"].permute(1, 0, 2).contiguous(),
                                  encoder_last_hidden_state=out["encoder"])
        if not return_dict:
            return (res.logits, res.decoder_last_hidden_state, res.encoder_last_hidden_state)
        return res

    __call__ = forward

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, head_mask=None,
                                      decoder_head_mask=None, cross_attn_head_mask=None, use_cache=None,
                                      encoder_outputs=None, **kwargs):
        """reference :452-479. (There ``past`` never reaches ``past_key_values``, so the whole prefix is fed.)"""
        return {"decoder_input_ids": input_ids, "past_key_values": past_key_values,
                "encoder_outputs": encoder_outputs, "attention_mask": attention_mask, "use_cache": use_cache}

    @staticmethod
    def _reorder_cache(past, beam_idx):
        """reference :484-512. The engine addresses its KV cache through a beam ancestry table instead of copying
        it, so there is nothing to reorder; kept for callers that drive their own loop."""
        return past

    def _forced(self, input_ids, attention_mask, tokens: torch.Tensor, rows_per_query: int, want_logits: bool,
                want_hidden: bool, want_scores: bool, precision: Optional[str] = None):
        """rb200_engine_forward: teacher-forced decoder pass. tokens int32 [B*rows_per_query, T]."""
        B, S = input_ids.shape
        R, T = tokens.shape
        assert R == B * rows_per_query
        L = _lib.lib()
        while True:
            mode = self.resolve_precision(precision)
            try:
                eng = self.get_engine(B, rows_per_query, S, mode)
            except ValueError as err:
                if (precision or self.precision) == "auto" and mode == "fp16x3" and "fp16 range" in str(err):
                    self.fp16_ok = False
                    continue
                raise
            dev = torch.device("cuda", eng.key[0])
            with torch.cuda.device(dev):
                ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
                mask = attention_mask.to(device=dev, dtype=torch.int64).contiguous()
                tok = tokens.to(device=dev, dtype=torch.int32).contiguous()
                V, d = self.config.decoder_vocab_sizes[0], self.config.d_model
                logits = torch.empty((T, R, V), dtype=torch.float32, device=dev) if want_logits else None
                hidden = torch.empty((T, R, d), dtype=torch.float32, device=dev) if want_hidden else None
                scores = torch.empty((R,), dtype=torch.float32, device=dev) if want_scores else None
                _lib.check(L.rb200_engine_forward(
                    eng.h, ids.data_ptr(), mask.data_ptr(), B, S, rows_per_query, tok.data_ptr(), T,
                    logits.data_ptr() if want_logits else None, hidden.data_ptr() if want_hidden else None,
                    scores.data_ptr() if want_scores else None, _lib.stream_ptr()))
                p = C.c_void_p()
                _lib.check(L.rb200_engine_encoder_states(eng.h, C.byref(p)))
                enc = _device_view(p.value, B * S * d * 4, dev).view(torch.float32).view(B, S, d).clone()
                probe = scores if want_scores else (logits if want_logits else None)
                if (precision or self.precision) == "auto" and mode == "fp16x3" and probe is not None and \
                        bool(torch.isnan(probe.reshape(-1)[:1]).item()):
                    self.fp16_ok = False                                 # fp16 range overflow: redo in tf32x3
                    self.drop_engine("fp16x3")
                    continue
            return {"logits": logits, "hidden": hidden, "scores": scores, "encoder": enc, "precision": mode}

    def resolve_precision(self, precision: Optional[str] = None) -> str:
        """'auto' -> 'fp16x3' until an fp16 range overflow has been seen on this model, then 'tf32x3'."""
        precision = precision or self.precision
        if precision != "auto":
            return precision
        return "fp16x3" if self.fp16_ok else "tf32x3"

    @classmethod
    def from_pretrained(cls, path: str, config: Optional[T5forDocIDConfig] = None) -> "T5ForDocIDGeneration":
        config = config or T5forDocIDConfig.from_pretrained(path)
        st_path = os.path.join(path, "model.safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file
            sd = load_file(st_path)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        return cls(config, sd)

    def save_pretrained(self, save_dir: str) -> None:
        os.makedirs(save_dir, exist_ok=True)
        c = self.config
        cfg = dict(decoder_vocab_sizes=c.decoder_vocab_sizes, decoding=c.decoding,
                   decoder_start_token_path=c.decoder_start_token_path,
                   scaleup_output_hidden=c.scaleup_output_hidden,
                   shared_output_input_embeds=c.shared_output_input_embeds, d_model=c.d_model, d_kv=c.d_kv,
                   d_ff=c.d_ff, num_layers=c.num_layers, num_decoder_layers=c.num_decoder_layers,
                   num_heads=c.num_heads, vocab_size=c.vocab_size,
                   relative_attention_num_buckets=c.relative_attention_num_buckets,
                   relative_attention_max_distance=c.relative_attention_max_distance,
                   layer_norm_epsilon=c.layer_norm_epsilon)
        with open(os.path.join(save_dir, "config.json"), "w") as f:
            json.dump(cfg, f)
        torch.save({k: v.detach().cpu() for k, v in self._weights.items()}, os.path.join(save_dir, "pytorch_model.bin"))


class T5SeqAQEncoder:
    """reference t5_generative_retriever.py:772-855 (inference surface only)."""

    def __init__(self, model_name_or_path=None, shared_output_input_embeds=None, multi_vocab_sizes=None,
                 base_model: Optional[T5ForDocIDGeneration] = None):
        if base_model is None:
            config = T5forDocIDConfig.from_pretrained(model_name_or_path)
            config.decoding = False
            if shared_output_input_embeds is not None:
                assert shared_output_input_embeds in [False, True]
                config.shared_output_input_embeds = shared_output_input_embeds
            base_model = T5ForDocIDGeneration.from_pretrained(model_name_or_path, config=config)
        self.base_model = base_model
        self.config = base_model.config
        self.model_args = None

    @classmethod
    def from_pretrained(cls, model_name_or_path=None, shared_output_input_embeds=None, multi_vocab_sizes=False):
        return cls(model_name_or_path, shared_output_input_embeds, multi_vocab_sizes)

    @classmethod
    def from_weights(cls, dims: T5Dims, weights: Dict[str, torch.Tensor]) -> "T5SeqAQEncoder":
        return cls(base_model=T5ForDocIDGeneration(T5forDocIDConfig.from_dims(dims), weights))

    def eval(self):
        return self

    def to(self, device):
        self.base_model.to(device)
        return self

    def query_encode(self, **inputs):
        """reference :786-792."""
        text_reps = self.base_model(**inputs).decoder_last_hidden_state
        assert text_reps.size(1) - 1 == 0
        return text_reps[:, 0, :]

    def output_table(self, i: int) -> torch.Tensor:
        sd = self.base_model.state_dict()
        name = "list_decoder_embeds" if self.config.shared_output_input_embeds else "list_output_embeds"
        return sd[f"{name}.{i}.weight"]

    def decode(self, text_encodings: torch.Tensor, summation: bool = False) -> torch.Tensor:
        """reference :812-832: rows of the per-position output tables [bz, smtid_length, d_model]."""
        embeds = [self.output_table(i).to(text_encodings.device)[text_encodings[:, i]].unsqueeze(1)
                  for i in range(text_encodings.size(1))]
        text_embeds = torch.cat(embeds, dim=1)
        return text_embeds.sum(dim=1) if summation else text_embeds

    def rerank_forward(self, **inputs) -> torch.Tensor:
        """reference :794-798: sum over positions of <decoder hidden state, output embedding of the doc's code> = the
        teacher-forced sum of the doc's logits; computed inside the engine (rb200_engine_forward scores)."""
        tq = inputs["tokenized_query"]
        doc = inputs["doc_encoding"].to(torch.int64)
        dec = tq["decoder_input_ids"].to(torch.int64)
        assert dec.shape == doc.shape, (dec.shape, doc.shape)
        assert torch.equal(dec[:, 1:].cpu(), doc[:, :-1].cpu()), \
            "decoder_input_ids must be the start token followed by the doc's first L-1 codes"
        out = self.base_model._forced(tq["input_ids"], tq["attention_mask"], doc.to(torch.int32).cpu(), 1,
                                      want_logits=False, want_hidden=False, want_scores=True)
        return out["scores"]

    def score_docids(self, input_ids, attention_mask, codes: torch.Tensor) -> torch.Tensor:
        """Engine extension of rerank_forward: n candidate DocIDs per query share one encoder pass.
        codes int [B, n, L] -> scores fp32 [B, n]."""
        B, n, L = codes.shape
        out = self.base_model._forced(input_ids, attention_mask, codes.reshape(B * n, L).to(torch.int32).cpu(), n,
                                      want_logits=False, want_hidden=False, want_scores=True)
        return out["scores"].view(B, n)

    def save_pretrained(self, save_dir):
        self.base_model.save_pretrained(save_dir)
