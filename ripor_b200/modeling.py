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
        try:
            if getattr(self, "h", None):
                _lib.lib().rb200_engine_free(self.h)
                self.h = None
        except Exception:
            pass


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
        e = self._engines.get(precision)
        if e is not None:
            d0, mb, nb, ms, pr = e.key
            if d0 == dev and nb == num_beams and mb >= batch and ms >= src_len:
                return e
            del self._engines[precision]
            del e
            torch.cuda.synchronize()
        with torch.cuda.device(dev):
            self._engines[precision] = _Engine(self.config, self._weights, dev, batch, num_beams, max(src_len, 8),
                                               precision)
        return self._engines[precision]

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

    def save_pretrained(self, save_dir):
        self.base_model.save_pretrained(save_dir)
