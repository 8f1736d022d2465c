"""ORACLE (test infrastructure, never on the product path): the allowed-next-token mask of
``PrefixConstrainLogitProcessorFastSparse.__call__`` (reference ``t5_pretrainer/tasks/generation.py:603-677``)
restated as a binary search over the lexicographically sorted code rows.

Why a second restatement: ``oracle/beam.py::TrieMaskOracle`` follows the reference's dict-of-strings construction
literally and is pinned against the literal class (tests/test_oracle_literal.py), but at 8.8 M documents x 32 codes
that dict needs ~2.6e8 string keys (tens of GB) and cannot be built on a test box. This one
needs 283 MB, shares no code and no data structure with the product's flattened trie (csrc/trie.cu), and is pinned
against ``TrieMaskOracle`` (hence against the reference) on random tries in tests/test_host_logic.py. The parity
checks at the benchmarked size use it as the mask so that the product is never its own oracle.

Semantics restated: the prefix ``input_ids[r, 1:]`` (column 0 is the decoder start token) selects the code rows that
start with it; the allowed tokens are the distinct values of the next column over those rows; a prefix that no row
starts with gives an all-zero row (generation.py:656-661,675).
"""
from __future__ import annotations

import numpy as np
import torch


class SortedCodesMask:
    def __init__(self, codes: np.ndarray, vocab_size: int):
        codes = np.asarray(codes)
        assert codes.ndim == 2
        self.L, self.V = codes.shape[1], vocab_size
        # big-endian fixed-width bytes: byte-wise lexicographic order == row-wise numeric order
        self.width = 1 if vocab_size <= 256 else 2
        be = codes.astype(">u1" if self.width == 1 else ">u2")
        keys = np.ascontiguousarray(be).view(f"S{self.L * self.width}").reshape(-1)
        order = np.argsort(keys, kind="stable")
        self.keys = keys[order]
        self.sorted_codes = np.ascontiguousarray(codes[order])
        self._cache = {}

    def _key(self, prefix: np.ndarray, fill: int) -> bytes:
        row = np.full(self.L, fill, dtype=np.int64)
        row[: len(prefix)] = prefix
        return row.astype(">u1" if self.width == 1 else ">u2").tobytes()

    def ranges(self, prefixes: np.ndarray):
        """[lo, hi) into the sorted rows for every prefix (rows of an int array [R, t])."""
        R, t = prefixes.shape
        top = 255 if self.width == 1 else 65535
        lo_keys = np.array([self._key(p, 0) for p in prefixes], dtype=self.keys.dtype)
        hi_keys = np.array([self._key(p, top) for p in prefixes], dtype=self.keys.dtype)
        lo = np.searchsorted(self.keys, lo_keys, side="left")
        hi = np.searchsorted(self.keys, hi_keys, side="right")
        ok = (prefixes >= 0).all(axis=1) & (prefixes < self.V).all(axis=1) if t else np.ones(R, dtype=bool)
        hi = np.where(ok, hi, lo)
        return lo, hi

    def __call__(self, input_ids: torch.Tensor, scores=None) -> torch.Tensor:
        ids = input_ids.cpu().numpy()
        R, T = ids.shape
        t = T - 1
        assert t < self.L
        lo, hi = self.ranges(ids[:, 1:].astype(np.int64))
        mask = np.zeros((R, self.V), dtype=np.float64)
        for r in range(R):
            if hi[r] > lo[r]:
                key = (int(lo[r]), int(hi[r]), t)
                allowed = self._cache.get(key)
                if allowed is None:
                    allowed = np.unique(self.sorted_codes[lo[r]: hi[r], t])
                    if hi[r] - lo[r] > 4096:                 # the few big ranges near the root repeat across rows
                        self._cache[key] = allowed
                mask[r, allowed] = 1.0
        return torch.from_numpy(mask)
