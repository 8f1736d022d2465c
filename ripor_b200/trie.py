"""DocID trie handle: host-side mirror of the structures the reference builds in Python.

Stands in for three reference objects at once (reference t5_pretrainer/evaluate.py:400-446):
``list_smtid_to_nextids`` (per-level prefix -> next ids dicts), the CSR tables inside
``PrefixConstrainLogitProcessorFastSparse`` (tasks/generation.py:604-642) and ``smtid_to_docids``.
All tables live behind the C ABI (``rb200_trie_*`` in include/riporb200.h).
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib


class DocidTrie:
    def __init__(self, handle: C.c_void_p, docids: Optional[List[str]] = None):
        self._h = handle
        self.docids = docids          # row index -> docid string (json order); None = str(row)
        info = _lib.TrieInfo()
        _lib.check(_lib.lib().rb200_trie_get_info(self._h, C.byref(info)))
        self.L, self.V = info.L, info.V
        self.n_docs, self.n_unique = info.n_docs, info.n_unique

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_codes(cls, codes: np.ndarray, V: int, docids: Optional[List[str]] = None, n_threads: int = 0
                   ) -> "DocidTrie":
        """codes[N, L] integer array; row i belongs to docids[i] (or str(i))."""
        codes = np.asarray(codes)
        if codes.ndim != 2:
            raise ValueError("codes must be [n_docs, L]")
        if codes.size and (codes.min() < 0 or codes.max() >= V):
            raise ValueError(f"codes must lie in [0, {V})")
        dt = np.uint8 if V <= 256 else np.uint16
        arr = np.ascontiguousarray(codes.astype(dt, copy=False))
        h = C.c_void_p()
        _lib.check(_lib.lib().rb200_trie_build(arr.ctypes.data, arr.itemsize, arr.shape[0], arr.shape[1], V,
                                               n_threads, C.byref(h)))
        if docids is not None and len(docids) != arr.shape[0]:
            raise ValueError("docids and codes disagree on the number of documents")
        return cls(h, docids)

    @classmethod
    def from_docid_to_smtid(cls, docid_to_smtids: Dict[str, Sequence[int]], V: int,
                            max_new_token_for_docid: Optional[int] = None) -> "DocidTrie":
        """The dict of docid_to_smtid.json: {docid: [-1, c1..cL]} (evaluate.py:400-401,441)."""
        docids = list(docid_to_smtids.keys())
        first = docid_to_smtids[docids[0]]
        L = len(first) - 1 if max_new_token_for_docid is None else max_new_token_for_docid
        codes = np.empty((len(docids), L), dtype=np.int64)
        for i, d in enumerate(docids):
            smt = docid_to_smtids[d]
            assert smt[0] == -1, smt
            codes[i] = smt[1: 1 + L]
        return cls.from_codes(codes, V, docids)

    @classmethod
    def from_json(cls, path: str, V: int, max_new_token_for_docid: Optional[int] = None) -> "DocidTrie":
        with open(path) as f:
            return cls.from_docid_to_smtid(json.load(f), V, max_new_token_for_docid)

    @classmethod
    def from_list_smtid_to_nextids(cls, list_smtid_to_nextids: List[Dict[str, Iterable[int]]], V: int) -> "DocidTrie":
        """The reference's pickle format (aq_preprocess/build_list_smtid_to_nextids.py:23-41): the last
        level's keys and values spell every distinct full code."""
        last = list_smtid_to_nextids[-1]
        rows = []
        for key, nxt in last.items():
            prefix = [int(x) for x in key.split("_")[1:]]
            for n in nxt:
                rows.append(prefix + [int(n)])
        return cls.from_codes(np.asarray(rows, dtype=np.int64).reshape(len(rows), len(list_smtid_to_nextids)), V)

    @classmethod
    def load(cls, path: str, docids: Optional[List[str]] = None) -> "DocidTrie":
        h = C.c_void_p()
        _lib.check(_lib.lib().rb200_trie_load(path.encode(), C.byref(h)))
        return cls(h, docids)

    def save(self, path: str) -> None:
        _lib.check(_lib.lib().rb200_trie_save(self._h, path.encode()))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().rb200_trie_free(self._h)
                self._h = None
        except Exception:
            pass

    # ---- queries ------------------------------------------------------------------------------
    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def info(self) -> _lib.TrieInfo:
        info = _lib.TrieInfo()
        _lib.check(_lib.lib().rb200_trie_get_info(self._h, C.byref(info)))
        return info

    def level_counts(self) -> List[int]:
        out = np.zeros(self.L, dtype=np.int64)
        _lib.check(_lib.lib().rb200_trie_level_counts(self._h, out.ctypes.data))
        return out.tolist()

    def upload(self, device: int = 0) -> "DocidTrie":
        _lib.check(_lib.lib().rb200_trie_upload(self._h, device))
        return self

    def mask(self, input_ids):
        """valid_mask float64 [R, V] for decoder prefixes input_ids [R, T] (torch tensor, host or CUDA)."""
        import torch
        ids = input_ids.to(torch.int64).contiguous()
        R, T = ids.shape
        if ids.is_cuda:
            out = torch.empty((R, self.V), dtype=torch.float64, device=ids.device)
            self.upload(ids.device.index or 0)
            _lib.check(_lib.lib().rb200_trie_mask_device(self._h, ids.data_ptr(), R, T, out.data_ptr(),
                                                         _lib.stream_ptr()))
            return out
        out = torch.empty((R, self.V), dtype=torch.float64)
        _lib.check(_lib.lib().rb200_trie_mask_host(self._h, ids.data_ptr(), R, T, out.data_ptr()))
        return out

    def leaf_rows(self, leaf: int) -> np.ndarray:
        ptr, n = C.c_void_p(), C.c_int64()
        _lib.check(_lib.lib().rb200_trie_leaf_docs(self._h, leaf, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.empty(0, dtype=np.int64)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(n.value,)).copy()

    def rows_for_range(self, lo: int, hi: int) -> np.ndarray:
        """Input rows under the leaves [lo, hi) in input (json) order, as evaluate.py:439-446 lists them."""
        if hi <= lo:
            return np.empty(0, dtype=np.int64)
        rows = np.concatenate([self.leaf_rows(u) for u in range(lo, hi)])
        rows.sort()
        return rows

    def docids_for_range(self, lo: int, hi: int) -> List[str]:
        rows = self.rows_for_range(lo, hi)
        if self.docids is None:
            return [str(int(r)) for r in rows]
        return [self.docids[int(r)] for r in rows]

    def find_leaf(self, code: Sequence[int]) -> int:
        arr = np.ascontiguousarray(np.asarray(code, dtype=np.int32))
        assert arr.shape == (self.L,)
        out = C.c_int64()
        _lib.check(_lib.lib().rb200_trie_find_leaf(self._h, arr.ctypes.data, C.byref(out)))
        return out.value
