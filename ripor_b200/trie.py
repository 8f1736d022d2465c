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


class DocidStrings:
    """Read-only list of the docid strings of docid_to_smtid.json in file order, kept as one byte buffer + offsets
    (8.8 M Python str objects would cost ~0.6 GB and seconds to build; the output mapping touches a few thousand)."""

    def __init__(self, key_bytes: np.ndarray, offsets: np.ndarray):
        self._bytes, self._off = key_bytes, offsets

    def __len__(self) -> int:
        return len(self._off) - 1

    def __getitem__(self, i: int) -> str:
        if i < 0:
            i += len(self)
        return self._bytes[self._off[i]: self._off[i + 1]].tobytes().decode("utf-8")

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def source_tag(path: str) -> int:
    """64-bit tag of a source file (size, mtime, first/last 64 KB) stored in the trie cache header, so that a cache
    made from another docid_to_smtid.json is rebuilt instead of silently mapping beams to the wrong documents."""
    import hashlib
    import os
    st = os.stat(path)
    h = hashlib.blake2b(digest_size=8)
    h.update(f"{st.st_size}:{st.st_mtime_ns}".encode())
    with open(path, "rb") as f:
        h.update(f.read(65536))
        if st.st_size > 65536:
            f.seek(max(st.st_size - 65536, 0))
            h.update(f.read(65536))
    return int.from_bytes(h.digest(), "little") or 1


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
    def from_json(cls, path: str, V: int, max_new_token_for_docid: Optional[int] = None, n_threads: int = 0
                  ) -> "DocidTrie":
        """docid_to_smtid.json -> trie through the streaming C++ reader (rb200_docid_json_open): no Python dict of
        8.8 M lists (the reference's ujson.load + dict walk, evaluate.py:400-446)."""
        L = _lib.lib()
        tab = C.c_void_p()
        _lib.check(L.rb200_docid_json_open(path.encode(), int(max_new_token_for_docid or 0), C.byref(tab)))
        try:
            h = C.c_void_p()
            _lib.check(L.rb200_trie_build_from_table(tab, V, n_threads, C.byref(h)))
            n, key_bytes = C.c_int64(), C.c_int64()
            _lib.check(L.rb200_docid_json_info(tab, C.byref(n), None, None, C.byref(key_bytes)))
            kb, ko = C.c_void_p(), C.c_void_p()
            _lib.check(L.rb200_docid_json_keys(tab, C.byref(kb), C.byref(ko)))
            keys = np.ctypeslib.as_array(C.cast(kb, C.POINTER(C.c_uint8)), shape=(max(key_bytes.value, 1),)).copy()
            offs = np.ctypeslib.as_array(C.cast(ko, C.POINTER(C.c_int64)), shape=(n.value + 1,)).copy()
        finally:
            L.rb200_docid_json_free(tab)
        return cls(h, DocidStrings(keys, offs))

    @staticmethod
    def read_json_codes(path: str, max_new_token_for_docid: Optional[int] = None):
        """(codes int32 [N, L], DocidStrings) of a docid_to_smtid.json, via the C++ reader."""
        L = _lib.lib()
        tab = C.c_void_p()
        _lib.check(L.rb200_docid_json_open(path.encode(), int(max_new_token_for_docid or 0), C.byref(tab)))
        try:
            n, ln, key_bytes = C.c_int64(), C.c_int32(), C.c_int64()
            _lib.check(L.rb200_docid_json_info(tab, C.byref(n), C.byref(ln), None, C.byref(key_bytes)))
            cp, kb, ko = C.c_void_p(), C.c_void_p(), C.c_void_p()
            _lib.check(L.rb200_docid_json_codes(tab, C.byref(cp)))
            _lib.check(L.rb200_docid_json_keys(tab, C.byref(kb), C.byref(ko)))
            codes = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_int32)), shape=(n.value, ln.value)).copy()
            keys = np.ctypeslib.as_array(C.cast(kb, C.POINTER(C.c_uint8)), shape=(max(key_bytes.value, 1),)).copy()
            offs = np.ctypeslib.as_array(C.cast(ko, C.POINTER(C.c_int64)), shape=(n.value + 1,)).copy()
        finally:
            L.rb200_docid_json_free(tab)
        return codes, DocidStrings(keys, offs)

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
    def load(cls, path: str, docids: Optional[List[str]] = None, expect_tag: Optional[int] = None) -> "DocidTrie":
        """Read a trie cache. ``expect_tag`` (see ``source_tag``) and ``docids`` are checked against the header: a
        cache made from another source raises ValueError (callers rebuild)."""
        h, tag = C.c_void_p(), C.c_uint64()
        _lib.check(_lib.lib().rb200_trie_load_tagged(path.encode(), C.byref(tag), C.byref(h)))
        t = cls(h, docids)
        if expect_tag is not None and tag.value != expect_tag:
            raise ValueError(f"{path} was built from another docid_to_smtid.json (tag {tag.value:#x} != {expect_tag:#x})")
        if docids is not None and len(docids) != t.n_docs:
            raise ValueError(f"{path} holds {t.n_docs} documents, the docid table {len(docids)}")
        return t

    def save(self, path: str, tag: int = 0) -> None:
        """Atomic (temporary file + rename) write of the cache with the source tag in its header."""
        _lib.check(_lib.lib().rb200_trie_save_tagged(self._h, path.encode(), tag))

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
            with torch.cuda.device(ids.device):          # the C ABI uses the tables of the current device
                self.upload(torch.cuda.current_device())
                _lib.check(_lib.lib().rb200_trie_mask_device(self._h, ids.data_ptr(), R, T, out.data_ptr(),
                                                             _lib.stream_ptr()))
            return out
        out = torch.empty((R, self.V), dtype=torch.float64)
        _lib.check(_lib.lib().rb200_trie_mask_host(self._h, ids.data_ptr(), R, T, out.data_ptr()))
        return out

    def expand_ranges(self, leaf_ranges, max_docs_per_row: int = 8):
        """Device-side smtid -> docids mapping (rb200_trie_leaf_expand). leaf_ranges: CUDA int32 [n, 2] as returned by
        the search; returns (doc rows int64 [n, k] in json order padded with -1, true counts int32 [n])."""
        import torch
        assert leaf_ranges.is_cuda and leaf_ranges.dtype == torch.int32
        lr = leaf_ranges.contiguous()
        n = lr.shape[0]
        docs = torch.empty((n, max_docs_per_row), dtype=torch.int64, device=lr.device)
        counts = torch.empty((n,), dtype=torch.int32, device=lr.device)
        with torch.cuda.device(lr.device):               # the C ABI uses the tables of the current device
            self.upload(torch.cuda.current_device())
            _lib.check(_lib.lib().rb200_trie_leaf_expand(self._h, lr.data_ptr(), n, max_docs_per_row, docs.data_ptr(),
                                                         counts.data_ptr(), _lib.stream_ptr()))
        return docs, counts

    def docid_of_row(self, row: int) -> str:
        return str(int(row)) if self.docids is None else self.docids[int(row)]

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
