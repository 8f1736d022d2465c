"""ORACLE (test infrastructure, never on the product path): the on-disk input formats of the retrieval path as the
reference reads them.

  * ``load_docid_to_smtid``: ``ujson.load`` of docid_to_smtid.json (reference t5_pretrainer/evaluate.py:400-401) and the
    assertions the reference makes on it (:441 ``smtids[0] == -1``), returned as (docids in file order, codes matrix).
  * ``unpack_bitstring_codes``: the residual-quantiser code unpacking of
    aq_preprocess/create_customized_smtid_file.py:38-45 (``faiss.BitstringReader(code, code_size)`` + ``read(bits)``
    M times). faiss (requirements: faiss-gpu, unpinned, not vendored under /root/reference and not installed here) packs
    and reads bit strings LSB first: bit i of a vector's stream is bit (i & 7) of byte (i >> 3)
    (faiss/utils/hamming.h, BitstringWriter/BitstringReader). numpy's ``unpackbits(bitorder="little")`` is that order.
    Parity is pinned on hand-computed vectors in tests/test_host_round2.py (no faiss golden data exists in the reference).
"""
from __future__ import annotations

import json
from typing import List, Tuple

import numpy as np


def load_docid_to_smtid(path: str, max_new_token_for_docid: int = 0) -> Tuple[List[str], np.ndarray]:
    with open(path) as fin:
        d = json.load(fin)
    docids = list(d.keys())
    L = len(d[docids[0]]) - 1
    if max_new_token_for_docid:
        L = min(L, max_new_token_for_docid)
    codes = np.empty((len(docids), L), dtype=np.int32)
    for i, k in enumerate(docids):
        smtids = d[k]
        assert smtids[0] == -1, smtids
        codes[i] = smtids[1: 1 + L]
    return docids, codes


def unpack_bitstring_codes(packed: np.ndarray, M: int, bits: int) -> np.ndarray:
    """packed uint8 [n, code_size] -> int32 [n, M]."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    stream = np.unpackbits(packed, axis=1, bitorder="little")[:, : M * bits].reshape(packed.shape[0], M, bits)
    weights = (1 << np.arange(bits, dtype=np.int64))
    return (stream.astype(np.int64) * weights).sum(axis=2).astype(np.int32)
