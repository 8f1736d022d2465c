"""Times the trie / top-k beam-step kernel alone, step by step (bench.py's `trie_topk` workload: 16 384 queries x beam 10
over the 8.8 M-doc trie, random logits): CUDA events around every launch, so the early steps (hundreds of allowed
children per beam) and the late ones (one or two) can be told apart.
    python tools/beam_bench.py [--nb 10] [--V 256] [--queries 16384] [--steps 6] [--reps 5]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ripor_b200 import _lib, synthetic as syn  # noqa: E402
from ripor_b200.trie import DocidTrie  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=10)
ap.add_argument("--V", type=int, default=256)
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--docs", type=int, default=8841823)
ap.add_argument("--queries", type=int, default=1 << 14)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--log-softmax", type=int, default=0)
a = ap.parse_args()

lib = _lib.lib()
trie = DocidTrie.from_codes(syn.make_codes(a.docs, a.L, a.V), a.V).upload(0)
hb = C.c_void_p()
_lib.check(lib.rb200_beam_create(0, a.queries, a.nb, a.L, a.V, C.byref(hb)))
big = torch.randn((a.queries * a.nb, a.V), device="cuda:0")
sp = _lib.stream_ptr()
per_step = [[] for _ in range(a.steps)]
for _ in range(a.reps + 1):
    _lib.check(lib.rb200_beam_reset(hb, trie.handle, a.queries, sp))
    for t in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rb200_beam_step(hb, trie.handle, big.data_ptr(), 1 if t == 0 else a.nb, a.log_softmax, None, None,
                                       0, sp))
        e1.record()
        torch.cuda.synchronize()
        per_step[t].append(e0.elapsed_time(e1) * 1e3)
lib.rb200_beam_free(hb)
us = [sorted(x[1:])[len(x[1:]) // 2] for x in per_step]          # median without the first (cold) repetition
row_bytes = a.V * 4 + (8 + 16) + (8 + 16 + 4 + 4) + 2 * a.L * 4 * 2
print(json.dumps({"nb": a.nb, "V": a.V, "queries": a.queries, "us_per_step": [round(u, 1) for u in us],
                  "gbs_steps_1_4": round(a.queries * a.nb * row_bytes / (sum(us[1:5]) / max(len(us[1:5]), 1)) / 1e3, 1)}))
