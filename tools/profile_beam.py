"""The trie / top-k beam-step kernel alone (bench.py's `trie_topk` workload: 16 384 queries x beam 10 over the 8.8 M-doc
trie, random logits) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ripor_b200 import _lib, synthetic as syn  # noqa: E402
from ripor_b200.trie import DocidTrie  # noqa: E402

nb, L, V, Rq = 10, 32, 256, 1 << 14
lib = _lib.lib()
trie = DocidTrie.from_codes(syn.make_codes(8841823, L, V), V).upload(0)
hb = C.c_void_p()
_lib.check(lib.rb200_beam_create(0, Rq, nb, L, V, C.byref(hb)))
big = torch.randn((Rq * nb, V), device="cuda:0")
sp = _lib.stream_ptr()
_lib.check(lib.rb200_beam_reset(hb, trie.handle, Rq, sp))
_lib.check(lib.rb200_beam_step(hb, trie.handle, big.data_ptr(), 1, 0, None, None, 0, sp))
_lib.check(lib.rb200_beam_step(hb, trie.handle, big.data_ptr(), nb, 0, None, None, 0, sp))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    _lib.check(lib.rb200_beam_step(hb, trie.handle, big.data_ptr(), nb, 0, None, None, 0, sp))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
lib.rb200_beam_free(hb)
