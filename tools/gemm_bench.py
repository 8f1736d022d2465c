"""Per-shape GEMM microbench through the C ABI (rb200_gemm_bench): device time per launch and fraction of the
measured dense peak for the decoder-step shapes of the bench workload."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ripor_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp16x3")
ap.add_argument("--M", type=int, default=2560)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--rotate-mb", type=int, default=300)
ap.add_argument("--shapes", default="")
a = ap.parse_args()
torch.cuda.init()
lib = _lib.lib()
peak = 1400.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass
shapes = [("qkv", 2304, 768, 0), ("o", 768, 768, 1), ("cq", 768, 768, 0), ("wi", 3072, 768, 2), ("wo", 768, 3072, 1),
          ("lmhead", 256, 768, 0), ("ckv", 18432, 768, 0)]
if a.shapes:
    shapes = [tuple([s.split(":")[0]] + [int(v) for v in s.split(":")[1:]]) for s in a.shapes.split(",")]
mult = 3 if a.precision.endswith("x3") else 1
tot = 0.0
for name, N, K, epi in shapes:
    us = C.c_double()
    _lib.check(lib.rb200_gemm_bench(_lib.PRECISIONS[a.precision], a.M, N, K, epi, a.iters, a.rotate_mb, C.byref(us),
                                    _lib.stream_ptr()))
    fl = 2.0 * a.M * N * K
    print(f"{name:8s} M={a.M} N={N:5d} K={K:5d} epi={epi}: {us.value:7.2f} us  alg {fl / us.value / 1e6:7.1f} TF/s  "
          f"issued {mult * fl / us.value / 1e6:7.1f} TF/s = {mult * fl / us.value / 1e6 / peak:5.1%} of {peak:.0f}")
