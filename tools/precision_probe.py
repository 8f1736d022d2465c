"""Which tensor-core arithmetic keeps top-k DocID parity with the fp32 oracle?

Emulates on CPU the operand roundings of the candidate tcgen05 GEMM modes (bf16, tf32, and the
error-compensated split forms bf16x3 / tf32x3 that issue three MMAs per k-block) inside the KV-cached
oracle decoder, runs the full constrained beam search, and counts how many queries keep a bit-exact
ranked smtid list and how far the scores move. fp64 linears give the noise floor of fp32 itself.
Test/design tooling: uses oracle/, never imported by the product.
"""
import argparse
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import beam as ob, t5_math  # noqa: E402
from ripor_b200 import synthetic as syn  # noqa: E402


def round_tf32(x):
    b = x.contiguous().view(torch.int32)
    b = (b + 0xFFF + ((b >> 13) & 1)) & ~0x1FFF
    return b.view(torch.float32)


def split(x, rnd):
    hi = rnd(x)
    return hi, rnd(x - hi)


def make_linear(mode):
    bf = lambda x: x.bfloat16().float()
    cache = {}

    def wsplit(W, rnd):
        k = (id(W), rnd)
        if k not in cache:
            cache[k] = split(W, rnd)
        return cache[k]

    if mode == "fp32":
        return t5_math._lin
    if mode == "fp64":
        return lambda x, W: (x.double() @ W.double().t()).float()
    if mode in ("bf16", "tf32"):
        rnd = bf if mode == "bf16" else round_tf32
        return lambda x, W: rnd(x) @ wsplit(W, rnd)[0].t()
    if mode in ("bf16x3", "tf32x3"):
        rnd = bf if mode == "bf16x3" else round_tf32

        def lin(x, W):
            xh, xl = split(x, rnd)
            wh, wl = wsplit(W, rnd)
            return (xl @ wh.t() + xh @ wl.t()) + xh @ wh.t()
        return lin
    raise ValueError(mode)


def run(mode, w, dims, ids, mask, lst, nb, L, V):
    lin = make_linear(mode)
    B = ids.shape[0]
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask, linear=lin)
        dec = t5_math.CachedDecoder(w, dims, enc, mask, nb, linear=lin)

        def step(dec_ids, bi):
            if bi is not None:
                dec.reorder(bi)
            return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])
        trace = []
        seqs, scores = ob.beam_search_oracle(step, ob.TrieMaskOracle(lst, V), B, nb, L, trace=trace)
    return seqs.view(B, nb, L + 1), scores.view(B, nb), trace


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--docs", type=int, default=200000)
    ap.add_argument("--nb", type=int, default=10)
    ap.add_argument("--L", type=int, default=32)
    ap.add_argument("--modes", default="fp64,tf32x3,bf16x3,tf32,bf16")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    dims = syn.T5Dims.t5_base(docid_len=a.L)
    w = syn.make_weights(dims)
    V = dims.decoder_vocab_size
    codes = syn.make_codes(a.docs, a.L, V)
    lst = ob.build_list_smtid_to_nextids(syn.codes_to_docid_to_smtid(codes))
    ids, mask = syn.make_queries(a.batch, S=32)
    t0 = time.time()
    ref_seq, ref_sc, ref_tr = run("fp32", w, dims, ids, mask, lst, a.nb, a.L, V)
    print(f"fp32 oracle: {time.time() - t0:.1f}s", flush=True)
    res = {}
    for mode in a.modes.split(","):
        t0 = time.time()
        seq, sc, tr = run(mode, w, dims, ids, mask, lst, a.nb, a.L, V)
        same_q = (seq == ref_seq).all(-1).all(-1)
        set_q = torch.tensor([set(map(tuple, seq[b].tolist())) == set(map(tuple, ref_seq[b].tolist()))
                              for b in range(a.batch)])
        valid = ref_sc > -1e6
        dsc = (sc - ref_sc).abs()[valid & same_q[:, None].expand_as(valid)]
        lg = max((x["processed"] - y["processed"]).abs()[y["processed"] > -1e8].max().item()
                 for x, y in zip(tr[:1], ref_tr[:1]))
        res[mode] = {"queries": a.batch, "ranked_list_exact": int(same_q.sum()), "top_set_exact": int(set_q.sum()),
                     "max_score_diff_on_exact": float(dsc.max()) if dsc.numel() else None,
                     "step0_max_logit_diff": lg, "seconds": round(time.time() - t0, 1)}
        print(mode, json.dumps(res[mode]), flush=True)
    if a.out:
        json.dump({"args": vars(a), "results": res}, open(a.out, "w"), indent=1)
