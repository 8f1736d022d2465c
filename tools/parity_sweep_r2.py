"""Parity sweep at the benchmarked size against the CPU oracle (not against another engine mode): BASELINE configs[1]
(t5-base, 8,841,823-doc trie, beam 10, L = 32) on the uniform and the Zipf-skewed trie, N queries per trie, in the
parity precisions. The mask of the oracle comes from oracle/range_mask.py (independent of the product's trie). Queries
whose oracle run had a near-tie at the beam cut (< 2e-4 between the last kept and the first dropped candidate) are
reported apart from real mismatches (hard part 1). Writes profiles/parity_sweep_r02.json.

    python tools/parity_sweep_r2.py --queries 256
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import beam as ob, t5_math  # noqa: E402
from oracle.range_mask import SortedCodesMask  # noqa: E402
from ripor_b200 import synthetic as syn  # noqa: E402
from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search  # noqa: E402
from ripor_b200.modeling import T5SeqAQEncoder  # noqa: E402
from ripor_b200.trie import DocidTrie  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=256)
ap.add_argument("--docs", type=int, default=8841823)
ap.add_argument("--precisions", default="fp16x3,tf32x3")
a = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)
L, V, nb, B = 32, 256, 10, 256
dims = syn.T5Dims.t5_base(docid_len=L)
w = syn.make_weights(dims)
model = T5SeqAQEncoder.from_weights(dims, w).to("cuda:0")
report = {"workload": f"t5-base, {a.docs:,}-doc trie (32x256), beam 10, batch 256, {a.queries} queries per trie",
          "oracle": "KV-cached fp32 CPU oracle + oracle/range_mask.py", "tries": {}}
for kind in ("uniform", "zipf"):
    codes = syn.make_codes(a.docs, L, V, skew=(kind == "zipf"))
    mask_fn = SortedCodesMask(codes, V)
    proc = PrefixConstrainLogitProcessorFastSparse.from_trie(DocidTrie.from_codes(codes, V))
    per = {p: {"queries": 0, "lists_exact": 0, "near_tie": 0, "real_mismatch": 0, "max_abs_score_diff": 0.0}
           for p in a.precisions.split(",")}
    gaps = []
    for off in range(0, a.queries, B):
        n = min(B, a.queries - off)
        ids, mask = syn.make_queries(B, S=32, seed=5000 + off)
        trace = []
        t0 = time.time()
        with torch.no_grad():
            enc = t5_math.encoder_forward(w, dims, ids[:n], mask[:n])
            dec = t5_math.CachedDecoder(w, dims, enc, mask[:n], nb)

            def step(dec_ids, bi):
                if bi is not None:
                    dec.reorder(bi)
                return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])
            ref_seq, ref_sc = ob.beam_search_oracle(step, mask_fn, n, nb, L, trace=trace)
        gaps.append(torch.stack([t["cut_gap"] for t in trace], 0).min(0).values)
        print(f"{kind}: oracle for {n} queries in {time.time() - t0:.1f} s", flush=True)
        for p in per:
            out = generate_for_constrained_prefix_beam_search(
                model.base_model, proc, input_ids=ids.to("cuda:0"), attention_mask=mask.to("cuda:0"), max_new_tokens=L,
                num_beams=nb, num_return_sequences=nb, output_scores=True, return_dict_in_generate=True, precision=p)
            seqs = out.sequences.view(B, nb, L + 1)[:n].reshape(n * nb, L + 1)
            sc = out.sequences_scores.view(B, nb)[:n].reshape(-1)
            real, near = helpers.compare_ranked_near_tie(seqs, sc, ref_seq, ref_sc, nb, trace)
            exact = int((seqs.cpu().view(n, -1) == ref_seq.view(n, -1)).all(1).sum())
            same = (seqs.cpu() == ref_seq).all(1)
            diff = float((sc.cpu()[same] - ref_sc[same]).abs().max()) if same.any() else 0.0
            per[p]["queries"] += n
            per[p]["lists_exact"] += exact
            per[p]["near_tie"] += near
            per[p]["real_mismatch"] += real
            per[p]["max_abs_score_diff"] = max(per[p]["max_abs_score_diff"], diff)
    g = torch.cat(gaps)
    report["tries"][kind] = {"per_precision": per, "min_cut_gap": float(g.min()),
                             "queries_with_cut_gap_below_2e-4": int((g < 2e-4).sum())}
    del proc, mask_fn, codes
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "parity_sweep_r02.json"), "w") as f:
    json.dump(report, f, indent=1)
print(json.dumps(report))
