"""How many ranked DocID lists change between GEMM precisions at the bench batch size?

Runs the same B queries through the engine in the exact fp32 (FFMA) mode and in each tensor-core mode,
counts queries whose ranked list differs from fp32's and the largest score difference, and (optionally)
checks the first --oracle-queries against the CPU oracle. Writes a JSON summary."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ripor_b200 import synthetic as syn  # noqa: E402
from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search  # noqa: E402
from ripor_b200.modeling import T5SeqAQEncoder  # noqa: E402
from ripor_b200.trie import DocidTrie  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--beams", type=int, default=10)
ap.add_argument("--docs", type=int, default=8841823)
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--modes", default="tf32x3,fp16x3,bf16x3,tf32,bf16")
ap.add_argument("--seeds", default="77,78,79")
ap.add_argument("--out", default="gpurun_out/parity_sweep.json")
a = ap.parse_args()
dims = syn.T5Dims.t5_base(docid_len=a.L)
w = syn.make_weights(dims)
trie = DocidTrie.from_codes(syn.make_codes(a.docs, a.L, 256), 256)
proc = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
res = {m: {"queries": 0, "lists_differ": 0, "top_set_differs": 0, "max_score_diff": 0.0} for m in a.modes.split(",")}
for seed in [int(x) for x in a.seeds.split(",")]:
    ids, mask = syn.make_queries(a.batch, S=32, seed=seed)
    ids, mask = ids.cuda(), mask.cuda()
    outs = {}
    for mode in ["fp32"] + a.modes.split(","):
        model = T5SeqAQEncoder.from_weights(dims, w).to("cuda:0")
        o = generate_for_constrained_prefix_beam_search(model.base_model, proc, input_ids=ids, attention_mask=mask,
                                                        max_new_tokens=a.L, num_beams=a.beams,
                                                        num_return_sequences=a.beams, output_scores=True,
                                                        return_dict_in_generate=True, precision=mode)
        torch.cuda.synchronize()
        outs[mode] = (o.sequences.view(a.batch, a.beams, -1).cpu(), o.sequences_scores.view(a.batch, a.beams).cpu())
        del model
        torch.cuda.empty_cache()
    rs, rc = outs["fp32"]
    for mode in a.modes.split(","):
        s, c = outs[mode]
        differ = (s != rs).any(-1).any(-1)
        setd = torch.tensor([set(map(tuple, s[b].tolist())) != set(map(tuple, rs[b].tolist())) for b in range(a.batch)])
        same = ~differ
        r = res[mode]
        r["queries"] += a.batch
        r["lists_differ"] += int(differ.sum())
        r["top_set_differs"] += int(setd.sum())
        if same.any():
            r["max_score_diff"] = max(r["max_score_diff"], float((c[same] - rc[same]).abs().max()))
    print(seed, json.dumps(res), flush=True)
json.dump({"args": vars(a), "vs": "fp32 FFMA mode of the same engine", "results": res}, open(a.out, "w"), indent=1)
