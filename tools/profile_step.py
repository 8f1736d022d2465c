"""One search of the bench workload between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ripor_b200 import synthetic as syn  # noqa: E402
from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search  # noqa: E402
from ripor_b200.modeling import T5SeqAQEncoder  # noqa: E402
from ripor_b200.trie import DocidTrie  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32x3")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--beams", type=int, default=10)
ap.add_argument("--docs", type=int, default=8841823)
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--steps-only", type=int, default=0, help="profile only this many decoder steps (0 = whole search)")
a = ap.parse_args()
dims = syn.T5Dims.t5_base(docid_len=a.L)
model = T5SeqAQEncoder.from_weights(dims, syn.make_weights(dims)).to("cuda:0")
proc = PrefixConstrainLogitProcessorFastSparse.from_trie(DocidTrie.from_codes(syn.make_codes(a.docs, a.L, 256), 256))
ids, mask = syn.make_queries(a.batch, S=32)
ids, mask = ids.cuda(), mask.cuda()


def run(L):
    return generate_for_constrained_prefix_beam_search(model.base_model, proc, input_ids=ids, attention_mask=mask,
                                                       max_new_tokens=L, num_beams=a.beams, num_return_sequences=a.beams,
                                                       output_scores=True, return_dict_in_generate=True,
                                                       precision=a.precision)


run(a.L)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(a.steps_only or a.L)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
