"""ORACLE tooling: generate the golden vectors under tests/golden/ by running the reference's OWN beam
loop and mask class (oracle/ref_literal.py -> /root/reference/t5_pretrainer/tasks/generation.py:253-575,
603-677) over the reference-faithful full-prefix model adapter. Run in the build container only:

    python -m oracle.make_golden

Each fixture stores the seeds/shape needed to rebuild the synthetic inputs plus the outputs of the literal
run (sequences int64 [B*nb, L+1], sequences_scores float32 [B*nb]) and the level counts the reference
prints (evaluate.py:425-426). ``c1_t5base`` is BASELINE.json configs[0]: t5-base, 1k docs, L=8, beam=5,
batch=4; it uses the reference's learned start embedding (t5_decoder_start_token_embeds/t5-base.npy,
model input data read from /root/reference and stored in the fixture).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import beam as ob, ref_literal, t5_math  # noqa: E402
from ripor_b200 import synthetic as syn  # noqa: E402

CASES = {
    # name: (dims kwargs or "t5_base", n_docs, L, V, nb, B, S, skew, log_softmax)
    "c1_t5base": dict(base="t5_base", n_docs=1000, L=8, V=256, nb=5, B=4, S=32, skew=False, log_softmax=False),
    "tiny_plain": dict(base="tiny", n_docs=400, L=4, V=16, nb=4, B=3, S=12, skew=False, log_softmax=False),
    "tiny_logsoftmax": dict(base="tiny", n_docs=400, L=4, V=16, nb=4, B=3, S=12, skew=True, log_softmax=True),
    "tiny_shared_scaleup": dict(base="tiny", n_docs=300, L=4, V=16, nb=6, B=2, S=9, skew=False, log_softmax=False,
                                shared=True, scaleup=True),
    "tiny_few_docs": dict(base="tiny", n_docs=3, L=4, V=16, nb=5, B=2, S=10, skew=False, log_softmax=False),
}


def case_dims(c):
    kw = dict(decoder_vocab_size=c["V"], docid_len=c["L"], shared_output_input_embeds=c.get("shared", False),
              scaleup_output_hidden=c.get("scaleup", False))
    return syn.T5Dims.t5_base(**kw) if c["base"] == "t5_base" else syn.T5Dims.tiny(**kw)


def case_inputs(name, c, start_embed=None):
    dims = case_dims(c)
    w = syn.make_weights(dims, start_token_embed=start_embed)
    codes = syn.make_codes(c["n_docs"], c["L"], c["V"], skew=c["skew"], dup_frac=0.05)
    ids, mask = syn.make_queries(c["B"], S=c["S"], vocab_size=dims.vocab_size)
    return dims, w, codes, ids, mask


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    start = np.load(os.path.join(ref_literal.REFERENCE_ROOT, "t5_decoder_start_token_embeds", "t5-base.npy"))
    for name, c in CASES.items():
        t0 = time.time()
        se = start if c["base"] == "t5_base" else None
        dims, w, codes, ids, mask = case_inputs(name, c, se)
        d2s = syn.codes_to_docid_to_smtid(codes)
        lst = ob.build_list_smtid_to_nextids(d2s)
        with torch.no_grad():
            enc = t5_math.encoder_forward(w, dims, ids, mask)
            out = ref_literal.literal_beam_search(ref_literal.FullPrefixModelAdapter(w, dims),
                                                  ref_literal.literal_processor(lst, c["V"]), c["B"], c["nb"], c["L"],
                                                  apply_log_softmax_for_scores=c["log_softmax"],
                                                  encoder_states=enc, attention_mask=mask)
        s2d = ob.build_smtid_to_docids(d2s, c["L"])
        run = ob.rankdata_for_batch(list(range(c["B"])), out.sequences, out.sequences_scores, s2d, c["nb"], c["L"],
                                    c["log_softmax"])
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            case=json.dumps(c), sequences=out.sequences.numpy(), sequences_scores=out.sequences_scores.numpy(),
            encoder_states=enc.numpy().astype(np.float32), level_counts=np.asarray([len(d) for d in lst]),
            input_ids=ids.numpy(), attention_mask=mask.numpy(), codes_sha=np.frombuffer(codes.tobytes()[:64], np.uint8),
            start_token_embed=(se if se is not None else np.zeros(0, np.float32)),
            run_json=json.dumps({str(q): r for q, r in run.items()}))
        print(f"{name}: {time.time() - t0:.1f}s  top score {out.sequences_scores[0]:.6f}", flush=True)


if __name__ == "__main__":
    main()
