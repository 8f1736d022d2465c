"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck): tiny-model searches in the two
parity precisions (step loop, per-query forced tail incl. the ragged pass, radix-select beam kernel, leaf expansion,
teacher-forced forward) plus the GEMM edge shapes M = 1 and M = 77. Results are checked against the CPU oracle so a
sanitizer-clean run is also a correct run.

    compute-sanitizer --tool memcheck  --print-limit 20 python tools/sanitizer_run.py
    compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_run.py --light
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ripor_b200 import _lib, synthetic as syn  # noqa: E402
from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search  # noqa: E402
from ripor_b200.modeling import T5SeqAQEncoder  # noqa: E402
from ripor_b200.trie import DocidTrie  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--light", action="store_true", help="fewer cases (racecheck is ~100x slower than native)")
a = ap.parse_args()
dev = "cuda:0"


def search(dims, w, codes, nb, L, B, S, precision, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    ids, mask = syn.make_queries(B, S=S, vocab_size=dims.vocab_size)
    ref_seq, ref_sc, _, _ = helpers.oracle_cached_search(w, dims, codes, ids, mask, nb, L)
    model = T5SeqAQEncoder.from_weights(dims, w).to(dev)
    trie = DocidTrie.from_codes(codes, dims.decoder_vocab_size)
    proc = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
    out = generate_for_constrained_prefix_beam_search(
        model.base_model, proc, input_ids=ids.to(dev), attention_mask=mask.to(dev), max_new_tokens=L, num_beams=nb,
        num_return_sequences=nb, output_scores=True, return_dict_in_generate=True, precision=precision)
    docs, counts = trie.expand_ranges(out.leaf_ranges, 4)
    torch.cuda.synchronize()
    bad = helpers.compare_ranked(out.sequences, out.sequences_scores, ref_seq, ref_sc, nb, atol=1e-3)
    assert bad == 0, (precision, nb, bad)
    for k in (env or {}):
        os.environ.pop(k, None)
    print(f"ok search {precision} nb={nb} L={L} B={B} frozen_at={out.frozen_at_step} launches={out.gpu_launches}", flush=True)
    return model, ids, mask


L = 8
dims = syn.T5Dims.tiny(docid_len=L)
w = syn.make_weights(dims)
V = dims.decoder_vocab_size
codes = syn.make_codes(3000, L, V, skew=True)
precisions = ("fp16x3",) if a.light else ("tf32x3", "fp16x3")
for prec in precisions:
    model, ids, mask = search(dims, w, codes, 5, L, 4, 12, prec)                          # warp beam kernel, ragged tail
    if not a.light:
        search(dims, w, codes, 20, L, 2, 40, prec)                                        # CTA beam kernel, S > 32
    search(dims, w, codes, 20, L, 2, 12, prec, env={"RB200_BEAM": "select"})              # radix-select beam kernel
    if not a.light:
        search(dims, w, codes, 5, L, 3, 12, prec, env={"RB200_TAIL": "0"})                # step loop to the end
    tok = torch.randint(0, V, (ids.shape[0] * 2, L), dtype=torch.int32)
    out = model.base_model._forced(ids, mask, tok, 2, want_logits=True, want_hidden=True, want_scores=True, precision=prec)
    torch.cuda.synchronize()
    assert torch.isfinite(out["scores"]).all()
    print(f"ok forward {prec}", flush=True)
lib = _lib.lib()
# the warp beam-step kernel in its compile-time codebook widths (V = 256, 1024) and the generic one (V = 64), alone,
# against the all-float64 CTA kernel: random logits, a few steps each
import ctypes as C  # noqa: E402


def beam_steps(trie, B, nb, Lb, Vb, logits):
    h = C.c_void_p()
    _lib.check(lib.rb200_beam_create(0, B, nb, Lb, Vb, C.byref(h)))
    _lib.check(lib.rb200_beam_reset(h, trie.handle, B, _lib.stream_ptr()))
    for t in range(Lb):
        lg = logits[t][: B] if t == 0 else logits[t]
        _lib.check(lib.rb200_beam_step(h, trie.handle, lg.data_ptr(), 1 if t == 0 else nb, 0, None, None, 0,
                                       _lib.stream_ptr()))
    seqs = torch.empty((B * nb, Lb + 1), dtype=torch.int64, device=dev)
    sc = torch.empty((B * nb,), dtype=torch.float32, device=dev)
    leaf = torch.empty((B * nb, 2), dtype=torch.int32, device=dev)
    _lib.check(lib.rb200_beam_finalize(h, trie.handle, nb, 1.0, seqs.data_ptr(), sc.data_ptr(), leaf.data_ptr(),
                                       _lib.stream_ptr()))
    torch.cuda.synchronize()
    lib.rb200_beam_free(h)
    return seqs.cpu(), sc.cpu(), leaf.cpu()


for Vb, nb in ((256, 10), (1024, 7), (64, 16)):
    Lb, B = 4, 5
    tr = DocidTrie.from_codes(syn.make_codes(4000, Lb, Vb, seed=3, dup_frac=0.05), Vb).upload(0)
    g = torch.Generator().manual_seed(Vb)
    logits = [(torch.randn(B * nb, Vb, generator=g) * 3).to(dev) for _ in range(Lb)]
    got = beam_steps(tr, B, nb, Lb, Vb, logits)
    os.environ["RB200_BEAM"] = "cta"
    want = beam_steps(tr, B, nb, Lb, Vb, logits)
    os.environ.pop("RB200_BEAM")
    assert all(torch.equal(x, y) for x, y in zip(got, want)), (Vb, nb)
    print(f"ok beam warp kernel V={Vb} nb={nb}", flush=True)
for prec in precisions:
    for M, N, K in ((1, 64, 64), (77, 136, 72)):
        A, W = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
        Cc = torch.zeros(M, N, device=dev)
        _lib.check(lib.rb200_gemm(_lib.PRECISIONS[prec], A.data_ptr(), W.data_ptr(), Cc.data_ptr(), M, N, K, 0, 0, None))
        torch.cuda.synchronize()
        err = (Cc - A.double().matmul(W.double().t()).float()).abs().max().item()
        assert err < 1e-3, (prec, M, err)
        print(f"ok gemm {prec} M={M} err={err:.2e}", flush=True)
print("SANITIZER_WORKLOAD_OK")
