#!/usr/bin/env python
"""bench.py — queries/sec of the constrained-beam-search retrieval path (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic queries: T5 encoder -> L prefix-constrained
beam-search decoder steps over the DocID trie -> ranked DocID lists -> documents (leaf expansion), through the C ABI.
Default workload = BASELINE.json configs[1]: t5-base, 8,841,823-doc trie (32 x 256 codes), beam 10, batch 256 per
GPU. Under torchrun every rank runs its own query shard (weak scaling); the only exchange is the NCCL all_gather of
the packed ranked lists at the end of every step, inside the timed region.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python bench.py --trie zipf ...          # the same workload on the Zipf(1.1)-skewed, 1 %-duplicated trie
    python bench.py --impl reference ...     # the reference-faithful CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0). Besides the contract keys it carries `parity` (the CUDA lists against the KV-cached CPU
oracle with an INDEPENDENT mask, >= 64 queries), `zipf` (sibling measurement + parity on the skewed trie) and
`configs` (BASELINE configs 3/4/5 and the reference's shipped topk=1000 launch, each with q/s and a parity check).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec at beam=10, L=32 DocID, 8.8M-doc trie; top-10 DocID parity"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("RB200_PRECISION", "auto"),
                    help="auto = fp16x3 with an automatic tf32x3 re-run on an fp16 range overflow")
    ap.add_argument("--model", default="t5-base", choices=["t5-base", "t5-large"])
    ap.add_argument("--batch", type=int, default=256, help="queries per GPU per step")
    ap.add_argument("--beams", type=int, default=10)
    ap.add_argument("--docid-len", type=int, default=32)
    ap.add_argument("--codebook", type=int, default=256)
    ap.add_argument("--docs", type=int, default=8841823)
    ap.add_argument("--src-len", type=int, default=32)
    ap.add_argument("--trie", default="uniform", choices=["uniform", "zipf"],
                    help="zipf: per-level Zipf(1.1) codes + 1 %% duplicated rows (SURVEY 8d), deeper branching")
    ap.add_argument("--cpu-queries", type=int, default=4, help="queries in the bounded CPU-baseline sample")
    ap.add_argument("--parity-queries", type=int, default=64)
    ap.add_argument("--zipf-parity-queries", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-zipf", action="store_true", help="skip the sibling measurement on the skewed trie")
    ap.add_argument("--zipf", action="store_true", help="run the skewed-trie sibling under torchrun too")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3/4/5 + topk=1000 block")
    ap.add_argument("--configs-only", default="", help="comma list out of c3,c4,c5,top1000,top1000x8,rank4x100 (default: all)")
    return ap.parse_args()


def workload_name(a, trie=None):
    kind = trie or a.trie
    return (f"{a.model}, {a.docs:,}-doc smtid trie ({a.docid_len}x{a.codebook}, {kind}), beam={a.beams}, "
            f"batch={a.batch}/GPU, S={a.src_len}")


def dims_for(model, docid_len, codebook):
    from ripor_b200 import synthetic as syn
    kw = dict(docid_len=docid_len, decoder_vocab_size=codebook)
    return syn.T5Dims.t5_base(**kw) if model == "t5-base" else syn.T5Dims.t5_large(**kw)


def make_codes(a, kind, docs=None, L=None, V=None):
    from ripor_b200 import synthetic as syn
    return syn.make_codes(docs or a.docs, L or a.docid_len, V or a.codebook, skew=(kind == "zipf"))


# ---------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference algorithm (oracle port) and the KV-cached oracle. Nothing here touches the product
# library: weights and codes come from the seeded numpy/torch generators, the mask from oracle/range_mask.py.
# ---------------------------------------------------------------------------------------------------
def cpu_reference_search(w, dims, mask_fn, ids, mask, nb, L):
    """Reference algorithm on the host cores: encoder once, every step re-runs the decoder over
    the FULL prefix for all B*nb rows with the encoder states expanded x nb, float64 mask add, top-2nb, python beam
    scorer. mask_fn: oracle/range_mask.SortedCodesMask (the reference's dict processor needs > 60 GB at 8.8 M docs;
    the two are pinned equal on random tries in tests/test_host_round2.py)."""
    import torch
    from oracle import beam as ob, t5_math
    B = ids.shape[0]
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        idx = torch.arange(B).repeat_interleave(nb)
        enc_r, mask_r = enc[idx], mask[idx]

        def full(dec_ids, bi):
            h = t5_math.decoder_full_prefix(w, dims, dec_ids, enc_r, mask_r)
            return t5_math.lm_logits_list(w, dims, h)[-1]

        return ob.beam_search_oracle(full, mask_fn, B, nb, L)


def cpu_cached_search(w, dims, mask_fn, ids, mask, nb, L):
    """KV-cached fp32 CPU oracle: same results as the reference algorithm with 1/16 of its decoder work."""
    import torch
    from oracle import beam as ob, t5_math
    B = ids.shape[0]
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        dec = t5_math.CachedDecoder(w, dims, enc, mask, nb)

        def cstep(dec_ids, bi):
            if bi is not None:
                dec.reorder(bi)
            return dec.step(None if dec_ids.shape[1] == 1 else dec_ids[:, -1])
        return ob.beam_search_oracle(cstep, mask_fn, B, nb, L)


def run_reference(a):
    """--impl reference: the reference's own CPU path (oracle port; the reference cannot be imported under the
    installed transformers) on a bounded sample of the same workload, all host threads. Imports
    nothing from the product's native library."""
    import torch
    from oracle.range_mask import SortedCodesMask
    from ripor_b200 import synthetic as syn
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dims = dims_for(a.model, a.docid_len, a.codebook)
    w = syn.make_weights(dims)
    mask_fn = SortedCodesMask(make_codes(a, a.trie), a.codebook)
    q = a.cpu_queries
    ids, mask = syn.make_queries(q, S=a.src_len)
    for _ in range(a.warmup):
        cpu_reference_search(w, dims, mask_fn, ids, mask, a.beams, a.docid_len)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_reference_search(w, dims, mask_fn, ids, mask, a.beams, a.docid_len)
    dt = time.perf_counter() - t0
    val = q * a.steps / dt
    sample = f"{q} queries/step of the same workload (full-prefix decoder, R={q * a.beams} rows, {a.docid_len} steps)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 logits / f64 beam scores", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def check_parity(out, B, nb, L, w, dims, mask_fn, ids, mask, pq):
    """The first pq queries of the CUDA result against the KV-cached CPU oracle (mask from oracle/range_mask.py, which
    shares nothing with the product's trie). DocID lists must be bit-exact, scores within 1e-3 (north star)."""
    import torch
    t0 = time.perf_counter()
    ref_seq, ref_sc = cpu_cached_search(w, dims, mask_fn, ids[:pq], mask[:pq], nb, L)
    dt = time.perf_counter() - t0
    got_seq = out.sequences.view(B, nb, L + 1)[:pq].reshape(pq * nb, L + 1).cpu()
    got_sc = out.sequences_scores.view(B, nb)[:pq].reshape(-1).cpu()
    exact = int((got_seq.view(pq, -1) == ref_seq.view(pq, -1)).all(dim=1).sum())
    return {"queries_checked": pq, "docid_lists_exact": exact,
            "max_abs_score_diff": float((got_sc - ref_sc).abs().max()), "score_tolerance": 1e-3,
            "oracle": "KV-cached fp32 CPU oracle (oracle/t5_math.py, oracle/beam.py)",
            "mask": "oracle/range_mask.py SortedCodesMask (binary search over the sorted codes; independent of the "
                    "product's trie)", "oracle_seconds": round(dt, 1),
            "oracle_qps": round(pq / dt, 3)}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from oracle.range_mask import SortedCodesMask
    from ripor_b200 import _lib, synthetic as syn
    from ripor_b200 import evaluate as ev
    from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse, generate_for_constrained_prefix_beam_search
    from ripor_b200.modeling import T5SeqAQEncoder
    from ripor_b200.trie import DocidTrie

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _lib.lib()
    cores = os.cpu_count() or 1
    WIDTH = ev.LEAF_EXPAND_WIDTH

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(model, proc, B, nb, L, S, steps, warmup, precision, seed_off=0, with_clocks=False):
        """warm-up + timed region of K steps with inputs resident in HBM; returns (result dict, last output, ids, mask).
        A step = search + device-side leaf expansion (+ the NCCL all_gather of the packed ranked lists when world > 1)."""
        ids, mask = syn.make_queries(B, S=S, seed=syn.QUERY_SEED + rank + seed_off)
        ids_d, mask_d = ids.to(dev), mask.to(dev)
        qids = torch.arange(B, device=dev) * world + rank
        trie = proc.trie

        def step():
            out = generate_for_constrained_prefix_beam_search(
                model.base_model, proc, input_ids=ids_d, attention_mask=mask_d, max_new_tokens=L, num_beams=nb,
                num_return_sequences=nb, output_scores=True, return_dict_in_generate=True, precision=precision)
            docs, counts = trie.expand_ranges(out.leaf_ranges, WIDTH)
            if world > 1:
                ev.gather_ranked_lists(qids, docs.view(B, nb * WIDTH), out.sequences_scores.view(B, nb))
            return out
        out = None
        for _ in range(max(warmup, 1)):
            out = step()
        sampler = ClockSampler(local) if with_clocks else None
        sync_all()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = step()
        e1.record()
        sync_all()
        clocks = sampler.stop() if sampler else None
        ms = max_over_ranks(e0.elapsed_time(e1))
        res = {"value": world * B * steps / (ms / 1e3), "ms_per_step": ms / steps, "precision": out.precision,
               "gpu_launches_per_step": out.gpu_launches + 1,
               "frozen_at_step": out.frozen_at_step, "forced_tail_rows": out.forced_tail_rows,
               "first_freeze_step": out.forced_tail_from}
        if clocks:
            res["clocks"] = clocks
        return res, out, ids, mask

    dims = dims_for(a.model, a.docid_len, a.codebook)
    w = syn.make_weights(dims)
    t0 = time.time()
    codes = make_codes(a, a.trie)
    trie = DocidTrie.from_codes(codes, a.codebook)
    trie_build_s = time.time() - t0
    proc = PrefixConstrainLogitProcessorFastSparse.from_trie(trie)
    model = T5SeqAQEncoder.from_weights(dims, w).to(dev)
    B, nb, L, S = a.batch, a.beams, a.docid_len, a.src_len

    main, out, ids, mask = measure(model, proc, B, nb, L, S, a.steps, a.warmup, a.precision, with_clocks=True)
    prec = main["precision"]                  # what 'auto' resolved to
    # ---- e2e: same metric through the public API with HOST buffers (H2D of the token batch, search, leaf expansion,
    #      D2H of the ranked documents + scores inside the timed region), plus the gather when world > 1 ----
    ids_h, mask_h = ids.pin_memory(), mask.pin_memory()
    qids = torch.arange(B, device=dev) * world + rank

    def e2e_step():
        (docs, counts, scores, leaf), o = ev.retrieve_batch(model.base_model, proc, ids_h, mask_h, L, nb, precision=a.precision)
        if world > 1:
            ev.gather_ranked_lists(qids, docs.view(B, nb * WIDTH).to(dev), scores.to(dev))
            torch.cuda.synchronize()
        return docs, counts, scores
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        docs_h, counts_h, scores_h = e2e_step()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sync_all()
    e2e = {"value": world * B * a.steps / e2e_s, "unit": "queries/s",
           "h2d_bytes_per_step": 2 * B * S * 8, "d2h_bytes_per_step": B * nb * (WIDTH * 8 + 4 + 4 + 8),
           "what": "ripor_b200.evaluate.retrieve_batch: pinned token batch -> search -> device leaf expansion -> "
                   "document rows + counts + scores + leaf ranges on the host"}
    # ---- roofline of the dominant kernel family (the decoder/encoder GEMMs), measured live with events ----
    eng = model.base_model.get_engine(B, nb, S, prec)
    ids_d, mask_d = ids.to(dev), mask.to(dev)

    def one_search():
        return generate_for_constrained_prefix_beam_search(
            model.base_model, proc, input_ids=ids_d, attention_mask=mask_d, max_new_tokens=L, num_beams=nb,
            num_return_sequences=nb, output_scores=True, return_dict_in_generate=True, precision=prec)
    _lib.check(lib.rb200_engine_set_profiling(eng.h, 1))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    one_search()
    p1.record()
    torch.cuda.synchronize()
    gm, gf, gl = C.c_double(), C.c_double(), C.c_int64()
    _lib.check(lib.rb200_engine_get_profile(eng.h, C.byref(gm), C.byref(gf), C.byref(gl)))
    gb = C.c_double()
    _lib.check(lib.rb200_engine_get_profile_bytes(eng.h, C.byref(gb)))
    _lib.check(lib.rb200_engine_set_profiling(eng.h, 0))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = gf.value / (gm.value / 1e3) / 1e12 if gm.value > 0 else 0.0
    mma_mult = 3 if prec in ("tf32x3", "bf16x3", "fp16x3") else 1
    traffic, traffic_note = None, None
    try:   # dram bytes per launch of the GEMM kernel from the committed ncu --set full capture of THIS kernel variant
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        ent = tj.get(prec, {})
        traffic, traffic_note = ent.get("dram_bytes_per_launch"), ent.get("note")
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note,
                "algorithmic_bytes_per_launch": gb.value / max(gl.value, 1),
                "hbm_gbs_of_gemms": gb.value / (gm.value / 1e3) / 1e9 if gm.value > 0 else None,
                "kernel": f"gemm_sm100_2cta_kernel[{prec}]" if prec != "fp32" else "gemm_simt_kernel",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)",
                "launches_per_step": int(gl.value), "avg_launch_us": gm.value * 1e3 / max(gl.value, 1),
                "gemm_share_of_step": gm.value / p0.elapsed_time(p1),
                "issued_mma_tflops": achieved * mma_mult,
                "note": "achieved = algorithmic 2*M*N*K of every GEMM launch of one step / their summed event time; "
                        f"{prec} issues {mma_mult} tensor-core MMA(s) per algorithmic product, so frac <= 1/{mma_mult} "
                        "by construction; frac_of_split_peak = issued / peak. The GEMM launches include the T5 "
                        "layer norms (NormFold, default): the residual add, the sum of squares and the next GEMM's "
                        "normalised operand planes run in the o / co / wo epilogues instead of 3 RMSNorm launches per "
                        "layer, which lengthens those launches (lower GEMM-only frac) and shortens the step (higher "
                        "step_level_frac); RB200_FOLD=0 restores the separate launches",
                "frac_of_split_peak": achieved * mma_mult / peak,
                # whole step, per GPU: SURVEY 8d's 70.5 GFLOP per query x queries/s / peak
                "step_level_frac": 70.5e9 * B / (main["ms_per_step"] / 1e3) / 1e12 / peak
                if (a.model, nb, L, a.codebook) == ("t5-base", 10, 32, 256) else None}
    # ---- secondary: achieved HBM GB/s of the trie / top-k beam-step kernel (BASELINE north_star) on a batch whose
    # logits (R x V fp32) do not fit L2; algorithmic bytes per launch = logits + beam/trie state in + state out ----
    trie_topk = None
    if rank == 0:
        try:
            Rq = 1 << 14                                             # queries -> 163,840 rows, 168 MB of logits at V=256
            hb = C.c_void_p()
            _lib.check(lib.rb200_beam_create(local, Rq, nb, L, a.codebook, C.byref(hb)))
            big = torch.randn((Rq * nb, a.codebook), device=dev)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nlaunch = min(4, L - 1)
            reps, total_ms = 4, 0.0                                  # repetition 0 is the warm-up (cold trie tables)
            for rep in range(reps):
                _lib.check(lib.rb200_beam_reset(hb, trie.handle, Rq, _lib.stream_ptr()))
                _lib.check(lib.rb200_beam_step(hb, trie.handle, big.data_ptr(), 1, 0, None, None, 0, _lib.stream_ptr()))
                torch.cuda.synchronize()
                b0.record()
                for _ in range(nlaunch):                             # steps 1..4: 168 MB of logits per launch (> L2)
                    _lib.check(lib.rb200_beam_step(hb, trie.handle, big.data_ptr(), nb, 0, None, None, 0,
                                                   _lib.stream_ptr()))
                b1.record()
                torch.cuda.synchronize()
                if rep > 0:
                    total_ms += b0.elapsed_time(b1)
            lib.rb200_beam_free(hb)
            us = total_ms * 1e3 / (nlaunch * (reps - 1))
            per_row = a.codebook * 4 + (8 + 16) + (8 + 16 + 4 + 4) + 2 * L * 4 * 2   # logits, score+state in, out, hist+anc in/out
            nbytes = Rq * nb * per_row
            survey_bytes = Rq * nb * (a.codebook * 4 + a.codebook // 8 + 8 + 4 + 8 + 20)   # ~1.07 KB per beam-step
            hbm = peaks.get("hbm_gbs", 6550.0)
            trie_topk = {"kernel": "beam_step_warp_kernel" if nb <= 16 else "beam_step_kernel", "rows": Rq * nb,
                         "avg_launch_us": us,
                         "algorithmic_bytes_per_launch": nbytes, "achieved_gbs": nbytes / us / 1e3, "peak_gbs": hbm,
                         "frac": nbytes / us / 1e3 / hbm, "bound": "hbm",
                         "survey_8d_bytes_per_launch": survey_bytes, "survey_8d_frac": survey_bytes / us / 1e3 / hbm,
                         "note": "one warp per query: fp32 pre-filter against proven per-beam thresholds, float64 "
                                 "ranking of the survivors, trie child lookup (dependent random reads into 338 MB of "
                                 "trie tables), history / ancestry reorder; issue-bound (ncu: 64 % issue slots busy), "
                                 "< 0.5 % of a search; 3 timed repetitions of steps 1..4 after one warm-up repetition"}
            del big
        except Exception as exc:                                     # the secondary figure must never break the bench line
            trie_topk = {"error": str(exc)[:200]}
    result = {
        "metric": METRIC, "value": main["value"], "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (fp32-grade split) + f64 beam scores", "bf16x3": "bf16x3 split",
                  "fp16x3": "fp16x3 (fp32-grade split, 11-bit planes) + f64 beam scores",
                  "tf32": "tf32", "bf16": "bf16"}[prec],
        "data": "synthetic",
        "config": {"workload": workload_name(a), "trie": a.trie, "precision": prec, "precision_requested": a.precision,
                   "global_batch": world * B, "frozen_at_step": main["frozen_at_step"],
                   "forced_tail_rows": main["forced_tail_rows"],
                   "l2": "per-step working set (fp32 KV cache + weights, >6 GB) exceeds the 126 MB L2",
                   "trie_build_s": round(trie_build_s, 2), "parallelism": f"query-sharded x{world}",
                   "collective": (f"all_gather_into_tensor of packed int64 [B, 1 + nb*{WIDTH} + nb] per rank per step "
                                  f"= {B * (1 + nb * WIDTH + nb) * 8} B/rank (document rows + scores + query ids)")
                   if world > 1 else "none (1 GPU)", "host_cores": cores},
        "clocks": main.get("clocks"), "e2e": e2e, "gpu_launches": int(main["gpu_launches_per_step"] * a.steps),
        "roofline": roofline, "trie_topk": trie_topk}
    # ---- parity at the benchmarked size + CPU baselines (rank 0) ---------------------------------------
    mask_fn = None
    if rank == 0:
        torch.set_num_threads(cores)
        pq = min(a.parity_queries, B)
        if pq > 0:
            mask_fn = SortedCodesMask(codes, a.codebook)
            result["parity"] = check_parity(out, B, nb, L, w, dims, mask_fn, ids, mask, pq)
            # the documents the e2e call returned are the ones under the oracle's DocIDs (first query, json order)
            ref_rows = [sorted(int(r) for r in (codes[:, :L] == out.sequences[j, 1:].cpu().numpy()).all(1).nonzero()[0])
                        for j in range(nb)]
            got_rows = [[int(r) for r in docs_h[0, j, : int(counts_h[0, j])]] for j in range(nb)]
            result["parity"]["leaf_expansion_checked"] = bool(ref_rows == got_rows)
        if world == 1 and not a.no_cpu_baseline:
            q = a.cpu_queries
            mask_fn = mask_fn or SortedCodesMask(codes, a.codebook)
            t0 = time.perf_counter()
            cpu_reference_search(w, dims, mask_fn, ids[:q], mask[:q], nb, L)
            dt = time.perf_counter() - t0
            result["cpu_baseline"] = {"value": q / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                                      "sample": f"{q} queries of the same workload, reference algorithm "
                                                f"(full-prefix decoder every step, R={q * nb} rows), {dt:.1f} s",
                                      "kv_cached_oracle_qps": result.get("parity", {}).get("oracle_qps"),
                                      "kv_cached_note": "best honest CPU: the KV-cached fp32 oracle timed in the parity "
                                                        f"leg on {pq} queries"}
    del mask_fn
    # ---- sibling: the same workload on the Zipf-skewed trie (beams reach single leaves at different steps) ----
    if not a.no_zipf and a.trie == "uniform" and (world == 1 or a.zipf):   # (scaling runs: only when asked, --zipf)
        del codes, trie, proc
        zcodes = make_codes(a, "zipf")
        ztrie = DocidTrie.from_codes(zcodes, a.codebook)
        zproc = PrefixConstrainLogitProcessorFastSparse.from_trie(ztrie)
        zres, zout, zids, zmask = measure(model, zproc, B, nb, L, S, a.steps, max(a.warmup, 1), a.precision)
        if rank == 0:
            zres.update({"trie": "zipf", "workload": workload_name(a, "zipf"), "unit": "queries/s"})
            zq = min(a.zipf_parity_queries, B)
            if zq > 0:
                zres["parity"] = check_parity(zout, B, nb, L, w, dims, SortedCodesMask(zcodes, a.codebook), zids, zmask, zq)
            result["zipf"] = zres
        del zcodes, ztrie, zproc, zout
    # ---- the other BASELINE configs + the reference's shipped launch, bounded (2 steps each; 4 for configs[4], 8 for the millisecond-long launches) ----
    if not a.no_configs and a.trie == "uniform" and (a.model, nb, L, a.codebook) == ("t5-base", 10, 32, 256):
        want = [c for c in a.configs_only.split(",") if c] or \
            (["c3"] if world > 1 else ["top1000", "rank4x100", "top1000x8", "c3", "c4", "c5"])
        result["configs"] = run_configs(a, want, model, w, dims, measure, rank, world)
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_configs(a, want, base_model, base_w, base_dims, measure, rank, world):
    """q/s (+ a small parity check on rank 0) for BASELINE configs[2..4] and the reference's shipped evaluation launch
    (full_scripts/full_evaluate_t5seq_aq_encoder.sh:191-199: batch 1, topk 1000). 2 timed steps each (8 for the millisecond-long launches)."""
    import torch
    from oracle.range_mask import SortedCodesMask
    from ripor_b200 import synthetic as syn
    from ripor_b200.generation import PrefixConstrainLogitProcessorFastSparse
    from ripor_b200.modeling import T5SeqAQEncoder
    from ripor_b200.trie import DocidTrie
    dev = base_model.base_model.device
    table = {
        "c3": dict(name="configs[2]: t5-base, 8.8M-doc trie, beam=100, batch=128/GPU (1024 over 8 GPUs)",
                   model="t5-base", L=32, V=256, nb=100, B=128, pq=4),
        "c4": dict(name="configs[3]: t5-large, 8.8M-doc trie (32x256), beam=10, batch=512", model="t5-large", L=32,
                   V=256, nb=10, B=512, pq=4),
        "c5": dict(name="configs[4]: t5-base, 16x1024 codebooks, beam=10, batch=256/GPU", model="t5-base", L=16, V=1024,
                   nb=10, B=256, pq=4),
        "top1000": dict(name="reference's shipped eval launch: t5-base, batch_size=1, topk=1000, L=32", model="t5-base",
                        L=32, V=256, nb=1000, B=1, pq=0),
        # the same launch with --batch_size 8 (what a user of this engine would pass: one SM-filling batch)
        "top1000x8": dict(name="shipped eval launch at --batch_size 8: t5-base, topk=1000, L=32", model="t5-base",
                          L=32, V=256, nb=1000, B=8, pq=0),
        # full_scripts/full_evaluate_t5seq_aq_encoder.sh:128-139: t5seq_aq_get_qid_to_smtid_rankdata, DocID prefixes
        "rank4x100": dict(name="reference's shipped training-data launch: t5-base, batch_size=4, topk=100, "
                               "max_new_token=16 (prefix search over the 32-code trie)", model="t5-base", L=16, trie_L=32,
                          V=256, nb=100, B=4, pq=2),
    }
    out = {}
    tries = {}
    for key in want:
        c = table[key]
        try:
            tk = (c.get("trie_L", c["L"]), c["V"])
            if tk not in tries:
                tries.clear()                      # one 8.8M-doc trie in host memory at a time
                codes = syn.make_codes(a.docs, tk[0], c["V"])
                tries[tk] = (codes, PrefixConstrainLogitProcessorFastSparse.from_trie(DocidTrie.from_codes(codes, c["V"])))
            codes, proc = tries[tk]
            if (c["model"], tk[0], c["V"]) == ("t5-base", base_dims.docid_len, base_dims.decoder_vocab_size):
                model, w, dims = base_model, base_w, base_dims
            else:
                for p in list(base_model.base_model._engines):       # make room: free the default model's engines
                    base_model.base_model.drop_engine(p)
                dims = dims_for(c["model"], tk[0], c["V"])
                w = syn.make_weights(dims)
                model = T5SeqAQEncoder.from_weights(dims, w).to(dev)
            # the small launches are milliseconds long: more warm-up and steps, so that the power state the previous
            # (heavy) config left behind does not colour them
            # (configs[4] steps in ~32 ms: 4 + 3 steps keep one hiccup from halving its figure)
            st, wu = (8, 8) if c["B"] * c["nb"] <= 1000 else ((4, 3) if key == "c5" else (2, 1))
            res, o, ids, mask = measure(model, proc, c["B"], c["nb"], c["L"], a.src_len, st, wu, a.precision, seed_off=1000)
            res.update({"workload": c["name"], "unit": "queries/s", "steps": st, "warmup": wu})
            if rank == 0 and c["pq"] > 0:
                res["parity"] = check_parity(o, c["B"], c["nb"], c["L"], w, dims, SortedCodesMask(codes, c["V"]), ids, mask,
                                             c["pq"])
            out[key] = res
            if model is not base_model:
                for p in list(model.base_model._engines):
                    model.base_model.drop_engine(p)
                del model, w
        except Exception as exc:                    # a config that does not fit must not take the headline line down
            out[key] = {"workload": c["name"], "error": str(exc)[:300]}
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
