"""Pin oracle/t5_math.py (restated HF T5 block arithmetic) against the installed transformers T5Stack."""
import copy

import pytest
import torch

from oracle import t5_math
from ripor_b200 import synthetic as syn

transformers = pytest.importorskip("transformers")


def _hf_stacks(dims, w):
    from transformers.models.t5.modeling_t5 import T5Config, T5Stack
    cfg = T5Config(vocab_size=dims.vocab_size, d_model=dims.d_model, d_kv=dims.d_kv, d_ff=dims.d_ff,
                   num_layers=dims.num_layers, num_decoder_layers=dims.num_decoder_layers,
                   num_heads=dims.num_heads, feed_forward_proj="relu", layer_norm_epsilon=dims.eps,
                   relative_attention_num_buckets=dims.num_buckets, dropout_rate=0.0)
    ec = copy.deepcopy(cfg); ec.is_decoder = False; ec.use_cache = False
    dc = copy.deepcopy(cfg); dc.is_decoder = True; dc.num_layers = dims.num_decoder_layers
    enc, dec = T5Stack(ec), T5Stack(dc)
    enc.embed_tokens = torch.nn.Embedding(dims.vocab_size, dims.d_model)
    esd = {k[len("encoder."):]: v for k, v in w.items() if k.startswith("encoder.")}
    esd["embed_tokens.weight"] = w["shared.weight"]
    dsd = {k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")}
    missing, unexpected = enc.load_state_dict(esd, strict=False)
    assert not unexpected, unexpected
    missing_d, unexpected_d = dec.load_state_dict(dsd, strict=False)
    assert not unexpected_d, unexpected_d
    assert all("embed_tokens" in m for m in missing_d), missing_d
    return enc.eval(), dec.eval()


@pytest.mark.parametrize("scaleup", [False, True])
def test_encoder_and_full_prefix_decoder_match_hf(scaleup):
    dims = syn.T5Dims.tiny(docid_len=6, scaleup_output_hidden=scaleup)
    w = syn.make_weights(dims)
    enc_hf, dec_hf = _hf_stacks(dims, w)
    ids, mask = syn.make_queries(4, S=10, vocab_size=dims.vocab_size)
    g = torch.Generator().manual_seed(3)
    dec_ids = torch.randint(0, dims.decoder_vocab_size, (4, 6), generator=g)
    dec_ids[:, 0] = 0
    with torch.no_grad():
        e_hf = enc_hf(input_ids=ids, attention_mask=mask).last_hidden_state
        e_me = t5_math.encoder_forward(w, dims, ids, mask)
        valid = mask.bool()
        assert torch.allclose(e_hf[valid], e_me[valid], atol=2e-5, rtol=1e-5)
        emb = t5_math.decoder_input_embeds(w, dims, dec_ids)
        d_hf = dec_hf(inputs_embeds=emb, encoder_hidden_states=e_me, encoder_attention_mask=mask,
                      use_cache=False).last_hidden_state
        if scaleup:
            d_hf = d_hf * dims.d_model ** -0.5
        d_me = t5_math.decoder_full_prefix(w, dims, dec_ids, e_me, mask)
        assert torch.allclose(d_hf, d_me, atol=2e-5, rtol=1e-5)


def test_cached_decoder_equals_full_prefix():
    dims = syn.T5Dims.tiny(docid_len=5, shared_output_input_embeds=True)
    w = syn.make_weights(dims)
    B, nb = 2, 3
    ids, mask = syn.make_queries(B, S=9, vocab_size=dims.vocab_size)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        enc = t5_math.encoder_forward(w, dims, ids, mask)
        idx = torch.arange(B).repeat_interleave(nb)
        dec = t5_math.CachedDecoder(w, dims, enc, mask, nb)
        dec_ids = torch.zeros(B * nb, 1, dtype=torch.long)
        for t in range(dims.docid_len):
            lg = dec.step(None if t == 0 else dec_ids[:, -1])
            h = t5_math.decoder_full_prefix(w, dims, dec_ids, enc[idx], mask[idx])
            ref = t5_math.lm_logits_list(w, dims, h)[-1]
            assert torch.allclose(lg, ref, atol=1e-5, rtol=1e-5)
            perm = torch.cat([torch.randperm(nb, generator=g) + b * nb for b in range(B)])
            dec.reorder(perm)
            nxt = torch.randint(0, dims.decoder_vocab_size, (B * nb, 1), generator=g)
            dec_ids = torch.cat([dec_ids[perm], nxt], dim=1)


def test_unidirectional_buckets_for_decode_distances():
    # for distances 0..31 only buckets 0..21 occur
    rel = -torch.arange(0, 32)
    b = t5_math.relative_bucket(rel, False, 32, 128)
    assert b[:16].tolist() == list(range(16))
    assert int(b.max()) == 21
