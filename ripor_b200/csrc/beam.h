// Beam state shared by beam.cu (the beam-step / finalize kernels) and engine.cu (the decoder reads the
// KV ancestry table the beam step maintains).
#pragma once
#include "rb_common.h"
#include "trie.h"

struct rb200_beam {
  int device = 0, max_batch = 0, nb = 0, L = 0, V = 0;
  int batch = 0;   // queries of the batch in flight
  int step = 0;    // number of steps taken since reset
  int cur = 0;     // which half of the double buffers holds the current state
  double* scores[2] = {nullptr, nullptr};       // [R] float64 running beam scores (generation.py:463,505)
  rb::TrieState* state[2] = {nullptr, nullptr};  // [R] trie state of every beam
  int32_t* hist[2] = {nullptr, nullptr};        // [R, L] tokens chosen so far (input_ids[:, 1:], generation.py:511)
  int32_t* anc[2] = {nullptr, nullptr};         // [R, L] anc[r][p] = row whose K/V at position p belongs to r's lineage
  int32_t* parent = nullptr;                    // [R] in-query beam index chosen at the last step
  int32_t* token = nullptr;                     // [R] token chosen at the last step
  int32_t* not_forced = nullptr;                // [1] rows of the current state whose trie range is not a single leaf
};

// ---- forced tail (engine.cu) --------------------------------------------------------------------------------------
// Once every beam of the batch sits on a single leaf range, the rest of its DocID is determined by the trie: the
// remaining T steps are evaluated in ONE teacher-forced pass over T*R rows (position-major: row = j*R + r).
namespace rb {
// x[(j*R + r), :] = list_decoder_embeds[t+j-1][code(r, t+j-1)] for j = 1..T-1 (block 0 was written by the last beam
// step) and hist[r][t..t+T-1] = the forced tokens
int launch_tail_prepare(rb200_beam* bm, const rb200_trie* trie, int T, const float* const* in_tabs_dev, float* x,
                        int d_model, cudaStream_t s);
// beam score += the forced tokens' logits in step order (float64, like generation.py:463); advances the state to L
int launch_tail_finish(rb200_beam* bm, const rb200_trie* trie, int T, const float* logits, int apply_log_softmax,
                       cudaStream_t s);
}  // namespace rb
