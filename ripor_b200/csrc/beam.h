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
};
