// On-disk inputs of the retrieval path (host code): docid_to_smtid.json and bit-packed residual-quantiser codes.
//
//   * docid_to_smtid.json = {"<docid>": [-1, c1, .., cL], ...} as written by the reference's
//     aq_preprocess/create_customized_smtid_file.py:47-58 / create_smtid_file.py (ujson.dump) and read back with
//     ujson.load in t5_pretrainer/evaluate.py:400-401. The reference then walks the Python dict of 8.8 M lists; here a
//     streaming parser turns the file straight into a codes matrix [N, L] plus a docid string table in file order
//     (evaluate.py:439-446 depends on that order).
//   * faiss packs the M codes of a vector, `bits` bits each, LSB first into ceil(M*bits/8) bytes; the reference unpacks
//     them with faiss.BitstringReader (create_customized_smtid_file.py:38-45) when bits != 8. faiss is a third-party
//     dependency that is not vendored (requirements: faiss-gpu, unpinned); rb200_unpack_codes restates
//     BitstringReader::read (faiss/utils/hamming.h): bit i of the stream is bit (i & 7) of byte (i >> 3).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "rb_common.h"
#include "trie.h"

struct rb200_docid_table {
  int64_t n_docs = 0;
  int32_t L = 0;              // codes per document (without the leading -1)
  int32_t max_code = -1;
  std::vector<int32_t> codes;      // [n_docs, L]
  std::vector<char> keys;          // concatenated docid strings (no terminators)
  std::vector<int64_t> key_off;    // [n_docs + 1]
};

namespace {

struct Cursor {
  const char* p;
  const char* end;
  int64_t line_hint = 0;
  bool ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    return p < end;
  }
};

int parse_failure(const Cursor& c, const char* base, const char* what) {
  return rb::fail(RB200_ERR_IO, "docid_to_smtid.json: %s at byte %lld", what, (long long)(c.p - base));
}

int parse_docid_json(const char* base, size_t size, int max_len, rb200_docid_table* t) {
  Cursor c{base, base + size};
  if (!c.ws() || *c.p != '{') return parse_failure(c, base, "expected '{'");
  ++c.p;
  // a first guess of the table sizes from the file size (MS MARCO: ~140 bytes per entry at L = 32)
  t->key_off.reserve(size / 64 + 16);
  t->key_off.push_back(0);
  bool first = true;
  for (;;) {
    if (!c.ws()) return parse_failure(c, base, "unexpected end of file");
    if (*c.p == '}') break;
    if (!first) {
      if (*c.p != ',') return parse_failure(c, base, "expected ',' between entries");
      ++c.p;
      if (!c.ws()) return parse_failure(c, base, "unexpected end of file");
    }
    first = false;
    if (*c.p != '"') return parse_failure(c, base, "expected a docid string");
    ++c.p;
    const char* k0 = c.p;
    bool escaped = false;
    while (c.p < c.end && *c.p != '"') {
      if (*c.p == '\\') { escaped = true; ++c.p; }
      ++c.p;
    }
    if (c.p >= c.end) return parse_failure(c, base, "unterminated docid string");
    if (!escaped) {
      t->keys.insert(t->keys.end(), k0, c.p);
    } else {                                       // rare: undo the JSON escapes that can occur in an id
      for (const char* q = k0; q < c.p; ++q) {
        if (*q == '\\' && q + 1 < c.p) {
          ++q;
          switch (*q) {
            case 'n': t->keys.push_back('\n'); break;
            case 't': t->keys.push_back('\t'); break;
            case 'r': t->keys.push_back('\r'); break;
            case 'b': t->keys.push_back('\b'); break;
            case 'f': t->keys.push_back('\f'); break;
            case 'u': return parse_failure(c, base, "\\u escapes in docids are not supported");
            default: t->keys.push_back(*q);
          }
        } else {
          t->keys.push_back(*q);
        }
      }
    }
    t->key_off.push_back((int64_t)t->keys.size());
    ++c.p;
    if (!c.ws() || *c.p != ':') return parse_failure(c, base, "expected ':'");
    ++c.p;
    if (!c.ws() || *c.p != '[') return parse_failure(c, base, "expected '['");
    ++c.p;
    int n = 0;          // values seen in this list, including the leading -1
    for (;;) {
      if (!c.ws()) return parse_failure(c, base, "unexpected end of file");
      if (*c.p == ']') { ++c.p; break; }
      if (n > 0) {
        if (*c.p != ',') return parse_failure(c, base, "expected ',' in a code list");
        ++c.p;
        if (!c.ws()) return parse_failure(c, base, "unexpected end of file");
      }
      bool neg = false;
      if (*c.p == '-') { neg = true; ++c.p; }
      if (c.p >= c.end || *c.p < '0' || *c.p > '9') return parse_failure(c, base, "expected an integer");
      int64_t v = 0;
      while (c.p < c.end && *c.p >= '0' && *c.p <= '9') {
        v = v * 10 + (*c.p - '0');
        if (v > 0x7fffffff) return parse_failure(c, base, "code out of range");
        ++c.p;
      }
      if (neg) v = -v;
      if (n == 0) {
        if (v != -1) return parse_failure(c, base, "a code list must start with -1 (evaluate.py:441)");
      } else if (max_len <= 0 || n <= max_len) {
        if (v < 0) return parse_failure(c, base, "negative code");
        t->codes.push_back((int32_t)v);
        if (v > t->max_code) t->max_code = (int32_t)v;
      }
      ++n;
    }
    const int len = (max_len > 0 && n - 1 > max_len) ? max_len : n - 1;
    if (t->n_docs == 0) {
      if (len < 1) return parse_failure(c, base, "empty code list");
      t->L = len;
      const size_t guess = size / (size_t)(c.p - base) + 16;      // entries, judged by the first one
      t->codes.reserve(guess * (size_t)len + 1024);
    } else if (len != t->L) {
      return parse_failure(c, base, "code lists of different lengths");
    }
    ++t->n_docs;
  }
  if (t->n_docs == 0) return rb::fail(RB200_ERR_IO, "docid_to_smtid.json holds no documents");
  return 0;
}

}  // namespace

extern "C" {

int rb200_docid_json_open(const char* path, int max_len, rb200_docid_table** out) {
  RB_REQUIRE(path && out, "null argument");
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return rb::fail(RB200_ERR_IO, "cannot open %s", path);
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) {
    close(fd);
    return rb::fail(RB200_ERR_IO, "cannot stat %s (or it is empty)", path);
  }
  void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return rb::fail(RB200_ERR_IO, "cannot map %s", path);
  madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
  rb200_docid_table* t = new (std::nothrow) rb200_docid_table();
  if (!t) { munmap(map, (size_t)st.st_size); return rb::fail(RB200_ERR_NOMEM, "out of memory"); }
  int status = 0;
  try {
    status = parse_docid_json(static_cast<const char*>(map), (size_t)st.st_size, max_len, t);
  } catch (const std::bad_alloc&) {
    status = rb::fail(RB200_ERR_NOMEM, "out of memory while reading %s", path);
  }
  munmap(map, (size_t)st.st_size);
  if (status != 0) { delete t; return status; }
  *out = t;
  return 0;
}

int rb200_docid_json_free(rb200_docid_table* t) {
  delete t;
  return 0;
}

int rb200_docid_json_info(const rb200_docid_table* t, int64_t* n_docs, int32_t* L, int32_t* max_code,
                          int64_t* key_bytes) {
  RB_REQUIRE(t, "null argument");
  if (n_docs) *n_docs = t->n_docs;
  if (L) *L = t->L;
  if (max_code) *max_code = t->max_code;
  if (key_bytes) *key_bytes = (int64_t)t->keys.size();
  return 0;
}

int rb200_docid_json_codes(const rb200_docid_table* t, const int32_t** codes) {
  RB_REQUIRE(t && codes, "null argument");
  *codes = t->codes.data();
  return 0;
}

int rb200_docid_json_keys(const rb200_docid_table* t, const char** key_bytes, const int64_t** key_offsets) {
  RB_REQUIRE(t && key_bytes && key_offsets, "null argument");
  *key_bytes = t->keys.data();
  *key_offsets = t->key_off.data();
  return 0;
}

int rb200_trie_build_from_table(const rb200_docid_table* t, int V, int n_threads, rb200_trie** out) {
  RB_REQUIRE(t && out, "null argument");
  RB_REQUIRE(t->max_code < V, "docid_to_smtid.json holds code %d but the model's decoder_vocab_size is %d", t->max_code, V);
  const int64_t n = t->n_docs * t->L;
  if (V <= 256) {
    std::vector<uint8_t> c8((size_t)n);
    for (int64_t i = 0; i < n; ++i) c8[i] = (uint8_t)t->codes[i];
    return rb200_trie_build(c8.data(), 1, t->n_docs, t->L, V, n_threads, out);
  }
  RB_REQUIRE(V <= 65536, "V=%d does not fit 2-byte codes", V);
  std::vector<uint16_t> c16((size_t)n);
  for (int64_t i = 0; i < n; ++i) c16[i] = (uint16_t)t->codes[i];
  return rb200_trie_build(c16.data(), 2, t->n_docs, t->L, V, n_threads, out);
}

int rb200_unpack_codes(const uint8_t* packed, int64_t n, int64_t code_size, int M, int bits, int32_t* out) {
  RB_REQUIRE(packed && out, "null argument");
  RB_REQUIRE(n >= 0 && M >= 1 && bits >= 1 && bits <= 24, "need n >= 0, M >= 1, 1 <= bits <= 24");
  RB_REQUIRE(code_size * 8 >= (int64_t)M * bits, "code_size=%lld bytes cannot hold %d codes of %d bits",
             (long long)code_size, M, bits);
  for (int64_t r = 0; r < n; ++r) {
    const uint8_t* code = packed + r * code_size;
    int64_t i = 0;                                  // bit cursor, as BitstringReader::i
    for (int m = 0; m < M; ++m) {
      uint32_t v = 0;
      for (int b = 0; b < bits; ++b, ++i) v |= (uint32_t)((code[i >> 3] >> (i & 7)) & 1u) << b;
      out[r * M + m] = (int32_t)v;
    }
  }
  return 0;
}

}  // extern "C"
