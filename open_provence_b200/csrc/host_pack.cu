// host_pack.cu -- host-side block assembly of process() (no device work; compiled into the same C-ABI
// library so that the reference-side binding has one shared object to load).
//
// Restates, over flat int32 arrays instead of per-token Python lists:
//   fragment windows            open_provence/modeling_open_provence_standalone.py:686-713
//   empty-fragment filter       :846-894  (the decode itself stays with the tokenizer: the caller passes
//                                          a per-token "decodes to something visible" table and is asked to
//                                          decode only the fragments that table cannot decide)
//   greedy block packing        :2222-2259 (+ truncation :2082)
//   block ids, context location, fragment ranges    :2104-2184
//   title-prefix offset quirk   :3076-3080
//   sentence -> fragment slots  :3094-3099
// Output is the packed table the device path consumes directly: ids [T], block offsets, fragment
// (block, start, end) slots and the sentence -> slot CSR.
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/opv.h"

void opv_detail_set_error(const char* message);  // engine.cu (thread-local string behind opv_last_error)

namespace {

struct RawFragment {
  int64_t start;  // offset into in.tokens
  int32_t len;
  int32_t context;
  int32_t sentence;  // sentence index inside the context
  int32_t fragment;  // window index inside the sentence
  int32_t global;    // index among the context's windows before filtering
};

struct Pack {
  // windows before the empty-text filter
  std::vector<RawFragment> raw;
  std::vector<uint8_t> raw_uncertain;
  std::vector<int64_t> raw_start;
  std::vector<int32_t> raw_len;
  int64_t n_uncertain = 0;
  int32_t needs_decode = 0;
  // packed table
  std::vector<int32_t> ids;
  std::vector<int64_t> block_offsets{0};
  std::vector<int32_t> block_context;
  std::vector<int32_t> frag_block;
  std::vector<int32_t> frag_local;  // [n_slots, 2]
  std::vector<int32_t> sent_slot_offsets{0};
  std::vector<int32_t> sent_slot_index;
  std::vector<int64_t> ctx_block_offsets{0};
  int64_t n_sentences = 0;
};

int fail(int code, const std::string& message) {
  opv_detail_set_error(message.c_str());
  return code;
}

// standalone:686-713
void make_windows(const opv_pack_input& in, Pack& p) {
  const int64_t step = in.max_fragment_tokens > 1 ? in.max_fragment_tokens : 1;
  for (int32_t c = 0; c < in.n_contexts; ++c) {
    int32_t global = 0;
    for (int64_t s = in.h_ctx_sent_offsets[c]; s < in.h_ctx_sent_offsets[c + 1]; ++s) {
      const int64_t begin = in.h_sent_offsets[s], n = in.h_sent_offsets[s + 1] - begin;
      if (n <= 0) continue;
      const int32_t s_local = static_cast<int32_t>(s - in.h_ctx_sent_offsets[c]);
      if (in.keep_sentence_boundaries && n <= in.max_fragment_tokens) {
        p.raw.push_back({begin, static_cast<int32_t>(n), c, s_local, 0, global++});
        continue;
      }
      int32_t f = 0;
      for (int64_t at = 0; at < n; at += step, ++f) {
        const int64_t len = n - at < step ? n - at : step;
        p.raw.push_back({begin + at, static_cast<int32_t>(len), c, s_local, f, global++});
      }
    }
  }
}

// A window is certainly non-empty text when one of its tokens decodes, on its own, to something visible.
void classify(const opv_pack_input& in, Pack& p) {
  const size_t n = p.raw.size();
  p.raw_uncertain.assign(n, 0);
  p.raw_start.resize(n);
  p.raw_len.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const RawFragment& f = p.raw[i];
    p.raw_start[i] = f.start;
    p.raw_len[i] = f.len;
    bool visible = false;
    if (in.h_token_visible != nullptr) {
      for (int32_t t = 0; t < f.len && !visible; ++t) {
        const int32_t id = in.h_tokens[f.start + t];
        visible = id >= 0 && id < in.vocab_size && in.h_token_visible[id] != 0;
      }
    }
    if (!visible) {
      p.raw_uncertain[i] = 1;
      ++p.n_uncertain;
    }
  }
}

void close_block(const opv_pack_input& in, Pack& p, int32_t c, const std::vector<RawFragment>& cur) {
  if (cur.empty()) return;
  const int32_t q = in.h_ctx_query[c];
  const int32_t* q_tok = in.h_query_tokens + in.h_query_offsets[q];
  const int64_t q_len = in.h_query_offsets[q + 1] - in.h_query_offsets[q];
  const size_t base = p.ids.size();
  p.ids.insert(p.ids.end(), in.h_head, in.h_head + in.n_head);
  p.ids.insert(p.ids.end(), q_tok, q_tok + q_len);
  p.ids.insert(p.ids.end(), in.h_mid, in.h_mid + in.n_mid);
  const size_t ctx_at = p.ids.size() - base;
  for (const RawFragment& f : cur) p.ids.insert(p.ids.end(), in.h_tokens + f.start, in.h_tokens + f.start + f.len);
  const size_t ctx_len = p.ids.size() - base - ctx_at;
  p.ids.insert(p.ids.end(), in.h_tail, in.h_tail + in.n_tail);
  const int64_t n = static_cast<int64_t>(p.ids.size() - base);
  const int32_t block = static_cast<int32_t>(p.block_context.size());
  p.block_context.push_back(c);
  p.block_offsets.push_back(static_cast<int64_t>(p.ids.size()));

  // the reference locates the context by its FIRST occurrence in the block (standalone:2159-2178)
  const int32_t* ids = p.ids.data() + base;
  size_t start = ctx_at;
  for (size_t i = 0; i < ctx_at; ++i) {
    if (ids[i] == ids[ctx_at] && std::memcmp(ids + i, ids + ctx_at, ctx_len * sizeof(int32_t)) == 0) {
      start = i;
      break;
    }
  }
  const int64_t sent0 = in.h_ctx_sent_offsets[c];
  int64_t cursor = static_cast<int64_t>(start);
  for (const RawFragment& f : cur) {
    int64_t s = cursor, e = cursor + f.len;
    cursor = e;
    // title quirk (standalone:3076-3080): token counts of the prefix sentences that precede this one
    const int32_t n_pre = f.sentence < in.h_ctx_prefix[c] ? f.sentence : in.h_ctx_prefix[c];
    const int64_t offset = in.h_sent_offsets[sent0 + n_pre] - in.h_sent_offsets[sent0];
    s = s - offset > 0 ? s - offset : 0;
    e = e - offset > s ? e - offset : s;
    if (e > n) e = n;
    if (s > n) s = n;
    p.frag_block.push_back(block);
    p.frag_local.push_back(static_cast<int32_t>(s));
    p.frag_local.push_back(static_cast<int32_t>(e));
  }
}

// standalone:2222-2259 for one context; `kept` are its fragments after the filter, in order
void pack_context(const opv_pack_input& in, Pack& p, int32_t c, const std::vector<RawFragment>& kept) {
  const int32_t q = in.h_ctx_query[c];
  const int64_t q_len = in.h_query_offsets[q + 1] - in.h_query_offsets[q];
  const int64_t available = static_cast<int64_t>(in.max_length) - 2;
  const int64_t base = q_len + in.sep_len;
  const int64_t capacity = available - base > 1 ? available - base : 1;
  const int32_t first_slot = static_cast<int32_t>(p.frag_block.size());
  std::vector<RawFragment> cur;
  int64_t cur_len = base;
  for (RawFragment f : kept) {
    if (cur_len + f.len <= available) {
      cur.push_back(f);
      cur_len += f.len;
      continue;
    }
    close_block(in, p, c, cur);
    cur.clear();
    if (f.len > capacity) f.len = static_cast<int32_t>(capacity);  // standalone:2082
    cur.push_back(f);
    cur_len = base + f.len;
  }
  close_block(in, p, c, cur);
  p.ctx_block_offsets.push_back(static_cast<int64_t>(p.block_context.size()));

  // sentence -> fragment slots (standalone:3094-3099); slots were issued in fragment order
  const int64_t n_sent = in.h_ctx_sent_offsets[c + 1] - in.h_ctx_sent_offsets[c];
  size_t k = 0;
  for (int64_t s = 0; s < n_sent; ++s) {
    while (k < kept.size() && kept[k].sentence == s) {
      p.sent_slot_index.push_back(first_slot + static_cast<int32_t>(k));
      ++k;
    }
    p.sent_slot_offsets.push_back(static_cast<int32_t>(p.sent_slot_index.size()));
  }
  p.n_sentences += n_sent;
}

int validate(const opv_pack_input& in) {
  if (in.abi_version != OPV_ABI_VERSION) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_input.abi_version mismatch");
  if (in.n_contexts < 0 || in.n_queries < 0) return fail(OPV_ERR_INVALID_ARGUMENT, "negative context / query count");
  if (in.n_contexts > 0 && (!in.h_ctx_sent_offsets || !in.h_ctx_query || !in.h_ctx_prefix || !in.h_sent_offsets ||
                            !in.h_query_offsets))
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: null offset table");
  if ((in.n_head > 0 && !in.h_head) || (in.n_mid > 0 && !in.h_mid) || (in.n_tail > 0 && !in.h_tail))
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: null special-token template");
  if (in.n_head < 0 || in.n_mid < 0 || in.n_tail < 0)
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: negative template length");
  for (int32_t c = 0; c < in.n_contexts; ++c) {
    if (in.h_ctx_query[c] < 0 || in.h_ctx_query[c] >= in.n_queries)
      return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: context " + std::to_string(c) + " names query " +
                                                std::to_string(in.h_ctx_query[c]) + " (have " +
                                                std::to_string(in.n_queries) + ")");
    if (in.h_ctx_sent_offsets[c + 1] < in.h_ctx_sent_offsets[c] || in.h_ctx_prefix[c] < 0)
      return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: sentence offsets must be non-decreasing");
  }
  for (int32_t q = 0; q < in.n_queries && in.n_contexts > 0; ++q)
    if (in.h_query_offsets[q + 1] < in.h_query_offsets[q] || in.h_query_offsets[q] < 0)
      return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: query offsets must be non-negative and non-decreasing");
  if (in.n_contexts > 0 && in.n_queries > 0 && in.h_query_offsets[in.n_queries] > 0 && !in.h_query_tokens)
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: null query token array");
  if (in.n_contexts > 0 && (in.h_ctx_sent_offsets[0] < 0 || in.h_sent_offsets[in.h_ctx_sent_offsets[0]] < 0))
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: offsets must be non-negative");
  if (in.max_length > (1 << 24) || in.max_fragment_tokens > (1 << 24))
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: max_length / max_fragment_tokens out of range");
  const int64_t n_sent = in.n_contexts > 0 ? in.h_ctx_sent_offsets[in.n_contexts] : 0;
  for (int64_t s = 0; s < n_sent; ++s)
    if (in.h_sent_offsets[s + 1] < in.h_sent_offsets[s])
      return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: token offsets must be non-decreasing");
  if (n_sent > 0 && in.h_sent_offsets[n_sent] > 0 && !in.h_tokens)
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: null token array");
  return OPV_OK;
}

}  // namespace

extern "C" {

int opv_pack_build(const opv_pack_input* in, opv_pack_handle* out) {
  if (in == nullptr || out == nullptr) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: null argument");
  *out = nullptr;
  if (int rc = validate(*in)) return rc;
  Pack* p = new (std::nothrow) Pack();
  if (p == nullptr) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: out of memory");
  try {
    make_windows(*in, *p);
    classify(*in, *p);
    if (in->h_frag_drop == nullptr && p->n_uncertain > 0) {
      // a zero-length window always decodes to "" (never produced above, kept for safety); everything else
      // that no visible token settles has to be decoded by the tokenizer: hand the list back
      p->needs_decode = 1;
      *out = p;
      return OPV_OK;
    }
    std::vector<RawFragment> kept;
    size_t i = 0;
    for (int32_t c = 0; c < in->n_contexts; ++c) {
      kept.clear();
      const size_t first = i;
      for (; i < p->raw.size() && p->raw[i].context == c; ++i) {
        const bool drop = p->raw_uncertain[i] && in->h_frag_drop != nullptr && in->h_frag_drop[i] != 0;
        if (!drop) kept.push_back(p->raw[i]);
      }
      if (kept.empty() && i > first) kept.push_back(p->raw[first]);  // standalone:826-842: never drop all
      pack_context(*in, *p, c, kept);
    }
  } catch (const std::bad_alloc&) {
    delete p;
    return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_build: out of memory");
  }
  *out = p;
  return OPV_OK;
}

int opv_pack_view_get(opv_pack_handle handle, opv_pack_view* view) {
  if (handle == nullptr || view == nullptr) return fail(OPV_ERR_INVALID_ARGUMENT, "opv_pack_view_get: null argument");
  const Pack* p = static_cast<const Pack*>(handle);
  view->needs_decode = p->needs_decode;
  view->n_raw_fragments = static_cast<int64_t>(p->raw.size());
  view->n_uncertain = p->n_uncertain;
  view->h_raw_uncertain = p->raw_uncertain.data();
  view->h_raw_start = p->raw_start.data();
  view->h_raw_len = p->raw_len.data();
  view->n_blocks = static_cast<int64_t>(p->block_context.size());
  view->n_tokens = static_cast<int64_t>(p->ids.size());
  view->n_slots = static_cast<int64_t>(p->frag_block.size());
  view->n_sentences = p->n_sentences;
  view->n_contexts = static_cast<int64_t>(p->ctx_block_offsets.size()) - 1;
  view->h_ids = p->ids.data();
  view->h_block_offsets = p->block_offsets.data();
  view->h_block_context = p->block_context.data();
  view->h_frag_block = p->frag_block.data();
  view->h_frag_local = p->frag_local.data();
  view->h_sent_slot_offsets = p->sent_slot_offsets.data();
  view->h_sent_slot_index = p->sent_slot_index.data();
  view->h_ctx_block_offsets = p->ctx_block_offsets.data();
  return OPV_OK;
}

int opv_pack_destroy(opv_pack_handle handle) {
  delete static_cast<Pack*>(handle);
  return OPV_OK;
}

}  // extern "C"
