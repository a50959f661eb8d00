// Fuzz driver for the host-side block assembly (open_provence_b200/csrc/host_pack.cu), built by
// tests/test_host_pack_sanitizers.py with g++ -fsanitize=address,undefined.  Random flat inputs (empty contexts and
// sentences, queries longer than the block capacity, every template shape) go through opv_pack_build /
// opv_pack_view_get / opv_pack_destroy; the driver checks the structural invariants of the returned table and exits
// non-zero on the first violation (the sanitizers abort on any out-of-bounds access or undefined behaviour).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../../include/opv.h"

static std::string g_error;
void opv_detail_set_error(const char* message) { g_error = message ? message : ""; }

#define CHECK(cond)                                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      std::fprintf(stderr, "case %d: invariant failed: %s (line %d)\n", iter, #cond, __LINE__); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

int main(int argc, char** argv) {
  const int iterations = argc > 1 ? std::atoi(argv[1]) : 2000;
  std::mt19937 rng(12345);
  auto uni = [&](int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); };
  for (int iter = 0; iter < iterations; ++iter) {
    const int n_q = uni(1, 3), n_ctx = uni(0, 5), vocab = 64;
    std::vector<int32_t> q_tok;
    std::vector<int64_t> q_off{0};
    for (int q = 0; q < n_q; ++q) {
      const int n = uni(0, iter % 7 == 0 ? 40 : 6);
      for (int i = 0; i < n; ++i) q_tok.push_back(uni(0, vocab - 1));
      q_off.push_back(static_cast<int64_t>(q_tok.size()));
    }
    std::vector<int32_t> tokens;
    std::vector<int64_t> sent_off{0}, ctx_sent{0};
    std::vector<int32_t> ctx_query, ctx_prefix;
    for (int c = 0; c < n_ctx; ++c) {
      const int n_sent = uni(0, 6);
      for (int s = 0; s < n_sent; ++s) {
        const int n = uni(0, 3) == 0 ? 0 : uni(1, 30);
        for (int i = 0; i < n; ++i) tokens.push_back(uni(-2, vocab + 1));  // ids outside the visibility table too
        sent_off.push_back(static_cast<int64_t>(tokens.size()));
      }
      ctx_sent.push_back(static_cast<int64_t>(sent_off.size()) - 1);
      ctx_query.push_back(uni(0, n_q - 1));
      ctx_prefix.push_back(uni(0, 2));
    }
    std::vector<uint8_t> visible(vocab);
    for (auto& v : visible) v = uni(0, 2) != 0;
    std::vector<int32_t> head(uni(0, 2), 1), mid(uni(0, 2), 2), tail(uni(0, 1), 2);
    opv_pack_input in{};
    in.abi_version = OPV_ABI_VERSION;
    in.max_length = uni(4, 48);
    in.max_fragment_tokens = uni(1, 24);
    in.keep_sentence_boundaries = uni(0, 1);
    in.sep_len = uni(0, 2);
    in.n_contexts = n_ctx;
    in.n_queries = n_q;
    in.vocab_size = uni(0, 3) == 0 ? 0 : vocab;
    in.n_head = static_cast<int32_t>(head.size());
    in.n_mid = static_cast<int32_t>(mid.size());
    in.n_tail = static_cast<int32_t>(tail.size());
    in.h_head = head.data();
    in.h_mid = mid.data();
    in.h_tail = tail.data();
    in.h_tokens = tokens.data();
    in.h_sent_offsets = sent_off.data();
    in.h_ctx_sent_offsets = ctx_sent.data();
    in.h_ctx_query = ctx_query.data();
    in.h_ctx_prefix = ctx_prefix.data();
    in.h_query_tokens = q_tok.data();
    in.h_query_offsets = q_off.data();
    in.h_token_visible = in.vocab_size ? visible.data() : nullptr;
    in.h_frag_drop = nullptr;

    opv_pack_handle handle = nullptr;
    opv_pack_view view{};
    CHECK(opv_pack_build(&in, &handle) == OPV_OK);
    CHECK(opv_pack_view_get(handle, &view) == OPV_OK);
    std::vector<uint8_t> drop;
    if (view.needs_decode) {
      CHECK(view.n_uncertain > 0);
      drop.assign(static_cast<size_t>(view.n_raw_fragments), 0);
      for (int64_t i = 0; i < view.n_raw_fragments; ++i) {
        CHECK(view.h_raw_len[i] > 0 && view.h_raw_start[i] >= 0);
        CHECK(view.h_raw_start[i] + view.h_raw_len[i] <= static_cast<int64_t>(tokens.size()));
        if (view.h_raw_uncertain[i]) drop[static_cast<size_t>(i)] = uni(0, 1);
      }
      opv_pack_destroy(handle);
      in.h_frag_drop = drop.data();
      CHECK(opv_pack_build(&in, &handle) == OPV_OK);
      CHECK(opv_pack_view_get(handle, &view) == OPV_OK);
      CHECK(view.needs_decode == 0);
    }
    // structural invariants of the packed table
    CHECK(view.n_contexts == n_ctx);
    CHECK(view.n_sentences == ctx_sent.back());
    CHECK(view.h_block_offsets[0] == 0 && view.h_block_offsets[view.n_blocks] == view.n_tokens);
    CHECK(view.h_ctx_block_offsets[0] == 0 && view.h_ctx_block_offsets[n_ctx] == view.n_blocks);
    CHECK(view.h_sent_slot_offsets[0] == 0 && view.h_sent_slot_offsets[view.n_sentences] == view.n_slots);
    for (int64_t b = 0; b < view.n_blocks; ++b) {
      const int64_t len = view.h_block_offsets[b + 1] - view.h_block_offsets[b];
      CHECK(len > 0);
      const int32_t c = view.h_block_context[b];
      CHECK(c >= 0 && c < n_ctx && view.h_ctx_block_offsets[c] <= b && b < view.h_ctx_block_offsets[c + 1]);
    }
    std::vector<int> seen(static_cast<size_t>(view.n_slots), 0);
    for (int64_t k = 0; k < view.n_slots; ++k) {
      const int32_t b = view.h_frag_block[k];
      CHECK(b >= 0 && b < view.n_blocks);
      const int64_t len = view.h_block_offsets[b + 1] - view.h_block_offsets[b];
      CHECK(0 <= view.h_frag_local[2 * k] && view.h_frag_local[2 * k] <= view.h_frag_local[2 * k + 1]);
      CHECK(view.h_frag_local[2 * k + 1] <= len);
      CHECK(k == 0 || view.h_frag_block[k - 1] <= b);  // slots are issued block by block
    }
    for (int64_t s = 0; s < view.n_sentences; ++s) {
      CHECK(view.h_sent_slot_offsets[s] <= view.h_sent_slot_offsets[s + 1]);
      for (int32_t j = view.h_sent_slot_offsets[s]; j < view.h_sent_slot_offsets[s + 1]; ++j) {
        const int32_t k = view.h_sent_slot_index[j];
        CHECK(k >= 0 && k < view.n_slots);
        ++seen[static_cast<size_t>(k)];
      }
    }
    for (int64_t k = 0; k < view.n_slots; ++k) CHECK(seen[static_cast<size_t>(k)] == 1);  // every slot in one sentence
    // a context with at least one token keeps at least one fragment (standalone:826-842)
    for (int c = 0; c < n_ctx; ++c) {
      const bool has_tokens = sent_off[static_cast<size_t>(ctx_sent[c + 1])] > sent_off[static_cast<size_t>(ctx_sent[c])];
      CHECK((view.h_ctx_block_offsets[c + 1] > view.h_ctx_block_offsets[c]) == has_tokens);
    }
    CHECK(opv_pack_destroy(handle) == OPV_OK);
  }
  // argument validation
  int iter = -1;
  opv_pack_handle handle = nullptr;
  CHECK(opv_pack_build(nullptr, &handle) == OPV_ERR_INVALID_ARGUMENT);
  opv_pack_input bad{};
  CHECK(opv_pack_build(&bad, &handle) == OPV_ERR_INVALID_ARGUMENT && !g_error.empty());
  CHECK(opv_pack_destroy(nullptr) == OPV_OK);
  std::printf("pack_fuzz: %d cases ok\n", iterations);
  return 0;
}
