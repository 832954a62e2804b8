// Host-side integer preprocessing of a question batch: bit-exact restatement of
// right_align (002_train_vqa_arch1/misc/RNNUtils.lua:54-61) and of the sort / time-major packing
// done by sort_encoding_onehot_right_align (misc/RNNUtils.lua:84-125), minus the dense one-hot
// expansion (the one-hot nn.Linear is executed as a gather, K2).  These run on the host in the
// reference as well (torch.sort / Lua loops inside dataset:next_batch, 002_train_baseline.lua:195-222).
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/nvqa.h"

namespace nvqa { void set_error(const std::string& msg); }

extern "C" int nvqa_right_align(const int32_t* seq, const int32_t* lengths, int32_t nq, int32_t T, int32_t* out) {
  if (!seq || !lengths || !out || nq < 0 || T <= 0) { nvqa::set_error("nvqa_right_align: bad argument"); return 1; }
  for (int32_t i = 0; i < nq; ++i) {
    int32_t n = lengths[i];
    if (n < 0 || n > T) { nvqa::set_error("nvqa_right_align: length out of range"); return 1; }
    int32_t* o = out + (int64_t)i * T;
    const int32_t* s = seq + (int64_t)i * T;
    std::fill(o, o + (T - n), 0);
    std::copy(s, s + n, o + (T - n));
  }
  return 0;
}

extern "C" int nvqa_pack_batch(const int32_t* q, const int32_t* lengths, int32_t B, int32_t T, int32_t* words,
                               int32_t* batch_sizes, int32_t* sort_index, int32_t* sort_index_inverse,
                               int32_t* n_words, int32_t* n_steps) {
  if (!q || !lengths || !words || !batch_sizes || !sort_index || !sort_index_inverse || B <= 0 || T <= 0) {
    nvqa::set_error("nvqa_pack_batch: bad argument");
    return 1;
  }
  std::vector<int32_t> order(B);
  std::iota(order.begin(), order.end(), 0);
  // torch.sort(len, true); TH's quicksort is unstable and its tie order unknowable:
  // the restatement fixes a STABLE descending order.
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return lengths[a] > lengths[b]; });
  for (int32_t r = 0; r < B; ++r) {
    if (lengths[order[r]] < 0 || lengths[order[r]] > T) { nvqa::set_error("nvqa_pack_batch: length out of range"); return 1; }
    sort_index[r] = order[r] + 1;                 // 1-based like Torch
    sort_index_inverse[order[r]] = r + 1;         // inverse_mapping(): x = y[inv]
  }
  const int32_t L = lengths[order[0]];
  int32_t cnt = 0;
  for (int32_t i = 0; i < L; ++i) {
    // rows (in sorted order) whose length >= L - i are active at packed step i  (:96-99)
    int32_t n = 0;
    while (n < B && lengths[order[n]] >= L - i) ++n;
    const int32_t col = T - L + i;
    for (int32_t r = 0; r < n; ++r) words[cnt + r] = q[(int64_t)order[r] * T + col];
    batch_sizes[i] = n;
    cnt += n;
  }
  if (n_words) *n_words = cnt;
  if (n_steps) *n_steps = L;
  return 0;
}

// Multiple-choice answer selection (002_train_vqa_arch1/004_eval_model.lua:257-271): for every question the argmax of
// its scores restricted to the non-zero candidate answer ids of MC_ans_test (1-based ids, 0 = padding); ties keep the
// FIRST candidate in list order (torch.max over the gathered values).  Host loop in the reference as well.
extern "C" int nvqa_mc_select(const float* scores, const int32_t* mc_ids, int32_t n, int32_t O, int32_t K, int32_t* out) {
  if (!scores || !mc_ids || !out || n < 0 || O <= 0 || K <= 0) { nvqa::set_error("nvqa_mc_select: bad argument"); return 1; }
  for (int32_t i = 0; i < n; ++i) {
    const float* s = scores + (int64_t)i * O;
    const int32_t* c = mc_ids + (int64_t)i * K;
    int32_t best = 0;
    float bv = 0.f;
    for (int32_t j = 0; j < K; ++j) {
      if (c[j] == 0) continue;
      if (c[j] < 1 || c[j] > O) { nvqa::set_error("nvqa_mc_select: candidate id out of range"); return 1; }
      const float v = s[c[j] - 1];
      if (best == 0 || v > bv) { best = c[j]; bv = v; }
    }
    out[i] = best;           // 0 when a question has no candidate (the reference would raise on an empty tensor)
  }
  return 0;
}
