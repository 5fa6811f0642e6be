// Host-side probe of emcee_b200/csrc/philox.cuh (the same header the kernels
// include), built by tests/test_host_philox.py with g++ and compared with the
// numpy statement of the draw specification (oracle/philox.py).
#include <cmath>
#include <cstdint>
using std::sqrt;
#include "../../emcee_b200/csrc/philox.cuh"

extern "C" {
void probe_draw_words(uint64_t seed, uint64_t step, uint32_t split, uint32_t tag, const uint32_t* index, int n,
                      uint32_t* out /* [n,4] */) {
  for (int i = 0; i < n; ++i) {
    eb::u32x4 w = eb::draw_words(seed, step, split, tag, index[i]);
    out[4 * i + 0] = w.x; out[4 * i + 1] = w.y; out[4 * i + 2] = w.z; out[4 * i + 3] = w.w;
  }
}
void probe_u53(const uint32_t* lo, const uint32_t* hi, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = eb::u53(lo[i], hi[i]);
}
void probe_bounded64(const uint32_t* lo, const uint32_t* hi, int n, uint64_t bound, uint64_t* out) {
  for (int i = 0; i < n; ++i) out[i] = eb::bounded64(lo[i], hi[i], bound);
}
void probe_split_permutation(uint64_t seed, uint64_t step, uint64_t n, int64_t* out) {
  eb::FeistelKeys fk = eb::feistel_keys(seed, step);
  int h = eb::feistel_half_bits(n);
  for (uint64_t w = 0; w < n; ++w) out[w] = (int64_t)eb::split_permute(w, n, h, fk);
}
void probe_de_pair(const uint64_t* m, int cnt, uint64_t n, uint64_t* p0, uint64_t* p1) {
  for (int i = 0; i < cnt; ++i) eb::de_pair_decode(m[i], n, p0[i], p1[i]);
}
}
