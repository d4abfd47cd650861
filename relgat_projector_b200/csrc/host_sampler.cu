// Host-side (CPU) batch construction, bit-exact with the reference's RNG stream — "next" row (f)-2.
//
// The reference draws corrupted tails with CPython's random.choice (dataset/edge.py:99-103: one draw
// per negative, redraw while it equals the true tail) over all_node_ids = range(N)
// (dataset/relgat_dataset.py:97).  CPython: choice(seq) = seq[_randbelow(len(seq))];
// _randbelow(n): k = n.bit_length(); r = getrandbits(k); while r >= n: r = getrandbits(k);
// getrandbits(k <= 32) = genrand_uint32() >> (32 - k) on the MT19937 generator.
// This file restates that generator step so a batch of B*(1+K) ids costs microseconds instead of
// B*K interpreter-level calls; the 625-word state is exchanged with random.getstate()/setstate(),
// so Python code before and after sees exactly the stream it would have seen.
#include <stdint.h>

namespace {

constexpr int kN = 624, kM = 397;

inline uint32_t mt_next(uint32_t* mt, uint32_t* pos) {
  if (*pos >= kN) {  // regenerate the block (identical to CPython's genrand_uint32)
    int kk;
    uint32_t y;
    for (kk = 0; kk < kN - kM; ++kk) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + kM] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < kN - 1; ++kk) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (kM - kN)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[kN - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    *pos = 0;
  }
  uint32_t y = mt[(*pos)++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

inline int bit_length(uint64_t n) {
  int k = 0;
  while (n) { ++k; n >>= 1; }
  return k;
}

inline uint64_t randbelow(uint32_t* mt, uint32_t* pos, uint64_t n, int k) {
  for (;;) {
    uint64_t r;
    if (k <= 32) {
      r = mt_next(mt, pos) >> (32 - k);
    } else {  // getrandbits(k > 32): little-endian 32-bit words, the last one shifted
      const uint64_t lo = mt_next(mt, pos);
      const uint64_t hi = mt_next(mt, pos) >> (64 - k);
      r = lo | (hi << 32);
    }
    if (r < n) return r;
  }
}

}  // namespace

// HOST pointers.  state: 624 MT words followed by the position index (random.getstate()[1]).
// edges: int64 [n_edges, 3] = (src, dst, rel); idxs: int64 [B] rows of `edges` forming the batch.
// Outputs int64 [B*(1+K)]: positives first, then K blocks of B negatives (K-major), exactly the
// layout of trainer/components/relgat_batching.py:5-19.
extern "C" int relgat_host_sample_batch(uint32_t* state, const long long* edges, long long n_edges,
                                        const long long* idxs, int B, int K, long long n_nodes,
                                        long long* src_out, long long* rel_out, long long* dst_out) {
  if (!state || !edges || (!idxs && B > 0) || B < 0 || K < 0 || n_nodes <= 0 || !src_out || !rel_out || !dst_out)
    return -1;
  if (K > 0 && n_nodes < 2) return -1;  // the rejection loop could not terminate
  uint32_t* mt = state;
  uint32_t pos = state[624];
  const int k = bit_length(static_cast<uint64_t>(n_nodes));
  for (int i = 0; i < B; ++i) {
    const long long ei = idxs[i];
    if (ei < 0 || ei >= n_edges) return -1;
    const long long s = edges[3 * ei], d = edges[3 * ei + 1], r = edges[3 * ei + 2];
    src_out[i] = s; rel_out[i] = r; dst_out[i] = d;
    for (int kk = 0; kk < K; ++kk) {
      long long c = static_cast<long long>(randbelow(mt, &pos, static_cast<uint64_t>(n_nodes), k));
      while (c == d) c = static_cast<long long>(randbelow(mt, &pos, static_cast<uint64_t>(n_nodes), k));
      const long long o = static_cast<long long>(B) + static_cast<long long>(kk) * B + i;
      src_out[o] = s; rel_out[o] = r; dst_out[o] = c;
    }
  }
  state[624] = pos;
  return 0;
}

// In-place random.shuffle of a permutation (CPython: for i in reversed(range(1, n)): j = _randbelow(i + 1);
// swap), used for the train / eval split of dataset/relgat_dataset.py:70-88.
extern "C" int relgat_host_shuffle(uint32_t* state, long long* perm, long long n) {
  if (!state || (!perm && n > 0) || n < 0) return -1;
  uint32_t pos = state[624];
  for (long long i = n - 1; i >= 1; --i) {
    const uint64_t m = static_cast<uint64_t>(i + 1);
    const long long j = static_cast<long long>(randbelow(state, &pos, m, bit_length(m)));
    const long long t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
  state[624] = pos;
  return 0;
}
