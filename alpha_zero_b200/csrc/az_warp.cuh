// az_warp.cuh — the warp-per-game programming layer.
//
// Every per-game routine in az_board.cuh / az_tree.cuh is written as a sequence of *phases*: a
// W_FOR(i, n) loop in which lane L owns elements L, L+32, ... of a per-game array, separated by
// w_sync().  On the GPU one warp runs one game, lanes cooperate through shuffles / ballots and
// shared-memory scratch.  With -DAZ_EMU the same source is compiled for the host with a "warp" of
// width 1: that build is TEST INFRASTRUCTURE ONLY (tests/emu, never loaded by the package) and lets the
// game / tree logic be checked against the oracle without a GPU.  A phase body may therefore only
// write element i (or use the w_* collectives); anything read from another lane's element must come
// from an earlier phase.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef AZ_EMU
#include <string.h>
#define AZ_DEV static inline
#define AZ_LANE 0
#define AZ_WIDTH 1
AZ_DEV void w_sync() {}
AZ_DEV bool w_any(bool p) { return p; }
AZ_DEV uint32_t w_ballot(bool p) { return p ? 1u : 0u; }
AZ_DEV uint32_t w_lanemask_lt() { return 0u; }
AZ_DEV int w_sum_i(int v) { return v; }
AZ_DEV int w_max_i(int v) { return v; }
AZ_DEV int w_min_i(int v) { return v; }
AZ_DEV double w_sum_d(double v) { return v; }
AZ_DEV float w_sum_f(float v) { return v; }
AZ_DEV void w_argmax(double& v, int& i) { (void)v; (void)i; }
AZ_DEV void w_argmax_f(float& v, int& i) { (void)v; (void)i; }
AZ_DEV int w_bcast_i(int v, int) { return v; }
AZ_DEV float w_bcast_f(float v, int) { return v; }
AZ_DEV float f_mul(float a, float b) { return a * b; }  // built with -ffp-contract=off
AZ_DEV float f_div(float a, float b) { return a / b; }
AZ_DEV float f_add(float a, float b) { return a + b; }
AZ_DEV double d_mul(double a, double b) { return a * b; }
AZ_DEV double d_add(double a, double b) { return a + b; }
AZ_DEV double d_div(double a, double b) { return a / b; }
AZ_DEV int atomic_add_i(int* p, int v) { int o = *p; *p = o + v; return o; }
AZ_DEV int atomic_or_i(int* p, int v) { int o = *p; *p = o | v; return o; }
AZ_DEV unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
#else
#include <cuda_runtime.h>
#define AZ_DEV __device__ __forceinline__
#define AZ_LANE ((int)(threadIdx.x & 31))
#define AZ_WIDTH 32
#define AZ_FULL 0xffffffffu
AZ_DEV void w_sync() { __syncwarp(); }
AZ_DEV bool w_any(bool p) { return __any_sync(AZ_FULL, p); }
AZ_DEV uint32_t w_ballot(bool p) { return __ballot_sync(AZ_FULL, p); }
AZ_DEV uint32_t w_lanemask_lt() { return (1u << AZ_LANE) - 1u; }
AZ_DEV int w_sum_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AZ_FULL, v, o);
  return v;
}
AZ_DEV int w_max_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(AZ_FULL, v, o));
  return v;
}
AZ_DEV int w_min_i(int v) {
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(AZ_FULL, v, o));
  return v;
}
AZ_DEV double w_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AZ_FULL, v, o);
  return v;
}
AZ_DEV float w_sum_f(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AZ_FULL, v, o);
  return v;
}
// Warp-shuffle argmax with numpy.argmax tie-breaking: the LOWEST index among equal maxima wins
// (core/mcts_v2.py:178).  Every lane ends up with the winner.
AZ_DEV void w_argmax(double& v, int& i) {
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(AZ_FULL, v, o);
    int oi = __shfl_xor_sync(AZ_FULL, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}
// The same for float32 scores (every level below a noised root) in two warp reductions instead of fifteen shuffles: the maximum
// of an order-preserving integer image of the scores (redux.sync.max), then the lowest index among the lanes that hold it
// (redux.sync.min).  A lane's (v, i) is its local best with the lowest-index tie-break already applied; v is never NaN.
AZ_DEV void w_argmax_f(float& v, int& i) {
  const uint32_t b = __float_as_uint(v);
  const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  const uint32_t top = __reduce_max_sync(AZ_FULL, key);
  i = (int)__reduce_min_sync(AZ_FULL, key == top ? (uint32_t)i : 0xffffffffu);
  const uint32_t tb = (top & 0x80000000u) ? (top & 0x7fffffffu) : ~top;
  v = __uint_as_float(tb);
}
AZ_DEV int w_bcast_i(int v, int src) { return __shfl_sync(AZ_FULL, v, src & 31); }
AZ_DEV float w_bcast_f(float v, int src) { return __shfl_sync(AZ_FULL, v, src & 31); }
// Round-to-nearest, never contracted into FMA: the selection arithmetic must reproduce numpy's
// element-wise float32 / float64 operations bit for bit.
AZ_DEV float f_mul(float a, float b) { return __fmul_rn(a, b); }
AZ_DEV float f_div(float a, float b) { return __fdiv_rn(a, b); }
AZ_DEV float f_add(float a, float b) { return __fadd_rn(a, b); }
AZ_DEV double d_mul(double a, double b) { return __dmul_rn(a, b); }
AZ_DEV double d_add(double a, double b) { return __dadd_rn(a, b); }
AZ_DEV double d_div(double a, double b) { return __ddiv_rn(a, b); }
AZ_DEV int atomic_add_i(int* p, int v) { return atomicAdd(p, v); }
AZ_DEV int atomic_or_i(int* p, int v) { return atomicOr(p, v); }
AZ_DEV unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
#endif

#define W_FOR(i, n) for (int i = AZ_LANE; i < (n); i += AZ_WIDTH)
#define W_LANE0 if (AZ_LANE == 0)

// Hint: pull the 128-byte lines of [p, p + bytes) towards L2 (one line per lane); results never depend on it.
AZ_DEV void w_prefetch(const void* p, int bytes) {
#ifndef AZ_EMU
  const int off = AZ_LANE * 128;
  if (off < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)p + off));
#else
  (void)p; (void)bytes;
#endif
}

// Warp-cooperative copy of n bytes between global buffers of ANY alignment (finished games: per-slot record -> sample ring).
// Destination words are written whole; each is assembled from the two aligned source words it straddles (funnel shift), four
// words per lane in flight.  The last source word read may extend up to 3 bytes past src + n: inside the allocation, whose size
// cudaMalloc rounds up to 256 bytes.
AZ_DEV void w_copy_bytes(void* dst_, const void* src_, size_t n) {
#ifdef AZ_EMU
  memmove(dst_, src_, n);
#else
  unsigned char* dst = (unsigned char*)dst_;
  const unsigned char* src = (const unsigned char*)src_;
  size_t head = (4 - ((size_t)dst & 3)) & 3;
  if (head > n) head = n;
  if ((size_t)AZ_LANE < head) dst[AZ_LANE] = src[AZ_LANE];
  uint32_t* d4 = (uint32_t*)(dst + head);
  const unsigned char* s = src + head;
  const size_t nw = (n - head) >> 2;
  const uint32_t r = (uint32_t)((size_t)s & 3);
  const uint32_t* sa = (const uint32_t*)(s - r);
  const uint32_t sh = r * 8;
  for (size_t j0 = 0; j0 < nw; j0 += 4 * AZ_WIDTH) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t j = j0 + (size_t)k * AZ_WIDTH + AZ_LANE;
      lo[k] = 0; hi[k] = 0;
      if (j < nw) { lo[k] = sa[j]; if (r) hi[k] = sa[j + 1]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t j = j0 + (size_t)k * AZ_WIDTH + AZ_LANE;
      if (j < nw) d4[j] = r ? __funnelshift_r(lo[k], hi[k], sh) : lo[k];
    }
  }
  const size_t done = head + nw * 4;
  if ((size_t)AZ_LANE < n - done) dst[done + AZ_LANE] = src[done + AZ_LANE];
#endif
}

// Same for 4-byte elements (both sides 4-byte aligned).
AZ_DEV void w_copy_words(void* dst_, const void* src_, size_t nwords) {
#ifdef AZ_EMU
  memmove(dst_, src_, nwords * 4);
#else
  uint32_t* d4 = (uint32_t*)dst_;
  const uint32_t* s4 = (const uint32_t*)src_;
  for (size_t j0 = 0; j0 < nwords; j0 += 4 * AZ_WIDTH) {
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t j = j0 + (size_t)k * AZ_WIDTH + AZ_LANE;
      v[k] = j < nwords ? s4[j] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t j = j0 + (size_t)k * AZ_WIDTH + AZ_LANE;
      if (j < nwords) d4[j] = v[k];
    }
  }
#endif
}

// Counter-based RNG (SplitMix64 finaliser over (seed, stream, counter)): stateless, identical on
// every lane, so a warp can draw element-wise without communication.
AZ_DEV uint64_t az_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
AZ_DEV uint64_t az_rand64(uint64_t seed, uint64_t stream, uint64_t ctr) {
  return az_mix64(az_mix64(seed ^ az_mix64(stream)) + ctr * 0xD1342543DE82EF95ull);
}
AZ_DEV double az_u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
