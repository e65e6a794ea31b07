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
