// az_net_kernels.cuh — kernels shared by the fp32 and the tensor-core tower (input planes -> feature rows, heads).
// Included by both translation units; internal linkage so each gets its own device code.
#pragma once
#include "az_net_impl.h"

// ------------------------------------------------------------------------------------------------
// input: int8 observation planes -> padded feature rows (CIN_PAD channels, planes 17.. zero)
template <typename T>
static __global__ void k_net_input(const int8_t* __restrict__ obs_base, const int32_t* __restrict__ row_list,
                            const int32_t* __restrict__ n_rows, T* __restrict__ act, NetGeom g) {
  const int n = *n_rows;
  const long long total = (long long)n * g.nc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int leaf = (int)(i / g.nc), c = (int)(i - (long long)leaf * g.nc);
    const int y = c / g.n, x = c - y * g.n;
    const size_t src = (size_t)(row_list ? row_list[leaf] : leaf) * g.obs_bytes;
    T* dst = act + ((size_t)g.guard + (size_t)leaf * g.RP + (size_t)(y + g.off) * g.Wr + (x + g.off)) * g.cin_pad;
    // one feature row = cin_pad channels = 128 bytes: assemble the 32 leading channels in registers, store 16 bytes at a time
    __align__(16) T vals[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) vals[p] = (T)((p < g.planes) ? (float)obs_base[src + (size_t)p * g.nc + c] : 0.f);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const uint4* s4 = reinterpret_cast<const uint4*>(vals);
    constexpr int kData = (int)(32 * sizeof(T) / 16);
    const int nvec = (int)(g.cin_pad * sizeof(T) / 16);
#pragma unroll
    for (int k = 0; k < kData; ++k) d4[k] = s4[k];
    // in_dup (split-bf16 tower): channels 32..63 repeat the planes, the layer multiplies them with the low weight terms
    for (int k = kData; k < nvec; ++k) d4[k] = (g.in_dup && k < 2 * kData) ? s4[k - kData] : make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// heads (network.py:127-156) + softmax over ALL actions (pipeline.py:108).
// One CTA (8 warps) handles LPB leaves so that every FC weight row fetched from L2 is used LPB times.
//   phase 1  1x1 policy (2 ch) / value (1 ch) convs + folded BN + ReLU: one warp per board position
//   phase 2  policy FC (A rows) and value FC1 (fc rows): one warp per output row, LPB accumulators
//   phase 3  softmax / value FC2 + tanh: one warp per leaf
template <typename T>
__device__ __forceinline__ float head_ld(const T* p) { return (float)*p; }

template <typename T, int LPB>
static __global__ void __launch_bounds__(256) k_heads(const T* __restrict__ feat, const int32_t* __restrict__ row_list,
                                                      const int32_t* __restrict__ n_rows, HeadParams hp, NetGeom g, int C, int fstride,
                                                      int split, int A, int fc, float* __restrict__ priors, float* __restrict__ values,
                                                      int pri_stride) {
  const int n = *n_rows;
  const int leaf0 = blockIdx.x * LPB;
  if (leaf0 >= n) return;
  const int nl = min(LPB, n - leaf0);
  extern __shared__ float sh[];
  const int HW = g.Hc * g.Hc;
  const int per = 3 * HW + fc + A;       // floats per leaf: pol[2*HW] | val[HW] | fc1[fc] | logits[A]
  float* s_w = sh + (size_t)LPB * per;   // [3][C] folded 1x1 conv weights (policy 0, policy 1, value)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 3 * C; i += 256) s_w[i] = i < 2 * C ? hp.pol_w[i] : hp.val_w[i - 2 * C];
  __syncthreads();
  // ---- phase 1: one thread per board position, channels streamed 16 bytes at a time, no shuffles
  for (int idx = tid; idx < nl * HW; idx += 256) {
    const int l = idx / HW, pos = idx - l * HW;
    const int y = pos / g.Hc, x = pos - y * g.Hc;
    // a feature row holds `fstride` elements; split rows are [hi(C) | lo(C) | hi(C)] and the value of a channel is hi + lo
    const T* f = feat + ((size_t)g.guard + (size_t)(leaf0 + l) * g.RP + (size_t)y * g.Wr + x) * fstride;
    float p0 = 0.f, p1 = 0.f, v0 = 0.f;
    if (sizeof(T) == 2) {
      const uint4* f4 = reinterpret_cast<const uint4*>(f);
      const int nchunk = (split ? 2 * C : C) / 8;  // a multiple of 8 (C is a multiple of 64)
      for (int c80 = 0; c80 < nchunk; c80 += 4) {
        // four 16-byte loads of the row in flight before the first is consumed (rows of neighbouring threads are 256+ bytes apart:
        // the loop is latency bound, not bandwidth bound)
        uint4 raw4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) raw4[u] = f4[c80 + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c8 = c80 + u;
          const uint32_t w[4] = {raw4[u].x, raw4[u].y, raw4[u].z, raw4[u].w};
          const int cb = c8 * 8 >= C ? c8 * 8 - C : c8 * 8;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float a0 = __uint_as_float(w[k] << 16), a1 = __uint_as_float(w[k] & 0xffff0000u);
            const int c = cb + k * 2;
            p0 = fmaf(a0, s_w[c], fmaf(a1, s_w[c + 1], p0));
            p1 = fmaf(a0, s_w[C + c], fmaf(a1, s_w[C + c + 1], p1));
            v0 = fmaf(a0, s_w[2 * C + c], fmaf(a1, s_w[2 * C + c + 1], v0));
          }
        }
      }
    } else {
      for (int c = 0; c < C; ++c) {
        const float a = head_ld(f + c);
        p0 = fmaf(a, s_w[c], p0);
        p1 = fmaf(a, s_w[C + c], p1);
        v0 = fmaf(a, s_w[2 * C + c], v0);
      }
    }
    float* s = sh + (size_t)l * per;
    s[pos] = fmaxf(p0 + hp.pol_b[0], 0.f);
    s[HW + pos] = fmaxf(p1 + hp.pol_b[1], 0.f);
    s[2 * HW + pos] = fmaxf(v0 + hp.val_b[0], 0.f);
  }
  __syncthreads();
  // ---- phase 2: one thread per FC output row (policy logits, value hidden units), LPB accumulators,
  //      weights read k-major (transposed at az_set_weights) so that consecutive threads read consecutive floats
  for (int a = tid; a < A + fc; a += 256) {
    float acc[LPB];
#pragma unroll
    for (int l = 0; l < LPB; ++l) acc[l] = 0.f;
    const bool is_pol = a < A;
    const float* wt = is_pol ? hp.pol_fc_wT + a : hp.val_fc1_wT + (a - A);
    const int klen = is_pol ? 2 * HW : HW, ldw = is_pol ? A : fc, soff = is_pol ? 0 : 2 * HW;
    for (int k = 0; k < klen; ++k) {
      const float w = wt[(size_t)k * ldw];
#pragma unroll
      for (int l = 0; l < LPB; ++l) acc[l] = fmaf(sh[(size_t)l * per + soff + k], w, acc[l]);
    }
    const float bv = is_pol ? hp.pol_fc_b[a] : hp.val_fc1_b[a - A];
#pragma unroll
    for (int l = 0; l < LPB; ++l) {
      if (l < nl) {
        float* s = sh + (size_t)l * per;
        if (is_pol) s[3 * HW + fc + a] = acc[l] + bv;
        else s[3 * HW + (a - A)] = fmaxf(acc[l] + bv, 0.f);
      }
    }
  }
  __syncthreads();
  // ---- phase 3: softmax / value FC2 + tanh, one warp per leaf
  for (int l = warp; l < nl; l += 8) {
    float* s = sh + (size_t)l * per;
    float* lg = s + 3 * HW + fc;
    float mx = -INFINITY;
    for (int a = lane; a < A; a += 32) mx = fmaxf(mx, lg[a]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int a = lane; a < A; a += 32) {
      const float e = expf(lg[a] - mx);
      lg[a] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const size_t orow = (size_t)(row_list ? row_list[leaf0 + l] : leaf0 + l);
    for (int a = lane; a < A; a += 32) priors[orow * pri_stride + a] = lg[a] / sum;
    float acc = 0.f;
    for (int k = lane; k < fc; k += 32) acc = fmaf(s[3 * HW + k], hp.val_fc2_w[k], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) values[orow] = tanhf(acc + hp.val_fc2_b[0]);
  }
}

// host helper: launch the heads with the largest LPB whose shared memory fits the default 48 KB
template <typename T>
static inline void launch_heads(cudaStream_t stream, const T* feat, const int32_t* row_list, const int32_t* n_rows, const HeadParams& hp,
                                const NetGeom& g, int C, int fstride, int split, int A, int fc, float* priors, float* values, int pri_stride,
                                int max_rows) {
  const int HW = g.Hc * g.Hc;
  const size_t per = (size_t)(3 * HW + fc + A) * sizeof(float);
  const size_t wsm = (size_t)3 * C * sizeof(float);
  if (per * 8 + wsm <= 48 * 1024)
    k_heads<T, 8><<<(max_rows + 7) / 8, 256, per * 8 + wsm, stream>>>(feat, row_list, n_rows, hp, g, C, fstride, split, A, fc, priors, values, pri_stride);
  else if (per * 4 + wsm <= 48 * 1024)
    k_heads<T, 4><<<(max_rows + 3) / 4, 256, per * 4 + wsm, stream>>>(feat, row_list, n_rows, hp, g, C, fstride, split, A, fc, priors, values, pri_stride);
  else
    k_heads<T, 1><<<max_rows, 256, per + wsm, stream>>>(feat, row_list, n_rows, hp, g, C, fstride, split, A, fc, priors, values, pri_stride);
}
