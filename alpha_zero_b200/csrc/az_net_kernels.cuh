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
#pragma unroll 4
    for (int p = 0; p < g.planes; ++p) dst[p] = (T)(float)obs_base[src + (size_t)p * g.nc + c];
    for (int p = g.planes; p < g.cin_pad; ++p) dst[p] = (T)0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// heads (network.py:127-156) + softmax over ALL actions (pipeline.py:108): one CTA per leaf.
template <typename T>
static __global__ void __launch_bounds__(128) k_heads(const T* __restrict__ feat, const int32_t* __restrict__ row_list,
                                               const int32_t* __restrict__ n_rows, HeadParams hp, NetGeom g, int C, int A,
                                               int fc, float* __restrict__ priors, float* __restrict__ values, int pri_stride) {
  const int leaf = blockIdx.x;
  if (leaf >= *n_rows) return;
  extern __shared__ float sh[];
  const int HW = g.Hc * g.Hc;
  float* s_pol = sh;                 // [2*HW]  flatten order (c, y, x)
  float* s_val = s_pol + 2 * HW;     // [HW]
  float* s_fc = s_val + HW;          // [fc]
  float* s_log = s_fc + fc;          // [A]
  __shared__ float s_red[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1x1 convs: one warp per position, lanes split the channels
  for (int pos = warp; pos < HW; pos += 4) {
    const int y = pos / g.Hc, x = pos - y * g.Hc;
    const T* f = feat + ((size_t)g.guard + (size_t)leaf * g.RP + (size_t)y * g.Wr + x) * C;
    float p0 = 0.f, p1 = 0.f, v0 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float a = (float)f[c];
      p0 = fmaf(a, hp.pol_w[c], p0);
      p1 = fmaf(a, hp.pol_w[C + c], p1);
      v0 = fmaf(a, hp.val_w[c], v0);
    }
    for (int o = 16; o > 0; o >>= 1) {
      p0 += __shfl_xor_sync(0xffffffffu, p0, o);
      p1 += __shfl_xor_sync(0xffffffffu, p1, o);
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    }
    if (lane == 0) {
      s_pol[pos] = fmaxf(p0 + hp.pol_b[0], 0.f);
      s_pol[HW + pos] = fmaxf(p1 + hp.pol_b[1], 0.f);
      s_val[pos] = fmaxf(v0 + hp.val_b[0], 0.f);
    }
  }
  __syncthreads();
  // policy FC and value FC1: one warp per output row, coalesced weight reads
  for (int a = warp; a < A + fc; a += 4) {
    float acc = 0.f;
    if (a < A) {
      const float* wr = hp.pol_fc_w + (size_t)a * 2 * HW;
      for (int k = lane; k < 2 * HW; k += 32) acc = fmaf(s_pol[k], wr[k], acc);
    } else {
      const float* wr = hp.val_fc1_w + (size_t)(a - A) * HW;
      for (int k = lane; k < HW; k += 32) acc = fmaf(s_val[k], wr[k], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (a < A) s_log[a] = acc + hp.pol_fc_b[a];
      else s_fc[a - A] = fmaxf(acc + hp.val_fc1_b[a - A], 0.f);
    }
  }
  __syncthreads();
  // softmax
  float mx = -INFINITY;
  for (int a = tid; a < A; a += 128) mx = fmaxf(mx, s_log[a]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int a = tid; a < A; a += 128) {
    const float e = expf(s_log[a] - mx);
    s_log[a] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = s_red[0] + s_red[1] + s_red[2] + s_red[3];
  const size_t orow = (size_t)(row_list ? row_list[leaf] : leaf);
  for (int a = tid; a < A; a += 128) priors[orow * pri_stride + a] = s_log[a] / sum;
  if (warp == 0) {
    float acc = 0.f;
    for (int k = lane; k < fc; k += 32) acc = fmaf(s_fc[k], hp.val_fc2_w[k], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) values[orow] = tanhf(acc + hp.val_fc2_b[0]);
  }
}

