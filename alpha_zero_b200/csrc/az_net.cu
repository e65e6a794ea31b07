// az_net.cu — policy/value network forward (core/network.py:85-173), engine-owned parameters.
//
// Data layout in HBM.  A leaf's feature map is stored row-major as RP = (Hc+1)*(Hc+1) rows of C
// channels: row r = y*(Hc+1)+x, with column x == Hc and row y == Hc kept at zero.  Leaves are
// concatenated and the whole matrix has GUARD zero rows in front and behind.  With that padding the
// input of tap (ky,kx) for output row m is simply row m + (ky-1)*(Hc+1) + (kx-1) of the SAME matrix, so
// a 3x3 convolution is 9 accumulated GEMMs over shifted views — no im2col, every tile load is a plain
// 2-D box (which is what the TMA descriptors of the tcgen05 path need).  Go: Hc = N.  Gomoku:
// network.py:101 pads the first conv by 3, so the tower runs on Hc = N+4 with the observation placed
// at offset (2,2) and every conv is an ordinary pad-1 conv on that canvas.
//
// BatchNorm is folded on the host at az_set_weights time (eval mode, eps 1e-5):
//   w'[co] = w[co] * g/sqrt(var+eps),  b'[co] = beta - mean * g/sqrt(var+eps).
//
// Two towers share the input / head kernels:
//   AZ_NET_FP32  k_conv_f32: CUDA-core SGEMM tiles, fp32 everywhere — the parity mode.
//   AZ_NET_BF16  az_net_tc.cu: tcgen05.mma (bf16 x bf16 -> f32 in TMEM), TMA-staged operands.
#include <math.h>

#include <thread>
#include <vector>

#include "az_net.h"
#include "az_net_impl.h"
#include "az_net_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// fp32 implicit-GEMM 3x3 convolution: out = act(sum_tap A[m + off_tap] * W[tap] + bias (+ res))
// tile 64 rows x 64 output channels, K chunks of 16, 256 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256) k_conv_f32(const float* __restrict__ in, const float* __restrict__ w,
                                                  const float* __restrict__ bias, const float* res,
                                                  float* out, const int32_t* __restrict__ n_rows, NetGeom g,
                                                  int cin, int cout, int relu) {
  const long long M = (long long)(*n_rows) * g.RP;
  const long long m0 = (long long)blockIdx.x * 64;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * 64;
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int la_row = tid >> 2, la_k = (tid & 3) * 4;   // A tile: 64 rows x 16 k, one float4 per thread
  const int lb_k = tid >> 4, lb_n = (tid & 15) * 4;    // B tile: 16 k x 64 n
  for (int tap = 0; tap < 9; ++tap) {
    const int toff = (tap / 3 - 1) * g.Wr + (tap % 3 - 1);
    const float* a_base = in + ((size_t)g.guard + (size_t)(m0 + la_row) + toff) * cin;  // guard rows make this in-bounds
    const float* w_base = w + (size_t)tap * cin * cout;
    for (int k0 = 0; k0 < cin; k0 += 16) {
      const float4 av = *reinterpret_cast<const float4*>(a_base + k0 + la_k);
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + lb_n < cout) bv = *reinterpret_cast<const float4*>(w_base + (size_t)(k0 + lb_k) * cout + n0 + lb_n);
      __syncthreads();
      sA[la_k + 0][la_row] = av.x; sA[la_k + 1][la_row] = av.y; sA[la_k + 2][la_row] = av.z; sA[la_k + 3][la_row] = av.w;
      sB[lb_k][lb_n + 0] = bv.x; sB[lb_k][lb_n + 1] = bv.y; sB[lb_k][lb_n + 2] = bv.z; sB[lb_k][lb_n + 3] = bv.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int r = (int)(m % g.RP);
    const int yy = r / g.Wr, xx = r - yy * g.Wr;
    const bool valid = yy < g.Hc && xx < g.Hc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= cout) continue;
      float v = 0.f;
      if (valid) {
        v = acc[i][j] + bias[co];
        if (res) v += res[((size_t)g.guard + m) * cout + co];
        if (relu) v = fmaxf(v, 0.f);
      }
      out[((size_t)g.guard + m) * cout + co] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
static float* upload(AzRt& rt, std::vector<void*>& allocs, const std::vector<float>& v) {
  float* p = (float*)rt_alloc(v.size() * sizeof(float));
  if (!p) return nullptr;
  allocs.push_back(p);
  rt_h2d(rt, p, v.data(), v.size() * sizeof(float));
  return p;
}

// allocate on first use, refresh in place afterwards
static void dev_set(AzRt& rt, std::vector<void*>& allocs, float*& slot, const std::vector<float>& v) {
  if (!slot) {
    slot = (float*)rt_alloc(v.size() * sizeof(float));
    if (!slot) return;
    allocs.push_back(slot);
  }
  rt_h2d(rt, slot, v.data(), v.size() * sizeof(float));
}

AzNet* aznet_create(const AzDims& d, const az_config& cfg, AzRt& rt, int max_leaves, std::string& err) {
  AzNet* n = new AzNet();
  n->blocks = cfg.num_res_blocks;
  n->C = cfg.num_filters;
  n->fc = cfg.num_fc_units;
  n->A = d.A;
  n->precision = cfg.net_precision;
  n->max_leaves = max_leaves;
  NetGeom& g = n->g;
  g.n = d.n;
  g.nc = d.nc;
  g.planes = d.planes;
  g.obs_bytes = d.obs_bytes;
  g.off = d.game == AZ_GAME_GOMOKU ? 2 : 0;
  g.Hc = d.n + 2 * g.off;
  g.Wr = g.Hc + 1;
  g.RP = g.Wr * g.Wr;
  g.guard = ((g.Wr + 1 + 127) / 128) * 128;  // >= one board row + 1, and a whole number of 128-row tiles
  g.cin_pad = 32;
  if (n->C % 64 != 0 && n->precision == AZ_NET_BF16) { err = "bf16 tower needs num_filters to be a multiple of 64"; delete n; return nullptr; }
  if (n->C % 16 != 0 || n->C < 16) { err = "num_filters must be a multiple of 16"; delete n; return nullptr; }
  if (d.planes > g.cin_pad) { err = "observation has more than 32 planes"; delete n; return nullptr; }
  n->rows_total = (size_t)max_leaves * g.RP + 2 * (size_t)g.guard + 1024;
  const size_t esz = n->precision == AZ_NET_BF16 ? 2 : 4;
  n->act_in = rt_alloc(n->rows_total * g.cin_pad * esz);
  n->act_x = rt_alloc(n->rows_total * n->C * esz);
  n->act_mid = rt_alloc(n->rows_total * n->C * esz);
  if (!n->act_in || !n->act_x || !n->act_mid) { err = "activation buffers: out of device memory"; aznet_destroy(n); return nullptr; }
  // 2*MAC per evaluation, for the roofline
  const double hw = (double)g.Hc * g.Hc;
  n->flops = 2.0 * hw * 9.0 * d.planes * n->C + (double)n->blocks * 2.0 * (2.0 * hw * 9.0 * n->C * n->C) + 2.0 * hw * n->C * 3.0 +
             2.0 * (2.0 * hw) * d.A + 2.0 * hw * n->fc + 2.0 * n->fc;
  if (n->precision == AZ_NET_BF16) {
    if (aznet_tc_create(n, rt, err)) { aznet_destroy(n); return nullptr; }
  }
  return n;
}

void aznet_destroy(AzNet* n) {
  if (!n) return;
  aznet_tc_destroy(n);
  rt_free(n->act_in);
  rt_free(n->act_x);
  rt_free(n->act_mid);
  for (void* p : n->allocs) rt_free(p);
  delete n;
}

double aznet_flops_per_eval(const AzNet* n) { return n ? n->flops : 0.0; }
int aznet_ready(const AzNet* n) { return n && n->ready; }

// Fold conv(no bias)+BN into tap-major weights [9][cin_pad][cout] and a bias vector.
static void fold_conv3(const float* w, const float* gm, const float* bt, const float* mu, const float* var, int cout, int cin,
                       int cin_pad, std::vector<float>& wf, std::vector<float>& bf) {
  wf.assign((size_t)9 * cin_pad * cout, 0.f);
  bf.assign(cout, 0.f);
  for (int co = 0; co < cout; ++co) {
    const float sc = gm[co] / sqrtf(var[co] + 1e-5f);
    bf[co] = bt[co] - mu[co] * sc;
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < 9; ++t) wf[((size_t)t * cin_pad + ci) * cout + co] = w[((size_t)co * cin + ci) * 9 + t] * sc;
  }
}

int aznet_set_weights(AzNet* n, AzRt& rt, const float* const* T, const int64_t* numel, int nt, std::string& err) {
  const int C = n->C, nb = n->blocks, fc = n->fc, A = n->A;
  const int planes = n->g.planes, HW = n->g.Hc * n->g.Hc;
  const int expect = 21 + 10 * nb;
  if (nt != expect) { err = "expected " + std::to_string(expect) + " tensors (state_dict without num_batches_tracked), got " + std::to_string(nt); return AZ_ERR_BAD_ARG; }
  std::vector<int64_t> want;
  want.push_back((int64_t)C * planes * 9);
  for (int k = 0; k < 4; ++k) want.push_back(C);
  for (int b = 0; b < nb; ++b)
    for (int h = 0; h < 2; ++h) {
      want.push_back((int64_t)C * C * 9);
      for (int k = 0; k < 4; ++k) want.push_back(C);
    }
  want.push_back(2 * C);
  for (int k = 0; k < 4; ++k) want.push_back(2);
  want.push_back((int64_t)A * 2 * HW);
  want.push_back(A);
  want.push_back(C);
  for (int k = 0; k < 4; ++k) want.push_back(1);
  want.push_back((int64_t)fc * HW);
  want.push_back(fc);
  want.push_back(fc);
  want.push_back(1);
  for (int i = 0; i < nt; ++i)
    if (numel[i] != want[i]) {
      err = "tensor " + std::to_string(i) + " has " + std::to_string(numel[i]) + " elements, expected " + std::to_string(want[i]) +
            " (AlphaZeroNet geometry mismatch)";
      return AZ_ERR_BAD_ARG;
    }
  // device buffers are allocated on the first call and refreshed in place afterwards (checkpoint hot-swap path)
  const int n_conv = 1 + 2 * nb;
  if ((int)n->conv_w.size() != n_conv) {
    n->conv_w.assign(n_conv, nullptr);
    n->conv_b.assign(n_conv, nullptr);
  }
  n->host_w.assign(n_conv, std::vector<float>());
  n->host_b.assign(n_conv, std::vector<float>());
  n->layer_src.assign(n_conv, nullptr);
  n->layer_cin.assign(n_conv, 0);
  int ti = 0;
  for (int li = 0; li < n_conv; ++li) {
    n->layer_src[li] = T + ti;   // {conv weight, bn gamma, beta, running_mean, running_var}
    n->layer_cin[li] = li == 0 ? planes : C;
    ti += 5;
  }
  {
    // BatchNorm folding, one task per layer on a few host threads (the fp32 tap-major copy is only built for the parity tower)
    const bool need_f32 = n->precision == AZ_NET_FP32;
    auto work = [&](int t0, int step) {
      for (int li = t0; li < n_conv; li += step) {
        const float* const* L = n->layer_src[li];
        const int cin = n->layer_cin[li], cin_pad = li == 0 ? n->g.cin_pad : C;
        if (need_f32) fold_conv3(L[0], L[1], L[2], L[3], L[4], C, cin, cin_pad, n->host_w[li], n->host_b[li]);
        else {
          n->host_b[li].assign(C, 0.f);
          for (int co = 0; co < C; ++co) n->host_b[li][co] = L[2][co] - L[3][co] * (L[1][co] / sqrtf(L[4][co] + 1e-5f));
        }
      }
    };
    const int nthreads = std::min(8, n_conv);
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t, nthreads);
    work(0, nthreads);
    for (auto& th : pool) th.join();
  }
  for (int li = 0; li < n_conv; ++li) {
    dev_set(rt, n->allocs, n->conv_b[li], n->host_b[li]);
    if (n->precision == AZ_NET_FP32) dev_set(rt, n->allocs, n->conv_w[li], n->host_w[li]);
  }
  // heads: 1x1 conv + BN folded
  std::vector<float> pw(2 * C), pb(2), vw(C), vb(1);
  for (int o = 0; o < 2; ++o) {
    const float sc = T[ti + 1][o] / sqrtf(T[ti + 4][o] + 1e-5f);
    pb[o] = T[ti + 2][o] - T[ti + 3][o] * sc;
    for (int c = 0; c < C; ++c) pw[o * C + c] = T[ti][o * C + c] * sc;
  }
  ti += 5;
  std::vector<float> pfw(T[ti], T[ti] + (size_t)A * 2 * HW), pfb(T[ti + 1], T[ti + 1] + A);
  ti += 2;
  {
    const float sc = T[ti + 1][0] / sqrtf(T[ti + 4][0] + 1e-5f);
    vb[0] = T[ti + 2][0] - T[ti + 3][0] * sc;
    for (int c = 0; c < C; ++c) vw[c] = T[ti][c] * sc;
  }
  ti += 5;
  std::vector<float> v1w(T[ti], T[ti] + (size_t)fc * HW), v1b(T[ti + 1], T[ti + 1] + fc), v2w(T[ti + 2], T[ti + 2] + fc),
      v2b(T[ti + 3], T[ti + 3] + 1);
  std::vector<float> pfwT((size_t)2 * HW * A), v1wT((size_t)HW * fc);
  for (int a2 = 0; a2 < A; ++a2)
    for (int k = 0; k < 2 * HW; ++k) pfwT[(size_t)k * A + a2] = pfw[(size_t)a2 * 2 * HW + k];
  for (int j = 0; j < fc; ++j)
    for (int k = 0; k < HW; ++k) v1wT[(size_t)k * fc + j] = v1w[(size_t)j * HW + k];
  const std::vector<float>* hv[12] = {&pw, &pb, &pfw, &pfb, &pfwT, &v1wT, &vw, &vb, &v1w, &v1b, &v2w, &v2b};
  for (int i = 0; i < 12; ++i) dev_set(rt, n->allocs, n->head_dev[i], *hv[i]);
  HeadParams& hp = n->hp;
  hp.pol_w = n->head_dev[0]; hp.pol_b = n->head_dev[1]; hp.pol_fc_w = n->head_dev[2]; hp.pol_fc_b = n->head_dev[3];
  hp.pol_fc_wT = n->head_dev[4]; hp.val_fc1_wT = n->head_dev[5]; hp.val_w = n->head_dev[6]; hp.val_b = n->head_dev[7];
  hp.val_fc1_w = n->head_dev[8]; hp.val_fc1_b = n->head_dev[9]; hp.val_fc2_w = n->head_dev[10]; hp.val_fc2_b = n->head_dev[11];
  if (!hp.val_fc2_b) { err = "out of device memory"; return AZ_ERR_CUDA; }
  if (n->precision == AZ_NET_BF16) {
    int rc = aznet_tc_set_weights(n, rt, err);
    if (rc) return rc;
  }
  n->ready = 1;
  return AZ_OK;
}

int aznet_forward(AzNet* n, AzRt& rt, const int8_t* obs_base, const int32_t* row_list, const int32_t* n_rows_dev, int max_rows,
                  float* priors_base, float* values_base, int pri_stride) {
  if (!n || !n->ready) return AZ_ERR_STATE;
  if (max_rows > n->max_leaves) max_rows = n->max_leaves;
  const NetGeom& g = n->g;
  if (n->precision == AZ_NET_BF16) return aznet_tc_forward(n, rt, obs_base, row_list, n_rows_dev, max_rows, priors_base, values_base, pri_stride);
  {
    long long work = (long long)max_rows * g.nc;
    int blocks = (int)std::min<long long>((work + 255) / 256, 148 * 8);
    k_net_input<float><<<blocks, 256, 0, rt.stream>>>(obs_base, row_list, n_rows_dev, (float*)n->act_in, g);
    rt.launches++;
  }
  const long long Mmax = (long long)max_rows * g.RP;
  dim3 grid((unsigned)((Mmax + 63) / 64), (unsigned)((n->C + 63) / 64));
  float* X = (float*)n->act_x;
  float* MID = (float*)n->act_mid;
  k_conv_f32<<<grid, 256, 0, rt.stream>>>((const float*)n->act_in, n->conv_w[0], n->conv_b[0], nullptr, X, n_rows_dev, g, g.cin_pad, n->C, 1);
  rt.launches++;
  for (int b = 0; b < n->blocks; ++b) {
    k_conv_f32<<<grid, 256, 0, rt.stream>>>(X, n->conv_w[1 + 2 * b], n->conv_b[1 + 2 * b], nullptr, MID, n_rows_dev, g, n->C, n->C, 1);
    k_conv_f32<<<grid, 256, 0, rt.stream>>>(MID, n->conv_w[2 + 2 * b], n->conv_b[2 + 2 * b], X, X, n_rows_dev, g, n->C, n->C, 1);
    rt.launches += 2;
  }
  launch_heads<float>(rt.stream, X, row_list, n_rows_dev, n->hp, g, n->C, n->A, n->fc, priors_base, values_base, pri_stride, max_rows);
  rt.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_az_error = std::string("network launch: ") + cudaGetErrorString(e); return AZ_ERR_CUDA; }
  return AZ_OK;
}

