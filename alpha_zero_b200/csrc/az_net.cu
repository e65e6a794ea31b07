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

#include <algorithm>

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
  // the towers run on a channel count padded to their tile granularity (the reference's default Gomoku net has 40 filters,
  // training_gomoku.py:38): padded channels carry zero weights and zero bias, so they stay exactly zero through every layer
  n->C_src = cfg.num_filters;
  n->C = (cfg.num_filters + (cfg.net_precision == AZ_NET_FP32 ? 15 : 63)) / (cfg.net_precision == AZ_NET_FP32 ? 16 : 64) * (cfg.net_precision == AZ_NET_FP32 ? 16 : 64);
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
  g.in_dup = 0;
  if (n->precision != AZ_NET_FP32 && n->precision != AZ_NET_BF16 && n->precision != AZ_NET_BF16X3) { err = "unknown net_precision"; delete n; return nullptr; }
  if (cfg.num_filters < 1 || cfg.num_fc_units < 1 || cfg.num_res_blocks < 0) { err = "bad network dimensions"; delete n; return nullptr; }
  if (d.planes > g.cin_pad) { err = "observation has more than 32 planes"; delete n; return nullptr; }
  n->rows_total = (size_t)max_leaves * g.RP + 2 * (size_t)g.guard + 1024;
  const size_t esz = aznet_is_tc(n) ? 2 : 4;
  const size_t cw = (size_t)(n->precision == AZ_NET_BF16X3 ? 3 * n->C : n->C);  // split rows: [hi | lo | hi]
  n->act_in = rt_alloc(n->rows_total * g.cin_pad * esz);
  n->act_x = rt_alloc(n->rows_total * cw * esz);
  n->act_mid = rt_alloc(n->rows_total * cw * esz);
  n->dbg_count = (int32_t*)rt_alloc(sizeof(int32_t));
  for (int k = 0; k < AzNet::kTowerRing; ++k) { cudaEventCreate(&n->ev_tower[k][0]); cudaEventCreate(&n->ev_tower[k][1]); }
  if (!n->act_in || !n->act_x || !n->act_mid || !n->dbg_count) {
    err = "activation buffers: out of device memory (" + std::to_string((n->rows_total * (g.cin_pad + 2 * cw) * esz) >> 20) + " MB)";
    aznet_destroy(n);
    return nullptr;
  }
  // 2*MAC per evaluation, for the roofline: the network's own channel count, not the padded one
  const double hw = (double)g.Hc * g.Hc, cs = (double)n->C_src;
  n->flops = 2.0 * hw * 9.0 * d.planes * cs + (double)n->blocks * 2.0 * (2.0 * hw * 9.0 * cs * cs) + 2.0 * hw * cs * 3.0 +
             2.0 * (2.0 * hw) * d.A + 2.0 * hw * n->fc + 2.0 * n->fc;
  if (aznet_is_tc(n)) {
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
  rt_free(n->dbg_count);
  for (int k = 0; k < AzNet::kTowerRing; ++k)
    for (int j = 0; j < 2; ++j)
      if (n->ev_tower[k][j]) cudaEventDestroy(n->ev_tower[k][j]);
  for (void* p : n->allocs) rt_free(p);
  delete n;
}

double aznet_flops_per_eval(const AzNet* n) { return n ? n->flops : 0.0; }
float aznet_last_tower_ms(const AzNet* n, int* n_averaged) {
  if (n_averaged) *n_averaged = 0;
  if (!n || !n->n_forwards) return 0.f;
  const unsigned long long last = n->n_forwards - 1;
  if (cudaEventSynchronize(n->ev_tower[last % AzNet::kTowerRing][1]) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  const int cnt = (int)std::min<unsigned long long>(n->n_forwards, (unsigned long long)AzNet::kTowerRing);
  double sum = 0.0;
  int used = 0;
  for (int k = 0; k < cnt; ++k) {
    cudaEvent_t* ev = const_cast<cudaEvent_t*>(n->ev_tower[(last - k) % AzNet::kTowerRing]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) { sum += ms; used++; }
    else cudaGetLastError();
  }
  if (n_averaged) *n_averaged = used;
  return used ? (float)(sum / used) : 0.f;
}
int aznet_padded_filters(const AzNet* n) { return n ? n->C : 0; }
int aznet_tc_mode_of(const AzNet* n) { return n && aznet_is_tc(n) ? aznet_tc_mode(n) : -1; }
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
  const int C = n->C, Cs = n->C_src, nb = n->blocks, fc = n->fc, A = n->A;
  const int planes = n->g.planes, HW = n->g.Hc * n->g.Hc;
  const int expect = 21 + 10 * nb;
  if (nt != expect) { err = "expected " + std::to_string(expect) + " tensors (state_dict without num_batches_tracked), got " + std::to_string(nt); return AZ_ERR_BAD_ARG; }
  std::vector<int64_t> want;
  want.push_back((int64_t)Cs * planes * 9);
  for (int k = 0; k < 4; ++k) want.push_back(Cs);
  for (int b = 0; b < nb; ++b)
    for (int h = 0; h < 2; ++h) {
      want.push_back((int64_t)Cs * Cs * 9);
      for (int k = 0; k < 4; ++k) want.push_back(Cs);
    }
  want.push_back(2 * Cs);
  for (int k = 0; k < 4; ++k) want.push_back(2);
  want.push_back((int64_t)A * 2 * HW);
  want.push_back(A);
  want.push_back(Cs);
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
  // channel padding (C_src -> C): zero weight rows / columns, BN of the padded channels = identity with zero shift
  std::vector<std::vector<float>> padded;
  std::vector<const float*> Tp;
  if (C != Cs) {
    padded.resize(nt);
    Tp.assign(T, T + nt);
    auto pad_conv = [&](int i, int cout_s, int cin_s, int cout_p, int cin_p, int taps) {
      padded[i].assign((size_t)cout_p * cin_p * taps, 0.f);
      for (int co = 0; co < cout_s; ++co)
        for (int ci = 0; ci < cin_s; ++ci)
          memcpy(&padded[i][((size_t)co * cin_p + ci) * taps], T[i] + ((size_t)co * cin_s + ci) * taps, taps * sizeof(float));
      Tp[i] = padded[i].data();
    };
    auto pad_bn = [&](int i0) {  // gamma, beta, running_mean, running_var
      const float fill[4] = {1.f, 0.f, 0.f, 1.f};
      for (int k = 0; k < 4; ++k) {
        padded[i0 + k].assign(C, fill[k]);
        memcpy(padded[i0 + k].data(), T[i0 + k], Cs * sizeof(float));
        Tp[i0 + k] = padded[i0 + k].data();
      }
    };
    int i = 0;
    pad_conv(i, Cs, planes, C, planes, 9);
    pad_bn(i + 1);
    i += 5;
    for (int b = 0; b < 2 * nb; ++b) {
      pad_conv(i, Cs, Cs, C, C, 9);
      pad_bn(i + 1);
      i += 5;
    }
    pad_conv(i, 2, Cs, 2, C, 1);  // policy head 1x1 conv
    i += 7;
    pad_conv(i, 1, Cs, 1, C, 1);  // value head 1x1 conv
    T = Tp.data();
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
  if (aznet_is_tc(n)) {
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
  if (aznet_is_tc(n)) return aznet_tc_forward(n, rt, obs_base, row_list, n_rows_dev, max_rows, priors_base, values_base, pri_stride);
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
  cudaEvent_t* evp = n->ev_tower[n->n_forwards % AzNet::kTowerRing];
  cudaEventRecord(evp[0], rt.stream);
  k_conv_f32<<<grid, 256, 0, rt.stream>>>((const float*)n->act_in, n->conv_w[0], n->conv_b[0], nullptr, X, n_rows_dev, g, g.cin_pad, n->C, 1);
  rt.launches++;
  for (int b = 0; b < n->blocks; ++b) {
    k_conv_f32<<<grid, 256, 0, rt.stream>>>(X, n->conv_w[1 + 2 * b], n->conv_b[1 + 2 * b], nullptr, MID, n_rows_dev, g, n->C, n->C, 1);
    k_conv_f32<<<grid, 256, 0, rt.stream>>>(MID, n->conv_w[2 + 2 * b], n->conv_b[2 + 2 * b], X, X, n_rows_dev, g, n->C, n->C, 1);
    rt.launches += 2;
  }
  cudaEventRecord(evp[1], rt.stream);
  n->n_forwards++;
  launch_heads<float>(rt.stream, X, row_list, n_rows_dev, n->hp, g, n->C, n->C, 0, n->A, n->fc, priors_base, values_base, pri_stride, max_rows);
  rt.launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_az_error = std::string("network launch: ") + cudaGetErrorString(e); return AZ_ERR_CUDA; }
  return AZ_OK;
}


// ---- one conv layer on caller-supplied activations (tests: the benched kernel against a plain convolution) ----------------
// in: float [cnt][cin][Hc][Hc] on the canvas the tower runs on (cin = observation planes for layer 0, num_filters otherwise),
// res: float [cnt][num_filters][Hc][Hc] or null, out: float [cnt][num_filters][Hc][Hc].  Values are converted to the tower's
// storage type on the way in (bf16 rounding; hi + lo for the split tower) and back to float on the way out.
int aznet_debug_layer(AzNet* n, AzRt& rt, int li, const float* in, const float* res, int cnt, float* out, std::string& err) {
  if (!n || !n->ready) { err = "network weights not set"; return AZ_ERR_STATE; }
  const int n_conv = 1 + 2 * n->blocks;
  if (li < 0 || li >= n_conv) { err = "layer index out of range"; return AZ_ERR_BAD_ARG; }
  if (cnt < 1 || cnt > n->max_leaves) { err = "leaf count out of range"; return AZ_ERR_BAD_ARG; }
  if (res && (li < 2 || (li & 1))) { err = "only the second conv of a block takes a residual"; return AZ_ERR_BAD_ARG; }
  const NetGeom& g = n->g;
  const bool tc = aznet_is_tc(n), split = n->precision == AZ_NET_BF16X3;
  const int C = n->C, Cs = n->C_src, Hc = g.Hc;
  const int cin_src = li == 0 ? g.planes : Cs;
  const size_t rows = (size_t)2 * g.guard + (size_t)cnt * g.RP;
  const size_t esz = tc ? 2 : 4;
  const size_t w_in = li == 0 ? (size_t)(tc ? 64 : g.cin_pad) : (size_t)(split ? 3 * C : C);
  const size_t w_act = (size_t)(split ? 3 * C : C);
  auto put = [&](std::vector<unsigned char>& buf, size_t width, size_t row, int c, float v, bool first_layer) {
    if (!tc) { ((float*)buf.data())[row * width + c] = v; return; }
    __nv_bfloat16* b = (__nv_bfloat16*)buf.data() + row * width;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    b[c] = hi;
    if (split) {
      if (first_layer) b[32 + c] = hi;
      else { b[C + c] = __float2bfloat16(v - __bfloat162float(hi)); b[2 * C + c] = hi; }
    }
  };
  auto pack = [&](const float* src, int ch, size_t width, bool first_layer) {
    std::vector<unsigned char> buf(rows * width * esz, 0);
    for (int l = 0; l < cnt; ++l)
      for (int c = 0; c < ch; ++c)
        for (int y = 0; y < Hc; ++y)
          for (int x = 0; x < Hc; ++x)
            put(buf, width, (size_t)g.guard + (size_t)l * g.RP + (size_t)y * g.Wr + x, c, src[(((size_t)l * ch + c) * Hc + y) * Hc + x], first_layer);
    return buf;
  };
  void* d_in = li == 0 ? n->act_in : ((li & 1) ? n->act_x : n->act_mid);
  void* d_out = (li & 1) ? n->act_mid : n->act_x;
  {
    std::vector<unsigned char> b = pack(in, cin_src, w_in, li == 0);
    rt_h2d(rt, d_in, b.data(), b.size());
  }
  if (res) {
    std::vector<unsigned char> b = pack(res, Cs, w_act, false);
    rt_h2d(rt, n->act_x, b.data(), b.size());
  } else if (d_out != d_in) {
    rt_zero(rt, d_out, rows * w_act * esz);
  }
  int32_t c32 = cnt;
  rt_h2d(rt, n->dbg_count, &c32, sizeof(c32));
  if (tc) {
    int rc = aznet_tc_layer(n, rt, li, res != nullptr, n->dbg_count, cnt);
    if (rc) return rc;
  } else {
    const long long Mmax = (long long)cnt * g.RP;
    dim3 grid((unsigned)((Mmax + 63) / 64), (unsigned)((C + 63) / 64));
    k_conv_f32<<<grid, 256, 0, rt.stream>>>((const float*)d_in, n->conv_w[li], n->conv_b[li], res ? (const float*)n->act_x : nullptr, (float*)d_out,
                                            n->dbg_count, g, li == 0 ? g.cin_pad : C, C, 1);
    rt.launches++;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { err = std::string("layer launch: ") + cudaGetErrorString(ce); return AZ_ERR_CUDA; }
  std::vector<unsigned char> ob(rows * w_act * esz);
  rt_d2h(rt, ob.data(), d_out, ob.size());
  if (rt_sync(rt)) { err = g_az_error; return AZ_ERR_CUDA; }
  for (int l = 0; l < cnt; ++l)
    for (int c = 0; c < Cs; ++c)
      for (int y = 0; y < Hc; ++y)
        for (int x = 0; x < Hc; ++x) {
          const size_t row = (size_t)g.guard + (size_t)l * g.RP + (size_t)y * g.Wr + x;
          float v;
          if (!tc) v = ((const float*)ob.data())[row * w_act + c];
          else {
            const __nv_bfloat16* b = (const __nv_bfloat16*)ob.data() + row * w_act;
            v = __bfloat162float(b[c]);
            if (split) v += __bfloat162float(b[C + c]);
          }
          out[(((size_t)l * Cs + c) * Hc + y) * Hc + x] = v;
        }
  // the padding the next layer relies on must have survived: zero board rows / separator cells, and (split) hi copies equal
  for (int l = 0; l < cnt; ++l)
    for (int r = 0; r < g.RP; ++r) {
      const int y = r / g.Wr, x = r - y * g.Wr;
      const bool pad_cell = y >= Hc || x >= Hc;
      const size_t row = (size_t)g.guard + (size_t)l * g.RP + r;
      for (size_t c = 0; c < w_act; ++c) {
        float v = tc ? __bfloat162float(((const __nv_bfloat16*)ob.data())[row * w_act + c]) : ((const float*)ob.data())[row * w_act + c];
        if (pad_cell && v != 0.f) { err = "padding row written: leaf " + std::to_string(l) + " row " + std::to_string(r); return AZ_ERR_STATE; }
        if (!pad_cell && split && c < (size_t)C && v != __bfloat162float(((const __nv_bfloat16*)ob.data())[row * w_act + 2 * C + c])) {
          err = "split row: hi copies differ";
          return AZ_ERR_STATE;
        }
      }
    }
  return AZ_OK;
}
