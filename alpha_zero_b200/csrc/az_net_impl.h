// az_net_impl.h — internal structures shared by az_net.cu (fp32 tower, heads) and az_net_tc.cu (tcgen05 tower).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "az_net.h"

struct NetGeom {
  int n, nc, planes, obs_bytes;  // observation geometry
  int off;                       // where the N x N observation sits inside the canvas (2 for Gomoku)
  int Hc;                        // canvas side the tower runs on
  int Wr;                        // row stride in cells = Hc + 1 (one zero column)
  int RP;                        // rows per leaf = (Hc+1)*(Hc+1)
  int guard;                     // zero rows before the first and after the last leaf
  int cin_pad;                   // channels of the input feature rows (>= planes)
  int in_dup;                    // 1: input rows carry the planes twice (channels 0.. and 32..), split-bf16 tower
};

struct HeadParams {
  const float *pol_w, *pol_b, *pol_fc_w, *pol_fc_b;
  const float *pol_fc_wT, *val_fc1_wT;  // k-major copies of the FC weights: [2*HW][A], [HW][fc]
  const float *val_w, *val_b, *val_fc1_w, *val_fc1_b, *val_fc2_w, *val_fc2_b;
};

struct AzNetTc;  // tensor-core tower state (az_net_tc.cu)

struct AzNet {
  int blocks = 0, C = 0, C_src = 0, fc = 0, A = 0, precision = 0, max_leaves = 0, ready = 0;  // C: padded channel count the towers run on
  int32_t* dbg_count = nullptr;
  NetGeom g;
  size_t rows_total = 0;
  void *act_in = nullptr, *act_x = nullptr, *act_mid = nullptr;
  std::vector<float*> conv_w, conv_b;                // device, fp32 tower: [9][cin_pad][cout], [cout]
  std::vector<std::vector<float>> host_w, host_b;    // folded host copies (source for the bf16 packing)
  HeadParams hp;
  float* head_dev[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<const float* const*> layer_src;  // per conv layer: {weight, gamma, beta, mean, var} (valid during az_set_weights)
  std::vector<int> layer_cin;
  std::vector<void*> allocs;
  double flops = 0.0;
  AzNetTc* tc = nullptr;
  // CUDA events around the conv launches of the most recent forward (the tower alone: no input / heads kernels), for the roofline
  // (a ring of the last AZ_TOWER_RING forwards: az_last_net_ms reports their mean, one tick alone is a noisy sample)
  static constexpr int kTowerRing = 64;
  cudaEvent_t ev_tower[kTowerRing][2] = {};
  unsigned long long n_forwards = 0;
};

int aznet_tc_create(AzNet* n, AzRt& rt, std::string& err);
void aznet_tc_destroy(AzNet* n);
int aznet_tc_set_weights(AzNet* n, AzRt& rt, std::string& err);
int aznet_tc_forward(AzNet* n, AzRt& rt, const int8_t* obs_base, const int32_t* row_list, const int32_t* n_rows_dev, int max_rows,
                     float* priors_base, float* values_base, int pri_stride);
int aznet_tc_layer(AzNet* n, AzRt& rt, int li, bool with_res, const int32_t* n_rows_dev, int max_rows);
int aznet_tc_mode(const AzNet* n);
static inline bool aznet_is_tc(const AzNet* n) { return n->precision == AZ_NET_BF16 || n->precision == AZ_NET_BF16X3; }
