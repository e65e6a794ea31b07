// az_net.h — policy/value network of the engine (core/network.py:85-173, forward only).
// Two implementations behind one interface: fp32 CUDA-core tower (parity mode) and the tcgen05 bf16
// tower (throughput mode).  Activations live as rows [row = leaf*RP + y*(Wc+1) + x][channels] with one
// zero column / zero row of padding per board, so each 3x3 tap is a constant row offset of the same
// 2-D matrix (implicit GEMM without im2col); Gomoku's pad-3 first layer (network.py:101) is the same
// tower on a 17x17 canvas with the 13x13 observation placed at (2,2).
#pragma once
#include <stdint.h>

#include <string>

#include "az_rt.h"
#include "az_state.h"
#include "../../include/az_engine.h"

struct AzNet;
AzNet* aznet_create(const AzDims& d, const az_config& cfg, AzRt& rt, int max_leaves, std::string& err);
void aznet_destroy(AzNet* n);
int aznet_set_weights(AzNet* n, AzRt& rt, const float* const* tensors, const int64_t* numel, int n_tensors, std::string& err);
// obs_base[row_list[i]] (or obs_base[i] when row_list == nullptr) -> priors_base[row * pri_stride ..], values_base[row]
// for i < *n_rows_dev (device-side count, <= max_rows).
int aznet_forward(AzNet* n, AzRt& rt, const int8_t* obs_base, const int32_t* row_list, const int32_t* n_rows_dev, int max_rows,
                  float* priors_base, float* values_base, int pri_stride);
int aznet_debug_layer(AzNet* n, AzRt& rt, int li, const float* in, const float* res, int cnt, float* out, std::string& err);
int aznet_tc_mode_of(const AzNet* n);
int aznet_padded_filters(const AzNet* n);
double aznet_flops_per_eval(const AzNet* n);
// mean device time of the conv launches of a forward over the most recent (up to 64) forwards, CUDA events on the engine stream
float aznet_last_tower_ms(const AzNet* n, int* n_averaged);
int aznet_ready(const AzNet* n);
