// TEST INFRASTRUCTURE ONLY.  The host-emulation build of the engine (tests/emu/build_emu.py) has no neural
// network.  So that the device-resident self-play loop (game_advance / game_new / rings / drain) can still be
// exercised without a GPU, this stub stands in for the evaluator with a deterministic hash of the observation:
// priors = normalised pseudo-random weights, value = -0.88 / +0.88 (+- noise) for black / white to move, so that the
// resignation branch is reached.  It is never part of the product library.
#include <math.h>
#include <stdlib.h>

#include "az_net.h"

struct AzNet {
  int A, obs_bytes, ready;
};

AzNet* aznet_create(const AzDims& d, const az_config&, AzRt&, int, std::string&) {
  AzNet* n = new AzNet();
  n->A = d.A;
  n->obs_bytes = d.obs_bytes;
  n->ready = 0;
  return n;
}
void aznet_destroy(AzNet* n) { delete n; }
int aznet_set_weights(AzNet* n, AzRt&, const float* const*, const int64_t*, int, std::string&) {
  n->ready = 1;
  return 0;
}
int aznet_forward(AzNet* n, AzRt&, const int8_t* obs_base, const int32_t* row_list, const int32_t* n_rows_dev, int max_rows,
                  float* priors_base, float* values_base, int pri_stride) {
  const int rows = *n_rows_dev < max_rows ? *n_rows_dev : max_rows;
  // AZ_EMU_SHARP=k raises the pseudo-random prior weights to the k-th power: a peaked policy, hence deep, narrow trees
  const char* sh = getenv("AZ_EMU_SHARP");
  const double sharp = sh ? atof(sh) : 1.0;
  const char* vs = getenv("AZ_EMU_VALUE_SCALE");  // < 1 flattens the values so that the (peaked) priors steer the search
  const double vscale = vs ? atof(vs) : 1.0;
  for (int i = 0; i < rows; ++i) {
    const size_t row = row_list ? (size_t)row_list[i] : (size_t)i;
    const int8_t* o = obs_base + row * n->obs_bytes;
    uint64_t h = 1469598103934665603ull;
    for (int k = 0; k < n->obs_bytes; ++k) { h ^= (uint8_t)o[k]; h *= 1099511628211ull; }
    float* p = priors_base + row * pri_stride;
    double sum = 0.0;
    for (int a = 0; a < n->A; ++a) {
      uint64_t z = h + (uint64_t)(a + 1) * 0x9E3779B97F4A7C15ull;
      z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
      double w = 1.0 + (double)(z % 1000);
      if (sharp != 1.0) w = pow(w / 1000.0, sharp);
      p[a] = (float)w;
      sum += w;
    }
    for (int a = 0; a < n->A; ++a) p[a] = (float)(p[a] / sum);
    // value: the side to move is pessimistic when it is black (last plane = colour to play), optimistic when white, +- hash noise:
    // black's root and best-child values fall below any sensible resignation threshold, which exercises that branch
    const double u = (double)((h >> 16) % 20001) / 10000.0 - 1.0;  // [-1, 1]
    const bool black_to_play = o[n->obs_bytes - 1] != 0;
    values_base[row] = (float)(vscale * ((black_to_play ? -0.88 : 0.88) + 0.1 * u));
  }
  return 0;
}
double aznet_flops_per_eval(const AzNet*) { return 0.0; }
int aznet_ready(const AzNet* n) { return n && n->ready; }
int aznet_debug_layer(AzNet*, AzRt&, int, const float*, const float*, int, float*, std::string& err) {
  err = "the host emulation has no network kernels";
  return AZ_ERR_STATE;
}
int aznet_tc_mode_of(const AzNet*) { return -1; }
int aznet_padded_filters(const AzNet*) { return 0; }
float aznet_last_tower_ms(const AzNet*, int* n) { if (n) *n = 0; return 0.f; }
