// TEST INFRASTRUCTURE ONLY.  The host-emulation build of the engine (tests/emu/build_emu.py) has no
// network: the evaluator is always external (az_search_select / az_search_apply).
#include "az_net.h"
AzNet* aznet_create(const AzDims&, const az_config&, AzRt&, int, std::string& err) { err = "no network in the emulation build"; return nullptr; }
void aznet_destroy(AzNet*) {}
int aznet_set_weights(AzNet*, AzRt&, const float* const*, const int64_t*, int, std::string& err) { err = "no network"; return -7; }
int aznet_forward(AzNet*, AzRt&, const int8_t*, const int32_t*, const int32_t*, int, float*, float*, int) { return -7; }
double aznet_flops_per_eval(const AzNet*) { return 0.0; }
int aznet_ready(const AzNet*) { return 0; }
