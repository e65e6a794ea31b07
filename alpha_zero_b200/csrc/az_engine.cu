// az_engine.cu — host side of the C ABI declared in include/az_engine.h.
//
// Owns the device buffers (structure-of-arrays state of every game slot, az_state.h), builds the
// selection tables, validates arguments the way the reference's Python does, and launches the
// warp-per-game kernels of az_kernels.cuh and the network kernels of az_net*.cu on one stream.
// There is no CPU fallback: without a CUDA device az_create fails.
#include <math.h>

#include <algorithm>
#include <vector>

#include "az_kernels.cuh"
#include "az_net.h"

thread_local std::string g_az_error;

#ifndef AZ_NODE_CACHE_DEFAULT
#define AZ_NODE_CACHE_DEFAULT 1
#endif

struct az_engine {
  az_config cfg;
  AzRt rt;
  AzState E;
  std::vector<void*> allocs;
  size_t alloc_bytes = 0, alloc_failed = 0;  // dev_alloc bookkeeping: bytes held, size of the first failed request
  AzNet* net = nullptr;
  AzNet* net2 = nullptr;  // second weight set (evaluation matches), created by az_set_weights_for(e, 1, ...)
  // staging
  int32_t *d_slots = nullptr, *d_aux = nullptr, *d_out = nullptr;
  float* d_fout = nullptr;
  double* d_dout = nullptr;
  int8_t* d_stage_obs = nullptr;
  float *d_stage_pri = nullptr, *d_stage_val = nullptr;
  double *d_pbc_fresh = nullptr, *d_pbc_f32 = nullptr, *d_sqrt = nullptr;
  double tab_cb = -1, tab_ci = -1;
  std::vector<int32_t> active;  // slots of the current split-phase search, ascending
  int last_total = 0;
  bool selfplay = false;
  unsigned long long drained_games = 0, dropped_samples = 0;
  long long rp_samples_added = 0, rp_games_added = 0;  // UniformReplay.num_samples_added / num_games_added
  int32_t* d_rp_idx = nullptr;
  int8_t* d_rp_obs = nullptr;
  float *d_rp_pi = nullptr, *d_rp_z = nullptr;
  int rp_batch_cap = 0;
  float last_net_ms = 0.f;
  int last_net_evals = 0;
  // az_tick_profile: per-phase device time of the self-play tick, CUDA events on the engine stream
  // AZ_PIPELINE=1 (opt-in): two halves of the slots; the tree kernels of one half run on `tree_rt` while the network evaluates
  // the other half on `rt`
  bool defer_reroot = true;         // AZ_DEFER_REROOT (default 1): k_advance_d + k_reroot_payload instead of the one-warp re-root
  int payload_ctas = 592;
  bool fused_tree = false;          // AZ_FUSED_TREE=1: serial tick with the fused apply -> advance -> collect pass (measured: no gain, off)
  bool pipeline = false;
  int pipeline_mode = 0;            // 1: one launch per phase, 512-CTA grids; 2: persistent fused tree pass, <= one CTA per SM
  unsigned int* d_work_ctr = nullptr;  // [2] game counters of the persistent tree pass, one per half
  int num_sms = 148;
  AzRt tree_rt;
  bool prof_on = false;
  double prof_ms[5] = {0, 0, 0, 0, 0};  // collect+compact, network, apply, advance, whole tick
  int prof_ticks = 0;
#ifndef AZ_EMU
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_net[2] = {nullptr, nullptr}, ev_tree[2] = {nullptr, nullptr}, ev_p0 = nullptr, ev_p1 = nullptr;
  bool collect_occ = false;             // AZ_COLLECT_OCC=1: k_collect built for 7 CTAs per SM (one wave for 4096 games)
  std::vector<cudaEvent_t> prof_ev;     // 5 per tick of the last az_selfplay_tick call
  int prof_pending = 0;                 // ticks recorded and not yet folded into prof_ms
#endif
};

// move / re-root / game end of every slot whose search is complete; the payload of the re-roots is spread over the grid
static void launch_advance(az_engine* e) {
  const AzDims& d = e->E.d;
  if (!e->defer_reroot) { AZ_LAUNCH_WARPS(e->rt, k_advance, d.G, d, e->E); return; }
  rt_zero(e->rt, e->E.rr_count, sizeof(int32_t));
  AZ_LAUNCH_WARPS(e->rt, k_advance_d, d.G, d, e->E);
#ifdef AZ_EMU
  k_reroot_payload(e->E, 0);
#else
  k_reroot_payload<<<e->payload_ctas, AZ_WPB * 32, 0, e->rt.stream>>>(e->E, 0);
#endif
  e->rt.launches++;
}

static void launch_collect(az_engine* e) {
  const AzDims& d = e->E.d;
#ifndef AZ_EMU
  if (d.node_cache) {
    if (e->collect_occ) AZ_LAUNCH_WARPS(e->rt, k_collect_nc_occ, d.G, d, e->E);
    else AZ_LAUNCH_WARPS(e->rt, k_collect_nc, d.G, d, e->E);
    return;
  }
  if (e->collect_occ) { AZ_LAUNCH_WARPS(e->rt, k_collect_occ, d.G, d, e->E); return; }
#endif
  AZ_LAUNCH_WARPS(e->rt, k_collect, d.G, d, e->E);
}

#ifndef AZ_EMU
static void prof_fold(az_engine* e) {
  if (!e->prof_pending) return;
  cudaStreamSynchronize(e->rt.stream);
  for (int t = 0; t < e->prof_pending; ++t) {
    float ms = 0.f;
    for (int k = 0; k < 4; ++k)
      if (cudaEventElapsedTime(&ms, e->prof_ev[(size_t)t * 5 + k], e->prof_ev[(size_t)t * 5 + k + 1]) == cudaSuccess) e->prof_ms[k] += ms;
    if (cudaEventElapsedTime(&ms, e->prof_ev[(size_t)t * 5], e->prof_ev[(size_t)t * 5 + 4]) == cudaSuccess) e->prof_ms[4] += ms;
  }
  e->prof_ticks += e->prof_pending;
  e->prof_pending = 0;
}
#endif

// Every device buffer of the engine comes from here; a failure is remembered (first failing request and the running total) and
// checked once after a group of allocations, so that a partial out-of-memory can never hand a null buffer to a kernel.
template <class T>
static T* dev_alloc(az_engine* e, size_t count) {
  const size_t bytes = count * sizeof(T);
  void* p = rt_alloc(bytes);
  if (p) {
    e->allocs.push_back(p);
    e->alloc_bytes += bytes;
  } else if (!e->alloc_failed) {
    e->alloc_failed = bytes ? bytes : 1;
  }
  return (T*)p;
}

static int alloc_check(az_engine* e, const char* where) {
  if (!e->alloc_failed) return AZ_OK;
  const size_t want = e->alloc_failed;
  e->alloc_failed = 0;
  return az_fail(AZ_ERR_CUDA, std::string(where) + ": device allocation of " + std::to_string(want) + " bytes failed (" +
                                  std::to_string(e->alloc_bytes >> 20) + " MB already held by this engine)");
}

// One engine is bound to one CUDA device, but the calling thread's current device may be anything (another thread of the
// process, torch.cuda.set_device): every entry point makes the engine's device current before it allocates or launches.
// The caller's device is put back when the entry point returns (torch keeps its own notion of the current device per thread).
#ifndef AZ_EMU
struct AzDeviceGuard {
  int prev = -1;
  explicit AzDeviceGuard(int dev) {
    if (dev < 0) return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) {
      prev = cur;
      cudaSetDevice(dev);
    }
  }
  ~AzDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define AZ_ENTER(e) AzDeviceGuard az_device_guard_((e) ? (e)->rt.device : -1)
#else
#define AZ_ENTER(e) do { } while (0)
#endif

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" const char* az_last_error(void) { return g_az_error.c_str(); }
extern "C" int az_version(void) { return 100; }

static void build_tables(az_engine* e, double cb, double ci) {
  if (cb == e->tab_cb && ci == e->tab_ci) return;
  const int n = e->E.d.table_len;
  std::vector<double> fresh(n), f32(n), sq(n);
  for (int i = 0; i < n; ++i) {
    // fresh root: N is a Python float -> double arithmetic (mcts_v2.py:56-62,101)
    fresh[i] = log((1.0 + (double)i + cb) / cb) + ci;
    // np.float32 N: 1 + N, + c_base, / c_base all round to float32 before math.log (numpy>=2 scalar rules)
    volatile float s = 1.0f + (float)i;
    s = s + (float)cb;
    s = s / (float)cb;
    f32[i] = log((double)s) + ci;
    sq[i] = sqrt((double)i);
  }
  rt_h2d(e->rt, e->d_pbc_fresh, fresh.data(), n * sizeof(double));
  rt_h2d(e->rt, e->d_pbc_f32, f32.data(), n * sizeof(double));
  rt_h2d(e->rt, e->d_sqrt, sq.data(), n * sizeof(double));
  e->tab_cb = cb;
  e->tab_ci = ci;
}

extern "C" int az_create(const az_config* cfg, az_engine** out) {
  if (!cfg || !out) return az_fail(AZ_ERR_BAD_ARG, "az_create: null argument");
#ifndef AZ_EMU
  AzDeviceGuard az_device_guard_(cfg->device);  // the engine's device is current while it is built; the caller's comes back afterwards
#endif
  if (cfg->game != AZ_GAME_GO && cfg->game != AZ_GAME_GOMOKU) return az_fail(AZ_ERR_BAD_ARG, "az_create: unknown game");
  if (cfg->board_size < 3 || cfg->board_size > 19) return az_fail(AZ_ERR_BAD_ARG, "az_create: board_size must be in [3, 19]");
  if (cfg->num_stack < 1 || cfg->num_stack > 8) return az_fail(AZ_ERR_BAD_ARG, "az_create: num_stack must be in [1, 8]");
  if (cfg->num_games < 1) return az_fail(AZ_ERR_BAD_ARG, "az_create: num_games must be positive");
  if (cfg->max_simulations < 1 || cfg->max_parallel < 1) return az_fail(AZ_ERR_BAD_ARG, "az_create: max_simulations / max_parallel must be positive");
  az_engine* e = new az_engine();
  e->cfg = *cfg;
  int rc = rt_init(e->rt, cfg->device);
  if (rc) { delete e; return rc; }
  AzDims& d = e->E.d;
  d.game = cfg->game;
  d.n = cfg->board_size;
  d.nc = d.n * d.n;
  d.ncp = round_up(d.nc, 32);
  d.A = d.nc + (cfg->game == AZ_GAME_GO ? 1 : 0);
  d.Ap = round_up(d.A, 32);
  d.num_stack = cfg->num_stack;
  d.planes = 2 * cfg->num_stack + 1;
  d.obs_bytes = d.planes * d.nc;
  d.max_steps = cfg->max_steps > 0 ? cfg->max_steps : 2 * d.nc;
  d.num_to_win = cfg->num_to_win > 0 ? cfg->num_to_win : 5;
  d.komi = cfg->komi;
  d.G = cfg->num_games;
  d.Pmax = cfg->max_parallel;
  d.cap = cfg->max_simulations + 6 * cfg->max_parallel + 16;
  if (d.cap > AZ_CIDX_MASK) { delete e; return az_fail(AZ_ERR_CAPACITY, "az_create: node pool exceeds the 14-bit child links"); }
  d.pass_move = cfg->game == AZ_GAME_GO ? d.nc : -1;
  d.table_len = d.cap + 64;
  d.max_len = (cfg->game == AZ_GAME_GO ? d.max_steps : d.nc) + 1;
  {
    // default: room for every slot finishing a maximum-length game between two drains, twice over (bounded to 8 GB)
    const double per_sample = (double)d.obs_bytes + 4.0 * d.A + 6.0;
    double want = 2.0 * (double)d.G * (double)d.max_len;
    const double cap_bytes = 8.0 * 1024 * 1024 * 1024;
    if (want * per_sample > cap_bytes) want = std::max((double)d.G * d.max_len, cap_bytes / per_sample);
#ifndef AZ_EMU
    {
      // never ask for more than a quarter of what the device has free right now (several actors may share one GPU)
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && want * per_sample > 0.25 * (double)free_b)
        want = std::max((double)d.G * 8.0, 0.25 * (double)free_b / per_sample);
    }
#endif
    d.ring_cap = cfg->sample_ring > 0 ? cfg->sample_ring : (int)std::min(want, 2.0e9);
  }
  e->cfg.max_steps = d.max_steps;
  e->cfg.sample_ring = d.ring_cap;
  AzState& E = e->E;
  const size_t G = d.G, nodes = G * 2 * d.cap, rows = G * d.Pmax;
  E.board = dev_alloc<int8_t>(e, G * d.ncp);
  E.hist = dev_alloc<int8_t>(e, G * 8 * d.ncp);
  E.env_i = dev_alloc<int32_t>(e, G * ENV_INTS);
  E.root_legal = dev_alloc<uint8_t>(e, G * d.Ap);
  E.cN = dev_alloc<float>(e, nodes * d.Ap);
  E.cW = dev_alloc<float>(e, nodes * d.Ap);
  E.cP = dev_alloc<float>(e, nodes * d.Ap);
  E.cidx = dev_alloc<int16_t>(e, nodes * d.Ap);
  E.parent = dev_alloc<int16_t>(e, nodes);
  E.pmove = dev_alloc<int16_t>(e, nodes);
  E.nvloss = dev_alloc<int16_t>(e, nodes);
  E.nto_play = dev_alloc<int8_t>(e, nodes);
  E.expanded = dev_alloc<uint8_t>(e, nodes);
  E.tree_i = dev_alloc<int32_t>(e, G * TREE_INTS);
  E.root_nw = dev_alloc<double>(e, G * 2);
  E.root_p64 = dev_alloc<double>(e, G * d.Ap);
  E.noise = dev_alloc<double>(e, G * d.Ap);
  E.remap = dev_alloc<int16_t>(e, G * d.cap);
  E.rr_jobs = dev_alloc<int32_t>(e, G);
  E.rr_count = dev_alloc<int32_t>(e, 1);
  {
    // per-node position cache (one ply per descent instead of a replay from the root): AZ_NODE_CACHE=0 switches it off
    const char* nc = getenv("AZ_NODE_CACHE");
    d.node_cache = nc ? (atoi(nc) != 0) : AZ_NODE_CACHE_DEFAULT;
    E.nboard = nullptr; E.nlegal = nullptr; E.nko = nullptr;
    if (d.node_cache) {
      E.nboard = dev_alloc<int8_t>(e, nodes * d.ncp);
      E.nlegal = dev_alloc<uint8_t>(e, nodes * d.Ap);
      E.nko = dev_alloc<int16_t>(e, nodes);
    }
  }
  e->d_pbc_fresh = dev_alloc<double>(e, d.table_len);
  e->d_pbc_f32 = dev_alloc<double>(e, d.table_len);
  e->d_sqrt = dev_alloc<double>(e, d.table_len);
  E.pbc_fresh = e->d_pbc_fresh;
  E.pbc_f32 = e->d_pbc_f32;
  E.sqrt_tab = e->d_sqrt;
  E.leaf_node = dev_alloc<int16_t>(e, rows);
  E.leaf_obs = dev_alloc<int8_t>(e, rows * d.obs_bytes);
  E.priors = dev_alloc<float>(e, rows * d.Ap);
  E.values = dev_alloc<float>(e, rows);
  E.leaf_rows = dev_alloc<int32_t>(e, rows);
  E.leaf_rows2 = dev_alloc<int32_t>(e, rows);
  E.slot_net = dev_alloc<uint8_t>(e, G);
  E.leaf_total = dev_alloc<int32_t>(e, 8);  // [0..1] totals, [2] az_net_forward's count, [4..5] totals of the second half (pipeline)
  E.leaf_count = dev_alloc<int32_t>(e, G);
  E.leaf_pk = dev_alloc<int32_t>(e, rows * AZ_PATH);
  E.leaf_pn = dev_alloc<int16_t>(e, rows * AZ_PATH);
  E.leaf_depth = dev_alloc<int32_t>(e, rows);
  E.res_pi = dev_alloc<double>(e, G * d.Ap);
  E.res_q = dev_alloc<double>(e, G * 2);
  E.res_move = dev_alloc<int32_t>(e, G);
  E.g_obs = dev_alloc<int8_t>(e, G * d.max_len * d.obs_bytes + 16);  // + slack: w_copy_bytes reads whole aligned words of the record
  E.g_pi = dev_alloc<float>(e, G * d.max_len * d.A);
  E.g_to_play = dev_alloc<int8_t>(e, G * d.max_len);
  E.g_move = dev_alloc<int16_t>(e, G * d.max_len);
  E.r_obs = dev_alloc<int8_t>(e, (size_t)d.ring_cap * d.obs_bytes);
  E.r_pi = dev_alloc<float>(e, (size_t)d.ring_cap * d.A);
  E.r_z = dev_alloc<float>(e, d.ring_cap);
  E.r_move = dev_alloc<int16_t>(e, d.ring_cap);
  E.games_ring = dev_alloc<int32_t>(e, (size_t)AZ_GAMES_RING * GR_INTS);
  E.counters = dev_alloc<unsigned long long>(e, CT_COUNT);
  e->d_slots = dev_alloc<int32_t>(e, G);
  e->d_aux = dev_alloc<int32_t>(e, G);
  e->d_out = dev_alloc<int32_t>(e, G * 3);
  e->d_fout = dev_alloc<float>(e, 16);
  e->d_dout = dev_alloc<double>(e, 16);
  e->d_stage_obs = dev_alloc<int8_t>(e, rows * d.obs_bytes);
  e->d_stage_pri = dev_alloc<float>(e, rows * d.A);
  e->d_stage_val = dev_alloc<float>(e, rows);
  rc = alloc_check(e, "az_create");
  if (rc) { az_destroy(e); return rc; }
  memset(&E.s, 0, sizeof(E.s));
  E.s.seed = cfg->seed;
  if (cfg->num_filters > 0) {
    std::string err;
    e->net = aznet_create(d, e->cfg, e->rt, (int)rows, err);
    if (!e->net) { az_destroy(e); return az_fail(AZ_ERR_CUDA, "az_create: network: " + err); }
  }
#ifndef AZ_EMU
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);
  {
    // the node-cache kernel needs 108 registers unconstrained (4 CTAs per SM); built for 7 CTAs per SM it spills 16 bytes
    const char* oc = getenv("AZ_COLLECT_OCC");
    e->collect_occ = oc ? atoi(oc) != 0 : e->E.d.node_cache != 0;
  }
#endif
  {
    const char* dr = getenv("AZ_DEFER_REROOT");
    e->defer_reroot = dr ? atoi(dr) != 0 : true;
#ifndef AZ_EMU
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->rt.device);
    e->payload_ctas = sms * 4;
#endif
  }
  {
    const char* ft = getenv("AZ_FUSED_TREE");
    e->fused_tree = ft ? atoi(ft) != 0 : false;
  }
  {
    const char* pl = getenv("AZ_PIPELINE");
    e->pipeline = pl && atoi(pl) != 0 && d.node_cache && d.G >= 2 && e->net;
    e->pipeline_mode = e->pipeline ? atoi(pl) : 0;
    if (e->pipeline) {
      e->d_work_ctr = dev_alloc<unsigned int>(e, 2);
      if (alloc_check(e, "az_create")) { az_destroy(e); return AZ_ERR_CUDA; }
#ifndef AZ_EMU
      cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->rt.device);
#endif
#ifndef AZ_EMU
      bool ok = cudaStreamCreateWithFlags(&e->tree_rt.stream, cudaStreamNonBlocking) == cudaSuccess;
      e->tree_rt.device = e->rt.device;
      for (int h = 0; h < 2 && ok; ++h)
        ok = cudaEventCreateWithFlags(&e->ev_net[h], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&e->ev_tree[h], cudaEventDisableTiming) == cudaSuccess;
      ok = ok && cudaEventCreate(&e->ev_p0) == cudaSuccess && cudaEventCreate(&e->ev_p1) == cudaSuccess;
      if (!ok) { az_destroy(e); return az_fail(AZ_ERR_CUDA, "az_create: pipeline stream / events"); }
#endif
    }
  }
  // every slot starts as a freshly reset game
  std::vector<int32_t> all(G);
  for (size_t i = 0; i < G; ++i) all[i] = (int32_t)i;
  rt_h2d(e->rt, e->d_slots, all.data(), G * sizeof(int32_t));
  AZ_LAUNCH_WARPS(e->rt, k_env_reset, (int)G, d, E, e->d_slots);
  rc = rt_sync(e->rt);
  if (rc) { az_destroy(e); return rc; }
  *out = e;
  return AZ_OK;
}

extern "C" int az_destroy(az_engine* e) {
  AZ_ENTER(e);
  if (!e) return AZ_OK;
  rt_sync(e->rt);
  if (e->net) aznet_destroy(e->net);
  if (e->net2) aznet_destroy(e->net2);
  for (void* p : e->allocs) rt_free(p);
#ifndef AZ_EMU
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
  for (int h = 0; h < 2; ++h) {
    if (e->ev_net[h]) cudaEventDestroy(e->ev_net[h]);
    if (e->ev_tree[h]) cudaEventDestroy(e->ev_tree[h]);
  }
  if (e->ev_p0) cudaEventDestroy(e->ev_p0);
  if (e->ev_p1) cudaEventDestroy(e->ev_p1);
  if (e->tree_rt.stream) { cudaStreamSynchronize(e->tree_rt.stream); cudaStreamDestroy(e->tree_rt.stream); }
#endif
  rt_destroy(e->rt);
  delete e;
  return AZ_OK;
}

extern "C" int az_get_config(az_engine* e, az_config* out) {
  AZ_ENTER(e);
  if (!e || !out) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  *out = e->cfg;
  return AZ_OK;
}
extern "C" int az_num_actions(az_engine* e) { return e ? e->E.d.A : AZ_ERR_BAD_ARG; }
extern "C" int az_obs_bytes(az_engine* e) { return e ? e->E.d.obs_bytes : AZ_ERR_BAD_ARG; }

static int check_slot(az_engine* e, int slot) {
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (slot < 0 || slot >= e->E.d.G) return az_fail(AZ_ERR_BAD_ARG, "slot out of range");
  return AZ_OK;
}

static int upload_slots(az_engine* e, const int32_t* slots, int n) {
  if (!e || !slots || n < 1 || n > e->E.d.G) return az_fail(AZ_ERR_BAD_ARG, "bad slot list");
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= e->E.d.G) return az_fail(AZ_ERR_BAD_ARG, "slot out of range");
  rt_h2d(e->rt, e->d_slots, slots, n * sizeof(int32_t));
  return AZ_OK;
}

// ---- weights / network ---------------------------------------------------------------------------
extern "C" int az_set_weights(az_engine* e, const float* const* tensors, const int64_t* numel, int32_t n_tensors) {
  AZ_ENTER(e);
  if (!e || !tensors || !numel) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (!e->net) return az_fail(AZ_ERR_STATE, "engine was created without a network (num_filters == 0)");
  std::string err;
  int rc = aznet_set_weights(e->net, e->rt, tensors, numel, n_tensors, err);
  if (rc) return az_fail(rc, "az_set_weights: " + err);
  return rt_sync(e->rt);
}

extern "C" int az_set_weights_for(az_engine* e, int32_t which, const float* const* tensors, const int64_t* numel, int32_t n_tensors) {
  AZ_ENTER(e);
  if (!e || !tensors || !numel) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (which == 0) return az_set_weights(e, tensors, numel, n_tensors);
  if (which != 1) return az_fail(AZ_ERR_BAD_ARG, "az_set_weights_for: weight set must be 0 or 1");
  if (!e->net) return az_fail(AZ_ERR_STATE, "engine was created without a network (num_filters == 0)");
  std::string err;
  if (!e->net2) {
    e->net2 = aznet_create(e->E.d, e->cfg, e->rt, e->E.d.G * e->E.d.Pmax, err);
    if (!e->net2) return az_fail(AZ_ERR_CUDA, "az_set_weights_for: second network: " + err);
  }
  int rc = aznet_set_weights(e->net2, e->rt, tensors, numel, n_tensors, err);
  if (rc) return az_fail(rc, "az_set_weights_for: " + err);
  return rt_sync(e->rt);
}

extern "C" int az_net_forward(az_engine* e, const int8_t* obs, int32_t n, float* priors, float* values) {
  AZ_ENTER(e);
  if (!e || !obs || !priors || !values) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (!e->net || !aznet_ready(e->net)) return az_fail(AZ_ERR_STATE, "network weights not set");
  const AzDims& d = e->E.d;
  const int cap = d.G * d.Pmax;
  for (int off = 0; off < n; off += cap) {
    const int m = std::min(cap, n - off);
    rt_h2d(e->rt, e->d_stage_obs, obs + (size_t)off * d.obs_bytes, (size_t)m * d.obs_bytes);
    int32_t cnt = m;
    rt_h2d(e->rt, e->E.leaf_total + 2, &cnt, sizeof(cnt));
    int rc = aznet_forward(e->net, e->rt, e->d_stage_obs, nullptr, e->E.leaf_total + 2, m, e->E.priors, e->E.values, d.Ap);
    if (rc) return az_fail(rc, "az_net_forward: " + g_az_error);
    std::vector<float> tmp((size_t)m * d.Ap);
    rt_d2h(e->rt, tmp.data(), e->E.priors, tmp.size() * sizeof(float));
    for (int i = 0; i < m; ++i) memcpy(priors + (size_t)(off + i) * d.A, tmp.data() + (size_t)i * d.Ap, d.A * sizeof(float));
    rt_d2h(e->rt, values + off, e->E.values, m * sizeof(float));
  }
  return rt_sync(e->rt);
}

extern "C" int az_net_conv_layer(az_engine* e, int32_t layer, const float* in, const float* res, int32_t n, float* out) {
  AZ_ENTER(e);
  if (!e || !in || !out) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (!e->net || !aznet_ready(e->net)) return az_fail(AZ_ERR_STATE, "network weights not set");
  std::string err;
  int rc = aznet_debug_layer(e->net, e->rt, layer, in, res, n, out, err);
  if (rc) return az_fail(rc, "az_net_conv_layer: " + err);
  return AZ_OK;
}

extern "C" int az_net_info(az_engine* e, int32_t* tc_mode, int32_t* padded_filters, double* flops_per_eval) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (!e->net) return az_fail(AZ_ERR_STATE, "engine was created without a network (num_filters == 0)");
  if (tc_mode) *tc_mode = aznet_tc_mode_of(e->net);
  if (padded_filters) *padded_filters = aznet_padded_filters(e->net);
  if (flops_per_eval) *flops_per_eval = aznet_flops_per_eval(e->net);
  return AZ_OK;
}

// ---- BoardGameEnv ----------------------------------------------------------------------------------
extern "C" int az_env_reset(az_engine* e, const int32_t* slots, int32_t n) {
  AZ_ENTER(e);
  int rc = upload_slots(e, slots, n);
  if (rc) return rc;
  AZ_LAUNCH_WARPS(e->rt, k_env_reset, n, e->E.d, e->E, e->d_slots);
  return rt_sync(e->rt);
}

extern "C" int az_env_step(az_engine* e, const int32_t* slots, const int32_t* actions, int32_t n, float* rewards, int32_t* dones) {
  AZ_ENTER(e);
  int rc = upload_slots(e, slots, n);
  if (rc) return rc;
  if (!actions) return az_fail(AZ_ERR_BAD_ARG, "null actions");
  rt_h2d(e->rt, e->d_aux, actions, n * sizeof(int32_t));
  AZ_LAUNCH_WARPS(e->rt, k_env_step, n, e->E.d, e->E, e->d_slots, e->d_aux, e->d_out);
  std::vector<int32_t> out(n * 3);
  rt_d2h(e->rt, out.data(), e->d_out, out.size() * sizeof(int32_t));
  rc = rt_sync(e->rt);
  if (rc) return rc;
  int first_err = 0, err_i = -1;
  for (int i = 0; i < n; ++i) {
    if (rewards) rewards[i] = 0.5f * (float)out[i * 3 + 1];
    if (dones) dones[i] = out[i * 3 + 2];
    if (out[i * 3] && !first_err) { first_err = out[i * 3]; err_i = i; }
  }
  if (first_err == AZ_ERR_GAME_OVER) return az_fail(first_err, "Game is over, call reset before using step method.");
  if (first_err == AZ_ERR_INVALID_ACTION) return az_fail(first_err, "Invalid action. The action " + std::to_string(actions[err_i]) + " is out of bound.");
  if (first_err == AZ_ERR_ILLEGAL_ACTION) return az_fail(first_err, "Illegal action " + std::to_string(actions[err_i]) + ".");
  return AZ_OK;
}

extern "C" int az_env_replay(az_engine* e, const int32_t* slots, int32_t n, const int16_t* moves, const int32_t* offsets, int8_t* states,
                             int32_t* n_played, int32_t* status) {
  AZ_ENTER(e);
  int rc = upload_slots(e, slots, n);
  if (rc) return rc;
  if (!moves || !offsets || !n_played || !status) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (offsets[0] != 0) return az_fail(AZ_ERR_BAD_ARG, "offsets[0] must be 0");
  std::vector<char> seen((size_t)e->E.d.G, 0);
  for (int i = 0; i < n; ++i) {
    if (offsets[i + 1] < offsets[i]) return az_fail(AZ_ERR_BAD_ARG, "offsets must be non-decreasing");
    if (seen[slots[i]]++) return az_fail(AZ_ERR_BAD_ARG, "duplicate slot in replay list");
  }
  const size_t total = (size_t)offsets[n], ob = (size_t)e->E.d.obs_bytes;
  int16_t* d_moves = (int16_t*)rt_alloc((total + 1) * sizeof(int16_t));
  int32_t* d_off = (int32_t*)rt_alloc((size_t)(n + 1) * sizeof(int32_t));
  int32_t* d_res = (int32_t*)rt_alloc((size_t)n * 2 * sizeof(int32_t));
  int8_t* d_states = states && total ? (int8_t*)rt_alloc(total * ob) : nullptr;
  if (!d_moves || !d_off || !d_res || (states && total && !d_states)) {
    rt_free(d_moves); rt_free(d_off); rt_free(d_res); rt_free(d_states);
    return az_fail(AZ_ERR_CUDA, "az_env_replay: out of device memory for " + std::to_string(total) + " positions");
  }
  if (total) rt_h2d(e->rt, d_moves, moves, total * sizeof(int16_t));
  rt_h2d(e->rt, d_off, offsets, (size_t)(n + 1) * sizeof(int32_t));
  AZ_LAUNCH_WARPS(e->rt, k_env_replay, n, e->E.d, e->E, e->d_slots, d_moves, d_off, d_states, d_res);
  std::vector<int32_t> res((size_t)n * 2);
  rt_d2h(e->rt, res.data(), d_res, res.size() * sizeof(int32_t));
  if (d_states) rt_d2h(e->rt, states, d_states, total * ob);  // rows of games that stopped early stay zero past n_played
  rc = rt_sync(e->rt);
  rt_free(d_moves); rt_free(d_off); rt_free(d_res); rt_free(d_states);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) { n_played[i] = res[i * 2]; status[i] = res[i * 2 + 1]; }
  return AZ_OK;
}

extern "C" int az_env_observation(az_engine* e, int32_t slot, int8_t* out) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  AZ_LAUNCH_WARPS(e->rt, k_env_obs, 1, e->E.d, e->E, slot, e->d_stage_obs);
  rt_d2h(e->rt, out, e->d_stage_obs, e->E.d.obs_bytes);
  return rt_sync(e->rt);
}

extern "C" int az_env_legal_actions(az_engine* e, int32_t slot, uint8_t* out) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  rt_d2h(e->rt, out, e->E.root_legal + (size_t)slot * e->E.d.Ap, e->E.d.A);
  return AZ_OK;
}

extern "C" int az_env_board(az_engine* e, int32_t slot, int8_t* out) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  const AzDims& d = e->E.d;
  rt_d2h(e->rt, out, e->E.board + (size_t)slot * d.ncp, d.nc);
  if (d.game == AZ_GAME_GOMOKU)
    for (int i = 0; i < d.nc; ++i)
      if (out[i] == -1) out[i] = 2;  // reference ids (envs/base.py:33-34)
  return AZ_OK;
}

static int ext_player(const AzDims& d, int v) { return (d.game == AZ_GAME_GOMOKU && v == -1) ? 2 : v; }

extern "C" int az_env_scalars(az_engine* e, int32_t slot, int32_t* out) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  int32_t ei[ENV_INTS];
  rt_d2h(e->rt, ei, e->E.env_i + (size_t)slot * ENV_INTS, sizeof(ei));
  const AzDims& d = e->E.d;
  out[0] = ext_player(d, ei[EI_TO_PLAY]);
  out[1] = ei[EI_STEPS];
  out[2] = ei[EI_LAST_MOVE];
  out[3] = ext_player(d, ei[EI_LAST_PLAYER]);
  out[4] = ext_player(d, ei[EI_WINNER]);
  out[5] = ei[EI_DONE];
  out[6] = ei[EI_KO];
  out[7] = ei[EI_BY_RESIGN];
  out[8] = ei[EI_CAPS_B];
  out[9] = ei[EI_CAPS_W];
  out[10] = ei[EI_NUM_PASSES];
  return AZ_OK;
}

extern "C" int az_env_score(az_engine* e, int32_t slot, float* out_score) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  AZ_LAUNCH_WARPS(e->rt, k_env_score, 1, e->E.d, e->E, slot, e->d_fout);
  rt_d2h(e->rt, out_score, e->d_fout, sizeof(float));
  return rt_sync(e->rt);
}

extern "C" int az_env_copy(az_engine* e, int32_t src, int32_t dst) {
  AZ_ENTER(e);
  int rc = check_slot(e, src);
  if (!rc) rc = check_slot(e, dst);
  if (rc) return rc;
  if (src == dst) return AZ_OK;
  AZ_LAUNCH_WARPS(e->rt, k_env_copy, 1, e->E.d, e->E, src, dst);
  return rt_sync(e->rt);
}

extern "C" int az_env_state_bytes(az_engine* e) {
  AZ_ENTER(e);
  if (!e) return AZ_ERR_BAD_ARG;
  const AzDims& d = e->E.d;
  return d.ncp * 9 + ENV_INTS * 4 + d.Ap;
}

extern "C" int az_env_export(az_engine* e, int32_t slot, uint8_t* out) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  const AzDims& d = e->E.d;
  rt_d2h(e->rt, out, e->E.board + (size_t)slot * d.ncp, d.ncp);
  rt_d2h(e->rt, out + d.ncp, e->E.hist + (size_t)slot * 8 * d.ncp, 8 * d.ncp);
  rt_d2h(e->rt, out + 9 * d.ncp, e->E.env_i + (size_t)slot * ENV_INTS, ENV_INTS * 4);
  rt_d2h(e->rt, out + 9 * d.ncp + ENV_INTS * 4, e->E.root_legal + (size_t)slot * d.Ap, d.Ap);
  return AZ_OK;
}

extern "C" int az_env_import(az_engine* e, int32_t slot, const uint8_t* in) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  const AzDims& d = e->E.d;
  rt_h2d(e->rt, e->E.board + (size_t)slot * d.ncp, in, d.ncp);
  rt_h2d(e->rt, e->E.hist + (size_t)slot * 8 * d.ncp, in + d.ncp, 8 * d.ncp);
  rt_h2d(e->rt, e->E.env_i + (size_t)slot * ENV_INTS, in + 9 * d.ncp, ENV_INTS * 4);
  rt_h2d(e->rt, e->E.root_legal + (size_t)slot * d.Ap, in + 9 * d.ncp + ENV_INTS * 4, d.Ap);
  int32_t ti[TREE_INTS] = {0};
  rt_h2d(e->rt, e->E.tree_i + (size_t)slot * TREE_INTS, ti, sizeof(ti));
  return AZ_OK;
}

// ---- search ------------------------------------------------------------------------------------------
static int set_search_cfg(az_engine* e, const az_search_params* p) {
  if (!p) return az_fail(AZ_ERR_BAD_ARG, "null search params");
  if (p->num_simulations < 1) return az_fail(AZ_ERR_BAD_ARG, "Expect `num_simulations` to a positive integer, got " + std::to_string(p->num_simulations));
  const int P = p->num_parallel > 1 ? p->num_parallel : 1;
  const AzDims& d = e->E.d;
  if (P > d.Pmax) return az_fail(AZ_ERR_CAPACITY, "num_parallel exceeds the engine's max_parallel");
  const int sims_bound = p->num_simulations + (P > 1 ? P : 0);
  if (sims_bound + 4 * P + 16 > d.cap) return az_fail(AZ_ERR_CAPACITY, "num_simulations exceeds the engine's node pool (max_simulations)");
  if (!(p->c_puct_base > 0.0)) return az_fail(AZ_ERR_BAD_ARG, "c_puct_base must be positive");
  AzSearchCfg& s = e->E.s;  // everything validated: commit (a rejected call leaves the running configuration untouched)
  s.P = P;
  s.use_vloss = P > 1;
  s.tries = P > 1 ? 2 * P : 1;
  s.sims_bound = sims_bound;
  s.root_noise = p->root_noise;
  s.deterministic = p->deterministic;
  build_tables(e, p->c_puct_base, p->c_puct_init);
  return AZ_OK;
}

extern "C" int az_search_begin(az_engine* e, const int32_t* slots, const int32_t* reuse, int32_t n, const az_search_params* p,
                               int32_t warm_up, const double* noise) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (!slots || n < 1 || n > e->E.d.G) return az_fail(AZ_ERR_BAD_ARG, "bad slot list");
  for (int i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= e->E.d.G) return az_fail(AZ_ERR_BAD_ARG, "slot out of range");
  for (int i = 1; i < n; ++i)
    if (slots[i] <= slots[i - 1]) return az_fail(AZ_ERR_BAD_ARG, "az_search_begin: slots must be strictly ascending");
  int rc = set_search_cfg(e, p);
  if (rc) return rc;
  rc = upload_slots(e, slots, n);
  if (rc) return rc;
  const AzDims& d = e->E.d;
  // RuntimeError('Game is over.') (mcts_v2.py:360)
  for (int i = 0; i < n; ++i) {
    int32_t done = 0;
    rt_d2h(e->rt, &done, e->E.env_i + (size_t)slots[i] * ENV_INTS + EI_DONE, sizeof(done));
    if (done) return az_fail(AZ_ERR_GAME_OVER, "Game is over.");
  }
  e->E.s.selfplay = 0;
  e->selfplay = false;
  e->E.s.host_noise = (p->root_noise && noise) ? 1 : 0;
  if (e->E.s.host_noise) {
    std::vector<double> tmp((size_t)d.Ap, 0.0);
    for (int i = 0; i < n; ++i) {
      std::copy(noise + (size_t)i * d.A, noise + (size_t)(i + 1) * d.A, tmp.begin());
      rt_h2d(e->rt, e->E.noise + (size_t)slots[i] * d.Ap, tmp.data(), d.Ap * sizeof(double));
    }
  }
  std::vector<int32_t> ru(n, 0);
  if (reuse) ru.assign(reuse, reuse + n);
  rt_h2d(e->rt, e->d_aux, ru.data(), n * sizeof(int32_t));
  AZ_LAUNCH_THREADS(e->rt, k_clear_active, d.G, e->E);
  AZ_LAUNCH_WARPS(e->rt, k_search_begin, n, d, e->E, e->d_slots, e->d_aux, warm_up ? 1 : 0);
  e->active.assign(slots, slots + n);
  e->last_total = 0;
  return rt_sync(e->rt);
}

static int collect_and_compact(az_engine* e, int32_t tot[2]) {
  const AzDims& d = e->E.d;
  launch_collect(e);
#ifdef AZ_EMU
  e->rt.launches++;
  k_compact(e->E, d.G);
#else
  e->rt.launches++;
  k_compact<<<1, 1024, 0, e->rt.stream>>>(e->E, d.G);
#endif
  rt_d2h(e->rt, tot, e->E.leaf_total, 2 * sizeof(int32_t));
  return rt_sync(e->rt);
}

extern "C" int az_search_select(az_engine* e, int8_t* leaf_obs, int32_t* counts, int32_t* n_leaves, int32_t* n_active) {
  AZ_ENTER(e);
  if (!e || !n_leaves || !n_active) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (e->active.empty()) return az_fail(AZ_ERR_STATE, "az_search_select: no search in progress");
  const AzDims& d = e->E.d;
  int32_t tot[2] = {0, 0};
  int rc = collect_and_compact(e, tot);
  if (rc) return rc;
  e->last_total = tot[0];
  *n_leaves = tot[0];
  *n_active = tot[1];
  if (tot[0] > 0 && leaf_obs) {
    AZ_LAUNCH_THREADS(e->rt, k_gather_obs, (long long)tot[0] * d.obs_bytes, e->E, e->d_stage_obs);
    rt_d2h(e->rt, leaf_obs, e->d_stage_obs, (size_t)tot[0] * d.obs_bytes);
  }
  if (counts) {
    std::vector<int32_t> all(d.G);
    rt_d2h(e->rt, all.data(), e->E.leaf_count, d.G * sizeof(int32_t));
    for (size_t i = 0; i < e->active.size(); ++i) counts[i] = all[e->active[i]];
  }
  return rt_sync(e->rt);
}

extern "C" int az_search_apply(az_engine* e, const float* priors, const float* values, int32_t n_leaves) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (n_leaves != e->last_total) return az_fail(AZ_ERR_BAD_ARG, "az_search_apply: leaf count does not match the last select");
  const AzDims& d = e->E.d;
  if (n_leaves > 0) {
    if (!priors || !values) return az_fail(AZ_ERR_BAD_ARG, "null priors / values");
    rt_h2d(e->rt, e->d_stage_pri, priors, (size_t)n_leaves * d.A * sizeof(float));
    rt_h2d(e->rt, e->d_stage_val, values, (size_t)n_leaves * sizeof(float));
    AZ_LAUNCH_THREADS(e->rt, k_scatter_eval, (long long)n_leaves * d.A, e->E, e->d_stage_pri, e->d_stage_val);
  }
  AZ_LAUNCH_WARPS(e->rt, k_apply, d.G, d, e->E);
  e->last_total = 0;
  return rt_sync(e->rt);
}

extern "C" int az_search_run(az_engine* e) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (!e->net || !aznet_ready(e->net)) return az_fail(AZ_ERR_STATE, "az_search_run: network weights not set");
  const AzDims& d = e->E.d;
  for (int it = 0; it < 1000000; ++it) {
    int32_t tot[2];
    int rc = collect_and_compact(e, tot);
    if (rc) return rc;
    if (tot[1] == 0) break;
    if (tot[0] > 0) {
      rc = aznet_forward(e->net, e->rt, e->E.leaf_obs, e->E.leaf_rows, e->E.leaf_total, d.G * d.Pmax, e->E.priors, e->E.values, d.Ap);
      if (rc) return az_fail(rc, "network forward: " + g_az_error);
    }
    AZ_LAUNCH_WARPS(e->rt, k_apply, d.G, d, e->E);
  }
  return rt_sync(e->rt);
}

extern "C" int az_search_result(az_engine* e, int32_t slot, float* child_N, float* child_W, double* pi, double* root_q,
                                int32_t* argmax_move) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  const AzDims& d = e->E.d;
  int32_t ti[TREE_INTS];
  rt_d2h(e->rt, ti, e->E.tree_i + (size_t)slot * TREE_INTS, sizeof(ti));
  if (ti[TI_STATE] != ST_DONE) return az_fail(AZ_ERR_STATE, "az_search_result: search not finished");
  const size_t row0 = ((size_t)slot * 2 + ti[TI_BUF]) * d.cap * d.Ap;
  if (child_N) rt_d2h(e->rt, child_N, e->E.cN + row0, d.A * sizeof(float));
  if (child_W) rt_d2h(e->rt, child_W, e->E.cW + row0, d.A * sizeof(float));
  if (pi) rt_d2h(e->rt, pi, e->E.res_pi + (size_t)slot * d.Ap, d.A * sizeof(double));
  if (root_q) rt_d2h(e->rt, root_q, e->E.res_q + (size_t)slot * 2, sizeof(double));
  if (argmax_move) rt_d2h(e->rt, argmax_move, e->E.res_move + slot, sizeof(int32_t));
  return AZ_OK;
}

extern "C" int az_search_commit(az_engine* e, int32_t slot, int32_t move, double* best_child_q, int32_t* has_next) {
  AZ_ENTER(e);
  int rc = check_slot(e, slot);
  if (rc) return rc;
  if (move < 0 || move >= e->E.d.A) return az_fail(AZ_ERR_BAD_ARG, "az_search_commit: move out of range");
  AZ_LAUNCH_WARPS(e->rt, k_commit, 1, e->E.d, e->E, slot, move, e->d_dout);
  double out[2];
  rt_d2h(e->rt, out, e->d_dout, sizeof(out));
  if (best_child_q) *best_child_q = out[0];
  if (has_next) *has_next = (int32_t)out[1];
  return rt_sync(e->rt);
}

// ---- device-resident self-play -------------------------------------------------------------------------
extern "C" int az_selfplay_begin(az_engine* e, const az_selfplay_params* p) {
  AZ_ENTER(e);
  if (!e || !p) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (!e->net || !aznet_ready(e->net)) return az_fail(AZ_ERR_STATE, "az_selfplay_begin: network weights not set");
  int rc = set_search_cfg(e, &p->search);
  if (rc) return rc;
  AzSearchCfg& s = e->E.s;
  s.selfplay = 1;
  s.host_noise = 0;
  s.match = 0;
  s.warm_up_steps = p->warm_up_steps;
  s.check_resign_after = p->check_resign_after_steps;
  s.resign_threshold = p->resign_threshold;
  s.disable_resign_ratio = p->disable_resign_ratio;
  e->selfplay = true;
  e->active.clear();
  rt_zero(e->rt, e->E.counters, CT_COUNT * sizeof(unsigned long long));
  e->drained_games = 0;
  e->dropped_samples = 0;
  AZ_LAUNCH_WARPS(e->rt, k_selfplay_begin, e->E.d.G, e->E.d, e->E);
  return rt_sync(e->rt);
}

extern "C" int az_selfplay_update(az_engine* e, const az_selfplay_params* p) {
  AZ_ENTER(e);
  if (!e || !p) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (!e->selfplay) return az_fail(AZ_ERR_STATE, "az_selfplay_update: call az_selfplay_begin first");
  if (p->search.c_puct_base > 0.0) {  // new search parameters from the next leaf batch on; searches in flight run to the new bound
    int rc = rt_sync(e->rt);          // the tables / AzSearchCfg travel with the next launches; nothing of the old ones may be in flight
    if (!rc) rc = set_search_cfg(e, &p->search);
    if (rc) return rc;
  }
  AzSearchCfg& s = e->E.s;  // games in flight keep running; only the per-move / per-new-game policy knobs change
  s.warm_up_steps = p->warm_up_steps;
  s.check_resign_after = p->check_resign_after_steps;
  s.resign_threshold = p->resign_threshold;
  s.disable_resign_ratio = p->disable_resign_ratio;
  return AZ_OK;
}

extern "C" int az_selfplay_restart(az_engine* e, const int32_t* slots, int32_t n) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (!e->selfplay) return az_fail(AZ_ERR_STATE, "az_selfplay_restart: call az_selfplay_begin first");
  if (n == 0) return AZ_OK;
  int rc = upload_slots(e, slots, n);
  if (rc) return rc;
  std::vector<char> seen((size_t)e->E.d.G, 0);
  for (int i = 0; i < n; ++i)
    if (seen[slots[i]]++) return az_fail(AZ_ERR_BAD_ARG, "az_selfplay_restart: duplicate slot");
  AZ_LAUNCH_WARPS(e->rt, k_selfplay_restart, n, e->E.d, e->E, e->d_slots);
  return rt_sync(e->rt);
}

// Two-half software pipeline of the tick (AZ_PIPELINE=1).  Games are independent, so the slots are split into halves A and B;
// per leaf batch:   engine stream:  net(A)            net(B)            net(A) ...
//                   tree stream:           tree(A)           tree(B)          ...      tree(h) = apply, advance, collect, compact
// tree(A) of batch t runs while the network evaluates B's leaves of batch t, tree(B) while it evaluates A's of batch t+1: the
// 20 % of the tick the tree kernels take disappears behind the tensor-core kernels (which leave exactly one 72-register CTA of
// room per SM, see az_kernels.cuh).  One call still runs n complete leaf batches for every game and leaves every slot in the
// "after advance" state, like the serial tick: the first collect and the last apply / advance of a call are not overlapped.
static int selfplay_tick_pipelined(az_engine* e, int n_ticks) {
  const AzDims& d = e->E.d;
  const int g0[2] = {0, d.G / 2}, ng[2] = {d.G / 2, d.G - d.G / 2};
  int32_t* tot[2] = {e->E.leaf_total, e->E.leaf_total + 4};
  const bool fused = e->pipeline_mode >= 2;
  // tree pass of half h: the requested phases (AZ_PH_*), then - when leaves were collected - the compaction of its leaf rows
  auto tree_pass = [&](int h, int phases) {
    if (fused) {
      rt_zero(e->tree_rt, e->d_work_ctr + h, sizeof(unsigned int));
#ifdef AZ_EMU
      k_tree_p(e->E, g0[h], ng[h], phases, e->d_work_ctr + h);
#else
      const int grid = std::min(e->num_sms, (ng[h] + AZ_WPB - 1) / AZ_WPB);
      k_tree_p<<<grid, AZ_WPB * 32, AZ_WPB * az_sim_stride(d), e->tree_rt.stream>>>(e->E, g0[h], ng[h], phases, e->d_work_ctr + h);
#endif
      e->tree_rt.launches++;
    } else {
      if (phases & AZ_PH_APPLY) AZ_LAUNCH_WARPS(e->tree_rt, k_apply_r, ng[h], d, e->E, g0[h]);
      if (phases & AZ_PH_ADVANCE) AZ_LAUNCH_WARPS(e->tree_rt, k_advance_r, ng[h], d, e->E, g0[h]);
      if (phases & AZ_PH_COLLECT) AZ_LAUNCH_WARPS(e->tree_rt, k_collect_r, ng[h], d, e->E, g0[h]);
    }
    if (phases & AZ_PH_COLLECT) {
#ifdef AZ_EMU
      k_compact_r(e->E, g0[h], ng[h], tot[h]);
#else
      k_compact_r<<<1, 128, 0, e->tree_rt.stream>>>(e->E, g0[h], ng[h], tot[h]);
#endif
      e->tree_rt.launches++;
    }
#ifndef AZ_EMU
    cudaEventRecord(e->ev_tree[h], e->tree_rt.stream);
#endif
  };
  auto collect = [&](int h) { tree_pass(h, AZ_PH_COLLECT); };
#ifndef AZ_EMU
  // everything queued on the engine stream so far (weights, earlier calls) happens before the tree stream starts
  cudaEventRecord(e->ev_p0, e->rt.stream);
  cudaStreamWaitEvent(e->tree_rt.stream, e->ev_p0, 0);
#endif
  collect(0);
  collect(1);
  for (int t = 0; t < n_ticks; ++t) {
    for (int h = 0; h < 2; ++h) {
#ifndef AZ_EMU
      cudaStreamWaitEvent(e->rt.stream, e->ev_tree[h], 0);
      if (t == n_ticks - 1 && h == 1) cudaEventRecord(e->ev0, e->rt.stream);
#endif
      int rc = aznet_forward(e->net, e->rt, e->E.leaf_obs, e->E.leaf_rows + (size_t)g0[h] * d.Pmax, tot[h], ng[h] * d.Pmax, e->E.priors,
                             e->E.values, d.Ap);
      if (rc) return az_fail(rc, "network forward: " + g_az_error);
#ifndef AZ_EMU
      if (t == n_ticks - 1 && h == 1) cudaEventRecord(e->ev1, e->rt.stream);
      cudaEventRecord(e->ev_net[h], e->rt.stream);
      cudaStreamWaitEvent(e->tree_rt.stream, e->ev_net[h], 0);
#endif
      tree_pass(h, AZ_PH_APPLY | AZ_PH_ADVANCE | (t < n_ticks - 1 ? AZ_PH_COLLECT : 0));
    }
  }
#ifndef AZ_EMU
  // join: whatever follows on the engine stream (drain, counters, the next call) sees both halves finished
  cudaStreamWaitEvent(e->rt.stream, e->ev_tree[0], 0);
  cudaStreamWaitEvent(e->rt.stream, e->ev_tree[1], 0);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return az_fail(AZ_ERR_CUDA, std::string("CUDA launch error: ") + cudaGetErrorString(ce));
#endif
  e->rt.launches += e->tree_rt.launches;
  e->tree_rt.launches = 0;
  if (e->prof_on) e->prof_ticks += n_ticks;  // phases overlap here: only the tick count is kept, bench.py reads the step time instead
  return AZ_OK;
}

extern "C" int az_selfplay_tick(az_engine* e, int32_t n_ticks) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (!e->selfplay) return az_fail(AZ_ERR_STATE, "az_selfplay_tick: call az_selfplay_begin first");
  if (e->pipeline && n_ticks > 0) return selfplay_tick_pipelined(e, n_ticks);
  const AzDims& d = e->E.d;
#ifndef AZ_EMU
  const bool prof = e->prof_on;
  if (prof) {
    prof_fold(e);
    while (e->prof_ev.size() < (size_t)n_ticks * 5) {
      cudaEvent_t ev;
      if (cudaEventCreate(&ev) != cudaSuccess) return az_fail(AZ_ERR_CUDA, "az_selfplay_tick: cudaEventCreate failed");
      e->prof_ev.push_back(ev);
    }
  }
#define AZ_PROF_MARK(k) do { if (prof) cudaEventRecord(e->prof_ev[(size_t)t * 5 + (k)], e->rt.stream); } while (0)
#else
#define AZ_PROF_MARK(k) do { } while (0)
#endif
  if (e->fused_tree && !e->prof_on && d.node_cache && n_ticks > 0) {
    // Fused tree pass: expand/backup -> move / re-root -> leaf collection of a game run back to back in one warp (games are
    // independent), one launch instead of three and one straggler tail instead of three.  The per-phase profile (az_tick_profile)
    // keeps the separate launches below.
    auto tree_pass = [&](int phases) {
#ifdef AZ_EMU
      k_tree_p(e->E, 0, d.G, phases, nullptr);
      if (phases & AZ_PH_COLLECT) k_compact(e->E, d.G);
#else
      k_tree_p<<<(d.G + AZ_WPB - 1) / AZ_WPB, AZ_WPB * 32, AZ_WPB * az_sim_stride(d), e->rt.stream>>>(e->E, 0, d.G, phases, nullptr);
      if (phases & AZ_PH_COLLECT) k_compact<<<1, 1024, 0, e->rt.stream>>>(e->E, d.G);
#endif
      e->rt.launches += (phases & AZ_PH_COLLECT) ? 2 : 1;
    };
    tree_pass(AZ_PH_COLLECT);
    for (int t = 0; t < n_ticks; ++t) {
#ifndef AZ_EMU
      if (t == n_ticks - 1) cudaEventRecord(e->ev0, e->rt.stream);
#endif
      int rc = aznet_forward(e->net, e->rt, e->E.leaf_obs, e->E.leaf_rows, e->E.leaf_total, d.G * d.Pmax, e->E.priors, e->E.values, d.Ap);
      if (rc) return az_fail(rc, "network forward: " + g_az_error);
#ifndef AZ_EMU
      if (t == n_ticks - 1) cudaEventRecord(e->ev1, e->rt.stream);
#endif
      tree_pass(AZ_PH_APPLY | AZ_PH_ADVANCE | (t < n_ticks - 1 ? AZ_PH_COLLECT : 0));
    }
#ifndef AZ_EMU
    cudaError_t fe = cudaGetLastError();
    if (fe != cudaSuccess) return az_fail(AZ_ERR_CUDA, std::string("CUDA launch error: ") + cudaGetErrorString(fe));
#endif
    return AZ_OK;
  }
  for (int t = 0; t < n_ticks; ++t) {
    AZ_PROF_MARK(0);
    launch_collect(e);
#ifndef AZ_EMU
    e->rt.launches++;
    k_compact<<<1, 1024, 0, e->rt.stream>>>(e->E, d.G);
    if (t == n_ticks - 1) cudaEventRecord(e->ev0, e->rt.stream);
#else
    k_compact(e->E, d.G);
#endif
    AZ_PROF_MARK(1);
    int rc = aznet_forward(e->net, e->rt, e->E.leaf_obs, e->E.leaf_rows, e->E.leaf_total, d.G * d.Pmax, e->E.priors, e->E.values, d.Ap);
    if (rc) return az_fail(rc, "network forward: " + g_az_error);
#ifndef AZ_EMU
    if (t == n_ticks - 1) cudaEventRecord(e->ev1, e->rt.stream);
#endif
    AZ_PROF_MARK(2);
    AZ_LAUNCH_WARPS(e->rt, k_apply, d.G, d, e->E);
    AZ_PROF_MARK(3);
    launch_advance(e);
    AZ_PROF_MARK(4);
  }
#undef AZ_PROF_MARK
#ifndef AZ_EMU
  if (prof) e->prof_pending = n_ticks;
#else
  if (e->prof_on) e->prof_ticks += n_ticks;
#endif
#ifndef AZ_EMU
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return az_fail(AZ_ERR_CUDA, std::string("CUDA launch error: ") + cudaGetErrorString(ce));
#endif
  return AZ_OK;
}

// ---- evaluation matches on the device (pipeline.py:815-867 for many games at once) ---------------------------------------
extern "C" int az_match_begin(az_engine* e, const az_search_params* p, const uint8_t* black_net, int32_t games_per_slot, int32_t alternate) {
  AZ_ENTER(e);
  if (!e || !p) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  if (!e->net || !aznet_ready(e->net) || !e->net2 || !aznet_ready(e->net2))
    return az_fail(AZ_ERR_STATE, "az_match_begin: both weight sets must be loaded (az_set_weights_for 0 and 1)");
  if (games_per_slot < 1) return az_fail(AZ_ERR_BAD_ARG, "az_match_begin: games_per_slot must be positive");
  int rc = set_search_cfg(e, p);
  if (rc) return rc;
  const AzDims& d = e->E.d;
  AzSearchCfg& s = e->E.s;
  s.selfplay = 1;
  s.host_noise = 0;
  s.warm_up_steps = -1;          // warm_up=False at every ply (pipeline.py:838)
  s.check_resign_after = 1 << 30;
  s.resign_threshold = -1.0f;    // no resignation in evaluation games
  s.disable_resign_ratio = 1.0f;
  s.match = 1;
  s.match_games = games_per_slot;
  s.match_alternate = alternate ? 1 : 0;
  std::vector<uint8_t> bn((size_t)d.G, 0);
  if (black_net)
    for (int g = 0; g < d.G; ++g) bn[g] = black_net[g] ? 1 : 0;
  rt_h2d(e->rt, e->E.slot_net, bn.data(), bn.size());
  e->selfplay = true;
  e->active.clear();
  rt_zero(e->rt, e->E.counters, CT_COUNT * sizeof(unsigned long long));
  e->drained_games = 0;
  e->dropped_samples = 0;
  AZ_LAUNCH_WARPS(e->rt, k_selfplay_begin, d.G, d, e->E);
  return rt_sync(e->rt);
}

extern "C" int az_match_tick(az_engine* e, int32_t n_ticks, int32_t* n_running) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
  if (!e->selfplay || !e->E.s.match) return az_fail(AZ_ERR_STATE, "az_match_tick: call az_match_begin first");
  const AzDims& d = e->E.d;
  for (int t = 0; t < n_ticks; ++t) {
    launch_collect(e);
#ifdef AZ_EMU
    k_compact_match(e->E, d.G);
#else
    k_compact_match<<<1, 1024, 0, e->rt.stream>>>(e->E, d.G);
#endif
    e->rt.launches++;
    // each weight set evaluates the leaves of the games in which its colour is to move; rows are disjoint, outputs land in the same
    // per-slot priors / values buffers
    int rc = aznet_forward(e->net, e->rt, e->E.leaf_obs, e->E.leaf_rows, e->E.leaf_total, d.G * d.Pmax, e->E.priors, e->E.values, d.Ap);
    if (!rc) rc = aznet_forward(e->net2, e->rt, e->E.leaf_obs, e->E.leaf_rows2, e->E.leaf_total + 4, d.G * d.Pmax, e->E.priors, e->E.values, d.Ap);
    if (rc) return az_fail(rc, "network forward: " + g_az_error);
    AZ_LAUNCH_WARPS(e->rt, k_apply, d.G, d, e->E);
    launch_advance(e);
  }
  if (n_running) {
    rt_zero(e->rt, e->d_out, sizeof(int32_t));
    AZ_LAUNCH_THREADS(e->rt, k_count_active, d.G, e->E, e->d_out);
    int32_t act = 0;
    rt_d2h(e->rt, &act, e->d_out, sizeof(act));
    *n_running = act;
  }
  return rt_sync(e->rt);
}

extern "C" int az_sync(az_engine* e) {
  AZ_ENTER(e);
  return e ? rt_sync(e->rt) : az_fail(AZ_ERR_BAD_ARG, "null engine");
}

extern "C" int az_get_counters(az_engine* e, az_counters* out) {
  AZ_ENTER(e);
  if (!e || !out) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  unsigned long long c[CT_COUNT];
  rt_d2h(e->rt, c, e->E.counters, sizeof(c));
  out->simulations = c[CT_SIMS];
  out->evaluations = c[CT_EVALS];
  out->moves = c[CT_MOVES];
  out->games = c[CT_GAMES];
  out->nodes = c[CT_NODES];
  out->depth_sum = c[CT_DEPTH];
  out->descents = c[CT_DESCENTS];
  out->samples = c[CT_SAMPLES];
  out->ring_dropped = c[CT_DROPPED] + e->dropped_samples;
  out->errors = c[CT_ERRORS];
  out->kernel_launches = e->rt.launches;
  out->ticks = 0;
  return AZ_OK;
}

static int drain_impl(az_engine* e, az_game_record* records, int32_t max_games, int32_t* n_games, int8_t* states, float* pis,
                      float* values, int16_t* moves, int32_t max_samples, int32_t* n_samples, bool dst_on_device);

extern "C" int az_drain_games(az_engine* e, az_game_record* records, int32_t max_games, int32_t* n_games, int8_t* states,
                              float* pis, float* values, int16_t* moves, int32_t max_samples, int32_t* n_samples) {
  AZ_ENTER(e);
  return drain_impl(e, records, max_games, n_games, states, pis, values, moves, max_samples, n_samples, false);
}

extern "C" int az_gather_pack(az_engine* e, az_game_record* records, int32_t max_games, int32_t* n_games, void* d_states, void* d_pis,
                              void* d_values, int32_t max_samples, int32_t* n_samples) {
  AZ_ENTER(e);
  if (!d_states || !d_pis || !d_values) return az_fail(AZ_ERR_BAD_ARG, "az_gather_pack: null device buffer");
  return drain_impl(e, records, max_games, n_games, (int8_t*)d_states, (float*)d_pis, (float*)d_values, nullptr, max_samples, n_samples, true);
}

// Finished games leave the sample ring oldest first, as maximal contiguous runs, into host buffers (az_drain_games: DMA to
// the caller's — ideally pinned — memory) or into caller-owned device buffers (az_gather_pack: the send block of the
// all-gather, packed device to device).
static int drain_impl(az_engine* e, az_game_record* records, int32_t max_games, int32_t* n_games, int8_t* states, float* pis,
                      float* values, int16_t* moves, int32_t max_samples, int32_t* n_samples, bool dst_on_device) {
  if (!e || !n_games || !n_samples) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  const AzDims& d = e->E.d;
  int rc = rt_sync(e->rt);
  if (rc) return rc;
  unsigned long long c[CT_COUNT];
  rt_d2h(e->rt, c, e->E.counters, sizeof(c));
  const unsigned long long head = c[CT_GAMES_HEAD];
  if (head - e->drained_games > AZ_GAMES_RING) e->drained_games = head - AZ_GAMES_RING;
  int ng = 0, ns = 0;
  // all pending game records in one or two copies (the ring of records is contiguous modulo its size)
  const unsigned long long pending = head - e->drained_games;
  std::vector<int32_t> recs((size_t)pending * GR_INTS);
  if (pending) {
    const size_t r0 = (size_t)(e->drained_games % AZ_GAMES_RING);
    const size_t k1 = std::min((size_t)pending, (size_t)AZ_GAMES_RING - r0), k2 = (size_t)pending - k1;
    rt_d2h_async(e->rt, recs.data(), e->E.games_ring + r0 * GR_INTS, k1 * GR_INTS * sizeof(int32_t));
    if (k2) rt_d2h_async(e->rt, recs.data() + k1 * GR_INTS, e->E.games_ring, k2 * GR_INTS * sizeof(int32_t));
    rc = rt_sync(e->rt);
    if (rc) return rc;
  }
  // Games finish one after the other, so the samples of consecutive games are consecutive in the sample ring: the accepted
  // games are copied as maximal contiguous runs (normally ONE run per call: two copies per array when it wraps) instead of
  // game by game.
  unsigned long long run_first = 0;
  int run_len = 0, run_dst = 0;
  auto cp = [&](void* dst, const void* src, size_t n) {
    if (dst_on_device) rt_d2d(e->rt, dst, src, n);
    else rt_d2h_async(e->rt, dst, src, n);
  };
  auto flush_run = [&]() {
    if (!run_len) return;
    const size_t s0 = (size_t)(run_first % (unsigned long long)d.ring_cap);
    const size_t n1 = std::min((size_t)run_len, (size_t)d.ring_cap - s0), n2 = (size_t)run_len - n1;
    if (states) {
      cp(states + (size_t)run_dst * d.obs_bytes, e->E.r_obs + s0 * d.obs_bytes, n1 * d.obs_bytes);
      if (n2) cp(states + (size_t)(run_dst + n1) * d.obs_bytes, e->E.r_obs, n2 * d.obs_bytes);
    }
    if (pis) {
      cp(pis + (size_t)run_dst * d.A, e->E.r_pi + s0 * d.A, n1 * d.A * sizeof(float));
      if (n2) cp(pis + (size_t)(run_dst + n1) * d.A, e->E.r_pi, n2 * d.A * sizeof(float));
    }
    if (values) {
      cp(values + run_dst, e->E.r_z + s0, n1 * sizeof(float));
      if (n2) cp(values + run_dst + n1, e->E.r_z, n2 * sizeof(float));
    }
    if (moves) {
      cp(moves + run_dst, e->E.r_move + s0, n1 * sizeof(int16_t));
      if (n2) cp(moves + run_dst + n1, e->E.r_move, n2 * sizeof(int16_t));
    }
    run_len = 0;
  };
  for (unsigned long long k = 0; k < pending && ng < max_games; ++k) {
    const int32_t* gr = recs.data() + (size_t)k * GR_INTS;
    const int len = gr[GR_LEN];
    if (ns + len > max_samples) break;
    if (records) {
      az_game_record& r = records[ng];
      r.slot = gr[GR_SLOT];
      r.game_length = len;
      r.winner = ext_player(d, gr[GR_WINNER]);
      r.by_resign = gr[GR_BY_RESIGN];
      memcpy(&r.score, &gr[GR_SCORE_BITS], 4);
      r.num_passes = gr[GR_PASSES];
      r.is_resign_disabled = gr[GR_RESIGN_DISABLED];
      r.is_marked_for_resign = gr[GR_MARKED_FOR_RESIGN];
      r.is_could_won = gr[GR_COULD_WON];
      r.marked_resign_player = ext_player(d, gr[GR_MARKED_PLAYER]);
      r.first_sample = ns;
      r.reserved = gr[GR_UID];
    }
    const unsigned long long first = (unsigned long long)(uint32_t)gr[GR_FIRST_LO] | ((unsigned long long)(uint32_t)gr[GR_FIRST_HI] << 32);
    if (c[CT_RING_HEAD] - first > (unsigned long long)d.ring_cap) {  // overwritten before the host came by: drop the game
      e->dropped_samples += (unsigned long long)len;
      e->drained_games++;
      continue;
    }
    if (run_len && first != run_first + (unsigned long long)run_len) flush_run();
    if (!run_len) { run_first = first; run_dst = ns; }
    run_len += len;
    ns += len;
    ng++;
    e->drained_games++;
  }
  flush_run();
  rc = rt_sync(e->rt);
  if (rc) return rc;
  *n_games = ng;
  *n_samples = ns;
  return AZ_OK;
}

extern "C" int az_host_alloc(size_t bytes, void** out) {
  if (!out) return az_fail(AZ_ERR_BAD_ARG, "null argument");
#ifndef AZ_EMU
  void* p = nullptr;
  cudaError_t ce = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    return az_fail(AZ_ERR_CUDA, std::string("az_host_alloc: ") + cudaGetErrorString(ce) + " (" + std::to_string(bytes) + " bytes)");
  }
  *out = p;
#else
  *out = malloc(bytes ? bytes : 1);
  if (!*out) return az_fail(AZ_ERR_CUDA, "az_host_alloc: out of host memory");
#endif
  return AZ_OK;
}

extern "C" int az_host_free(void* p) {
#ifndef AZ_EMU
  if (p) cudaFreeHost(p);
#else
  free(p);
#endif
  return AZ_OK;
}

extern "C" int az_stream(az_engine* e, void** cuda_stream) {
  AZ_ENTER(e);
  if (!e || !cuda_stream) return az_fail(AZ_ERR_BAD_ARG, "null argument");
  *cuda_stream = (void*)e->rt.stream;
  return AZ_OK;
}

extern "C" int az_last_net_ms(az_engine* e, float* ms, int32_t* n_evals) {
  AZ_ENTER(e);
  if (!e || !ms) return az_fail(AZ_ERR_BAD_ARG, "null argument");
#ifndef AZ_EMU
  // the conv launches alone (events recorded inside the forward, around the tower), mean over the last (up to 64) network calls
  int averaged = 0;
  *ms = e->net ? aznet_last_tower_ms(e->net, &averaged) : 0.f;
  int32_t tot[2] = {0, 0};
  rt_d2h(e->rt, tot, e->E.leaf_total + (e->pipeline ? 4 : 0), sizeof(tot));  // pipeline: the events bracket the second half's tower
  if (n_evals) *n_evals = tot[0];
#else
  *ms = 0.f;
  if (n_evals) *n_evals = 0;
#endif
  return AZ_OK;
}

extern "C" int az_tick_profile(az_engine* e, int32_t enable, double* ms5, int32_t* n_ticks) {
  AZ_ENTER(e);
  if (!e) return az_fail(AZ_ERR_BAD_ARG, "null engine");
#ifndef AZ_EMU
  prof_fold(e);
#endif
  if (ms5) for (int k = 0; k < 5; ++k) ms5[k] = e->prof_ms[k];
  if (n_ticks) *n_ticks = e->prof_ticks;
  for (int k = 0; k < 5; ++k) e->prof_ms[k] = 0.0;
  e->prof_ticks = 0;
  e->prof_on = enable != 0;
  return AZ_OK;
}

// ---- device-resident replay: the learner's input path -----------------------------------------
// Mirrors UniformReplay (core/replay.py:35-116): circular storage, add_game / add, uniform sampling with replacement by
// caller-supplied indices (so a numpy RandomState on the host reproduces the reference's minibatches), plus the batch-wide
// dihedral transformation of utils/transformation.py:160 applied while gathering.
extern "C" int az_replay_create(az_engine* e, int32_t capacity) {
  AZ_ENTER(e);
  if (!e || capacity <= 0) return az_fail(AZ_ERR_BAD_ARG, "Expect capacity to be a positive integer");
  if (e->E.rp_obs) return az_fail(AZ_ERR_STATE, "replay already created");
  const AzDims& d = e->E.d;
  e->E.rp_obs = dev_alloc<int8_t>(e, (size_t)capacity * d.obs_bytes);
  e->E.rp_pi = dev_alloc<float>(e, (size_t)capacity * d.A);
  e->E.rp_z = dev_alloc<float>(e, capacity);
  if (alloc_check(e, "az_replay_create")) {
    e->E.rp_obs = nullptr; e->E.rp_pi = nullptr; e->E.rp_z = nullptr;  // whatever was obtained stays in e->allocs and is freed by az_destroy
    return AZ_ERR_CUDA;
  }
  e->E.rp_cap = capacity;
  e->rp_samples_added = e->rp_games_added = 0;
  return AZ_OK;
}

static void replay_copy_in(az_engine* e, const int8_t* so, const float* sp, const float* sz, size_t n, bool from_device) {
  const AzDims& d = e->E.d;
  size_t done = 0;
  while (done < n) {  // circular: at most two segments per call unless n exceeds the capacity
    const size_t pos = (size_t)(e->rp_samples_added % e->E.rp_cap);
    const size_t m = std::min(n - done, (size_t)e->E.rp_cap - pos);
    if (from_device) {
      rt_d2d(e->rt, e->E.rp_obs + pos * d.obs_bytes, so + done * d.obs_bytes, m * d.obs_bytes);
      rt_d2d(e->rt, e->E.rp_pi + pos * d.A, sp + done * d.A, m * d.A * sizeof(float));
      rt_d2d(e->rt, e->E.rp_z + pos, sz + done, m * sizeof(float));
    } else {
      rt_h2d(e->rt, e->E.rp_obs + pos * d.obs_bytes, so + done * d.obs_bytes, m * d.obs_bytes);
      rt_h2d(e->rt, e->E.rp_pi + pos * d.A, sp + done * d.A, m * d.A * sizeof(float));
      rt_h2d(e->rt, e->E.rp_z + pos, sz + done, m * sizeof(float));
    }
    e->rp_samples_added += (long long)m;
    done += m;
  }
}

extern "C" int az_replay_add(az_engine* e, const int8_t* states, const float* pis, const float* values, int32_t n, int32_t n_games) {
  AZ_ENTER(e);
  if (!e || !e->E.rp_obs) return az_fail(AZ_ERR_STATE, "replay not created");
  if (n < 0 || (n > 0 && (!states || !pis || !values))) return az_fail(AZ_ERR_BAD_ARG, "az_replay_add: bad arguments");
  replay_copy_in(e, states, pis, values, (size_t)n, false);
  e->rp_games_added += n_games;
  return rt_sync(e->rt);
}

// Move every finished, not yet consumed game from the sample ring into the replay, device to device (add_game per game).
extern "C" int az_replay_ingest(az_engine* e, int32_t* n_games, int32_t* n_samples) {
  AZ_ENTER(e);
  if (!e || !e->E.rp_obs) return az_fail(AZ_ERR_STATE, "replay not created");
  const AzDims& d = e->E.d;
  int rc = rt_sync(e->rt);
  if (rc) return rc;
  unsigned long long c[CT_COUNT];
  rt_d2h(e->rt, c, e->E.counters, sizeof(c));
  const unsigned long long head = c[CT_GAMES_HEAD];
  if (head - e->drained_games > AZ_GAMES_RING) e->drained_games = head - AZ_GAMES_RING;
  int ng = 0, ns = 0;
  while (e->drained_games < head) {
    int32_t gr[GR_INTS];
    rt_d2h(e->rt, gr, e->E.games_ring + (size_t)(e->drained_games % AZ_GAMES_RING) * GR_INTS, sizeof(gr));
    const int len = gr[GR_LEN];
    const unsigned long long first = (unsigned long long)(uint32_t)gr[GR_FIRST_LO] | ((unsigned long long)(uint32_t)gr[GR_FIRST_HI] << 32);
    e->drained_games++;
    if (c[CT_RING_HEAD] - first > (unsigned long long)d.ring_cap) { e->dropped_samples += (unsigned long long)len; continue; }
    const size_t s0 = (size_t)(first % (unsigned long long)d.ring_cap);
    const size_t n1 = std::min((size_t)len, (size_t)d.ring_cap - s0), n2 = (size_t)len - n1;
    replay_copy_in(e, e->E.r_obs + s0 * d.obs_bytes, e->E.r_pi + s0 * d.A, e->E.r_z + s0, n1, true);
    if (n2) replay_copy_in(e, e->E.r_obs, e->E.r_pi, e->E.r_z, n2, true);
    e->rp_games_added += 1;
    ng++;
    ns += len;
  }
  if (n_games) *n_games = ng;
  if (n_samples) *n_samples = ns;
  return rt_sync(e->rt);
}

extern "C" int az_replay_info(az_engine* e, int64_t* num_samples_added, int64_t* num_games_added, int32_t* size, int32_t* capacity) {
  AZ_ENTER(e);
  if (!e || !e->E.rp_obs) return az_fail(AZ_ERR_STATE, "replay not created");
  if (num_samples_added) *num_samples_added = e->rp_samples_added;
  if (num_games_added) *num_games_added = e->rp_games_added;
  if (size) *size = (int32_t)std::min<long long>(e->rp_samples_added, e->E.rp_cap);
  if (capacity) *capacity = e->E.rp_cap;
  return AZ_OK;
}

extern "C" int az_replay_sample(az_engine* e, const int32_t* indices, int32_t batch, int32_t transform, int8_t* states, float* pis,
                                float* values, int32_t outputs_on_device) {
  AZ_ENTER(e);
  if (!e || !e->E.rp_obs) return az_fail(AZ_ERR_STATE, "replay not created");
  if (!indices || batch <= 0 || !states || !pis || !values) return az_fail(AZ_ERR_BAD_ARG, "az_replay_sample: bad arguments");
  if (transform < 0 || transform > 5) return az_fail(AZ_ERR_BAD_ARG, "az_replay_sample: transform must be in [0, 5]");
  const AzDims& d = e->E.d;
  const int size = (int)std::min<long long>(e->rp_samples_added, e->E.rp_cap);
  for (int i = 0; i < batch; ++i)
    if (indices[i] < 0 || indices[i] >= size) return az_fail(AZ_ERR_BAD_ARG, "az_replay_sample: index out of range");
  if (batch > e->rp_batch_cap) {  // staging grows with the largest batch seen
    e->d_rp_idx = dev_alloc<int32_t>(e, batch);
    e->d_rp_obs = dev_alloc<int8_t>(e, (size_t)batch * d.obs_bytes);
    e->d_rp_pi = dev_alloc<float>(e, (size_t)batch * d.A);
    e->d_rp_z = dev_alloc<float>(e, batch);
    if (alloc_check(e, "az_replay_sample")) { e->rp_batch_cap = 0; return AZ_ERR_CUDA; }
    e->rp_batch_cap = batch;
  }
  rt_h2d(e->rt, e->d_rp_idx, indices, (size_t)batch * sizeof(int32_t));
  int8_t* o_obs = outputs_on_device ? states : e->d_rp_obs;
  float* o_pi = outputs_on_device ? pis : e->d_rp_pi;
  float* o_z = outputs_on_device ? values : e->d_rp_z;
  AZ_LAUNCH_THREADS(e->rt, k_replay_sample, (long long)batch * (d.obs_bytes + d.A), e->E, e->d_rp_idx, transform, o_obs, o_pi, o_z);
  if (!outputs_on_device) {
    rt_d2h(e->rt, states, o_obs, (size_t)batch * d.obs_bytes);
    rt_d2h(e->rt, pis, o_pi, (size_t)batch * d.A * sizeof(float));
    rt_d2h(e->rt, values, o_z, (size_t)batch * sizeof(float));
  }
  return rt_sync(e->rt);
}
