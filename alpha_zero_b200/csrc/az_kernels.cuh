// az_kernels.cuh — __global__ entry points: one warp per game slot (4 warps per CTA, scratch position
// in dynamic shared memory), plus the small bookkeeping kernels around the evaluator.
#pragma once
#include "az_rt.h"
#include "az_tree.cuh"

#define AZ_WPB 4  // warps (games) per CTA

#ifdef AZ_EMU
#include <vector>
#define AZ_GLOBAL static void
#define AZ_WARP_INDEX(nwarps) for (int az_g = 0; az_g < (nwarps); ++az_g)
#define AZ_THREAD_LOOP(i, n) for (long long i = 0; i < (long long)(n); ++i)
static inline unsigned char* az_emu_scratch(size_t bytes) {
  static thread_local std::vector<unsigned char> buf;
  if (buf.size() < bytes) buf.resize(bytes);
  return buf.data();
}
#define AZ_SCRATCH(d, S) sim_carve((d), (S), az_emu_scratch(sim_bytes(d) + 64))
#define AZ_LAUNCH_WARPS(rt, kern, nwarps, d, ...) \
  do { (rt).launches++; kern(__VA_ARGS__, (nwarps)); } while (0)
#define AZ_LAUNCH_THREADS(rt, kern, n, ...) \
  do { (rt).launches++; kern(__VA_ARGS__, (n)); } while (0)
#else
#define AZ_GLOBAL __global__ void
#define AZ_WARP_INDEX(nwarps)                                                  \
  const int az_g = blockIdx.x * (blockDim.x >> 5) + (int)(threadIdx.x >> 5);   \
  if (az_g < (nwarps))
#define AZ_THREAD_LOOP(i, n)                                                              \
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)(n); \
       i += (long long)gridDim.x * blockDim.x)
static __host__ __device__ inline size_t az_sim_stride(const AzDims& d) { return ((size_t)d.ncp * 18 + d.Ap + AZ_PATH * 6 + 15) & ~(size_t)15; }
#define AZ_SCRATCH(d, S)                                        \
  extern __shared__ __align__(16) unsigned char az_smem[];      \
  sim_carve((d), (S), az_smem + (threadIdx.x >> 5) * az_sim_stride(d))
#define AZ_LAUNCH_WARPS(rt, kern, nwarps, d, ...)                                                           \
  do {                                                                                                       \
    (rt).launches++;                                                                                         \
    kern<<<((nwarps) + AZ_WPB - 1) / AZ_WPB, AZ_WPB * 32, AZ_WPB * az_sim_stride(d), (rt).stream>>>(__VA_ARGS__, (nwarps)); \
  } while (0)
#define AZ_LAUNCH_THREADS(rt, kern, n, ...)                                                  \
  do {                                                                                        \
    (rt).launches++;                                                                          \
    long long _b = ((long long)(n) + 255) / 256;                                              \
    if (_b < 1) _b = 1;                                                                       \
    if (_b > 148 * 16) _b = 148 * 16;                                                         \
    kern<<<(int)_b, 256, 0, (rt).stream>>>(__VA_ARGS__, (n));                                 \
  } while (0)
#endif

AZ_DEV void flush_counters(const AzState& E, const LocalCounters& lc) {
  W_LANE0 {
    if (lc.sims) atomic_add_u64(&E.counters[CT_SIMS], lc.sims);
    if (lc.evals) atomic_add_u64(&E.counters[CT_EVALS], lc.evals);
    if (lc.nodes) atomic_add_u64(&E.counters[CT_NODES], lc.nodes);
    if (lc.depth) atomic_add_u64(&E.counters[CT_DEPTH], lc.depth);
    if (lc.descents) atomic_add_u64(&E.counters[CT_DESCENTS], lc.descents);
    if (lc.errors) atomic_add_u64(&E.counters[CT_ERRORS], lc.errors);
  }
}

AZ_GLOBAL k_collect(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
#ifdef AZ_EMU
    if (E.d.node_cache) game_collect_nc(E, az_g, S, lc);
    else
#endif
    game_collect(E, az_g, S, lc);
    flush_counters(E, lc);
  }
}

#ifndef AZ_EMU
// Same kernel compiled for 7 CTAs (28 warps) per SM: 4096 games then fit in ONE wave on 148 SMs (the unconstrained build
// allocates 96 registers -> 5 CTAs per SM -> 2960 resident warps, i.e. a 1.4-wave launch).  Chosen with AZ_COLLECT_OCC=1.
__global__ void __launch_bounds__(AZ_WPB * 32, 7) k_collect_occ(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    game_collect(E, az_g, S, lc);
    flush_counters(E, lc);
  }
}
// Node-cache variants (d.node_cache): one ply per descent, see game_collect_nc.
__global__ void k_collect_nc(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    game_collect_nc(E, az_g, S, lc);
    flush_counters(E, lc);
  }
}
__global__ void __launch_bounds__(AZ_WPB * 32, 7) k_collect_nc_occ(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    game_collect_nc(E, az_g, S, lc);
    flush_counters(E, lc);
  }
}
#endif

AZ_GLOBAL k_apply(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    game_apply(E, az_g, lc);
    flush_counters(E, lc);
  }
}

AZ_GLOBAL k_advance(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    game_advance(E, az_g, S);
  }
}

// Self-play tick: the few games that move in a tick (one in ~50) each re-root a subtree of some hundred nodes; done by their own
// warp alone that serial copy is the kernel's duration.  k_advance_d copies the new root only and queues the slot; k_reroot_payload
// (launched right behind it) spreads the remaining nodes of all queued slots over the whole grid.
AZ_GLOBAL k_advance_d(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    game_advance(E, az_g, S, true);
  }
}

#ifdef AZ_EMU
static void k_reroot_payload(AzState E, int) { reroot_payload(E, 0, 1); }
#else
__global__ void __launch_bounds__(AZ_WPB * 32) k_reroot_payload(AzState E, int) {
  reroot_payload(E, blockIdx.x * (blockDim.x >> 5) + (int)(threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5));
}
#endif

AZ_GLOBAL k_selfplay_begin(AzState E, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    W_LANE0 E.tree_i[(size_t)az_g * TREE_INTS + TI_SLOT_GAMES] = 0;
    w_sync();
    game_new(E, az_g, S);
    W_LANE0 E.tree_i[(size_t)az_g * TREE_INTS + TI_ACTIVE] = 1;
  }
}

// az_selfplay_restart: the running game of slots[i] is abandoned (nothing emitted) and a new one starts in its place
AZ_GLOBAL k_selfplay_restart(AzState E, const int32_t* slots, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    game_new(E, slots[az_g], S);
  }
}

AZ_GLOBAL k_env_reset(AzState E, const int32_t* slots, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    const int g = slots[az_g];
    env_reset(E, g, S);
    W_LANE0 {
      int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
      ti[TI_STATE] = ST_IDLE;
      ti[TI_NODES] = 0;
      ti[TI_ACTIVE] = 0;
    }
  }
}

// step() with the reference's validation order (go.py:90-95): game over, out of range, illegal.
AZ_GLOBAL k_env_step(AzState E, const int32_t* slots, const int32_t* actions, int32_t* out, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    const int g = slots[az_g];
    const int a = actions[az_g];
    const int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
    int err = 0;
    const bool resign = (a == -1 && E.d.game == 0);
    if (ei[EI_DONE]) err = -4;
    else if (!resign && (a < 0 || a >= E.d.A)) err = -2;
    else if (!resign && E.root_legal[(size_t)g * E.d.Ap + a] != 1) err = -3;
    int rx2 = 0, done = 0;
    if (!err) {
      const StepOut o = env_step(E, g, S, a);
      rx2 = o.reward_x2;
      done = o.done;
    }
    W_LANE0 { out[az_g * 3] = err; out[az_g * 3 + 1] = rx2; out[az_g * 3 + 2] = done; }
  }
}

// Replay of recorded games (core/eval_dataset.py:166-215, the loop of replay_sgf): one warp resets its slot and plays its
// move list, writing the observation BEFORE every move (the dataset's `state`) to states[(offsets[i] + t)].  Stops at the
// first move step() would reject (same validation order as k_env_step); out = {moves played, error code} per game.
AZ_GLOBAL k_env_replay(AzState E, const int32_t* slots, const int16_t* moves, const int32_t* offsets, int8_t* states, int32_t* out, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    const int g = slots[az_g];
    env_reset(E, g, S);
    W_LANE0 {
      int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
      ti[TI_STATE] = ST_IDLE;
      ti[TI_NODES] = 0;
      ti[TI_ACTIVE] = 0;
    }
    w_sync();
    const int first = offsets[az_g], n = offsets[az_g + 1] - first;
    const int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
    int t = 0, err = 0;
    for (; t < n; ++t) {
      const int a = moves[first + t];
      if (ei[EI_DONE]) err = -4;
      else if (a < 0 || a >= E.d.A) err = -2;
      else if (E.root_legal[(size_t)g * E.d.Ap + a] != 1) err = -3;
      if (err) break;
      if (states) {
        sim_load(E, g, S);
        sim_write_obs(E.d, S, states + (size_t)(first + t) * E.d.obs_bytes);
      }
      env_step(E, g, S, a);
    }
    W_LANE0 { out[az_g * 2] = t; out[az_g * 2 + 1] = err; }
  }
}

AZ_GLOBAL k_env_obs(AzState E, int slot, int8_t* out, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    sim_load(E, slot, S);
    sim_write_obs(E.d, S, out);
  }
}

AZ_GLOBAL k_env_score(AzState E, int slot, float* out, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    sim_load(E, slot, S);
    const float sc = E.d.game == 0 ? go_score(E.d, S) : 0.f;
    W_LANE0 out[0] = sc;
  }
}

AZ_GLOBAL k_env_copy(AzState E, int src, int dst, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    const AzDims& d = E.d;
    W_FOR(c, d.ncp) E.board[(size_t)dst * d.ncp + c] = E.board[(size_t)src * d.ncp + c];
    W_FOR(c, 8 * d.ncp) E.hist[(size_t)dst * 8 * d.ncp + c] = E.hist[(size_t)src * 8 * d.ncp + c];
    W_FOR(c, ENV_INTS) E.env_i[(size_t)dst * ENV_INTS + c] = E.env_i[(size_t)src * ENV_INTS + c];
    W_FOR(c, d.Ap) E.root_legal[(size_t)dst * d.Ap + c] = E.root_legal[(size_t)src * d.Ap + c];
    W_LANE0 {
      int32_t* ti = E.tree_i + (size_t)dst * TREE_INTS;
      ti[TI_STATE] = ST_IDLE;
      ti[TI_NODES] = 0;
      ti[TI_ACTIVE] = 0;
    }
  }
}

// number of slots still playing (match loop: slots retire after their games): out[0], zeroed by the first thread of the grid first
AZ_GLOBAL k_count_active(AzState E, int32_t* out, int n) {
#ifdef AZ_EMU
  int c = 0;
  for (int i = 0; i < n; ++i) c += E.tree_i[(size_t)i * TREE_INTS + TI_ACTIVE] ? 1 : 0;
  out[0] = c;
#else
  // single CTA launch is not guaranteed by AZ_LAUNCH_THREADS: count with one atomic per thread that finds an active slot
  AZ_THREAD_LOOP(i, n) if (E.tree_i[(size_t)i * TREE_INTS + TI_ACTIVE]) atomicAdd(out, 1);
#endif
}

AZ_GLOBAL k_clear_active(AzState E, int n) {
  AZ_THREAD_LOOP(i, n) E.tree_i[(size_t)i * TREE_INTS + TI_ACTIVE] = 0;
}

// Arm a split-phase search on slots[]; reuse[i] keeps the re-rooted subtree (root_node argument).
AZ_GLOBAL k_search_begin(AzState E, const int32_t* slots, const int32_t* reuse, int warm, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    const int g = slots[az_g];
    int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const bool keep = reuse[az_g] != 0 && ti[TI_NODES] > 0;
    W_LANE0 {
      ti[TI_ACTIVE] = 1;
      ti[TI_WARM] = warm;
      ti[TI_STATE] = keep ? ST_SEARCH_INIT : ST_NEED_ROOT;
    }
    w_sync();
    if (keep) search_enter(E, g);
  }
}

AZ_GLOBAL k_commit(AzState E, int slot, int move, double* out, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    double bq = 0.0;
    const int kept = game_commit(E, slot, move, &bq);
    W_LANE0 { out[0] = bq; out[1] = (double)kept; E.tree_i[(size_t)slot * TREE_INTS + TI_ACTIVE] = 0; }
  }
}

// Compact the occupied leaf rows (slot-major, ascending) for the evaluator; count running searches.
#ifdef AZ_EMU
static void k_compact(AzState E, int n) {
  int total = 0, running = 0;
  for (int g = 0; g < n; ++g) {
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int c = ti[TI_ACTIVE] ? ti[TI_NLEAVES] : 0;
    E.leaf_count[g] = c;
    for (int j = 0; j < c; ++j) E.leaf_rows[total++] = g * E.d.Pmax + j;
    const int st = ti[TI_STATE];
    if (ti[TI_ACTIVE] && (st == ST_NEED_ROOT || st == ST_SEARCH_INIT || st == ST_SEARCHING)) running++;
  }
  E.leaf_total[0] = total;
  E.leaf_total[1] = running;
}
#else
__global__ void k_compact(AzState E, int n) {  // one CTA of 1024 threads
  __shared__ int s_sum[1024];
  __shared__ int s_run[32];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int g0 = t * per, g1 = min(n, g0 + per);
  int local = 0, running = 0;
  for (int g = g0; g < g1; ++g) {
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int c = ti[TI_ACTIVE] ? ti[TI_NLEAVES] : 0;
    E.leaf_count[g] = c;
    local += c;
    const int st = ti[TI_STATE];
    if (ti[TI_ACTIVE] && (st == ST_NEED_ROOT || st == ST_SEARCH_INIT || st == ST_SEARCHING)) running++;
  }
  s_sum[t] = local;
  running = w_sum_i(running);
  if ((t & 31) == 0) s_run[t >> 5] = running;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
    int v = t >= off ? s_sum[t - off] : 0;
    __syncthreads();
    s_sum[t] += v;
    __syncthreads();
  }
  int pos = s_sum[t] - local;
  for (int g = g0; g < g1; ++g) {
    const int c = E.leaf_count[g];
    for (int j = 0; j < c; ++j) E.leaf_rows[pos++] = g * E.d.Pmax + j;
  }
  if (t == 1023) E.leaf_total[0] = s_sum[1023];
  if (t == 0) {
    int r = 0;
    for (int k = 0; k < 32; ++k) r += s_run[k];
    E.leaf_total[1] = r;
  }
}
#endif

// Leaf-row compaction of the match loop: two lists, one per weight set (leaf_rows / leaf_total[0], leaf_rows2 / leaf_total[4]);
// leaf_total[1] = running searches.  One CTA of 1024 threads, like k_compact.
#ifdef AZ_EMU
static void k_compact_match(AzState E, int n) {
  int tot[2] = {0, 0}, running = 0;
  for (int g = 0; g < n; ++g) {
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int c = ti[TI_ACTIVE] ? ti[TI_NLEAVES] : 0;
    E.leaf_count[g] = c;
    if (c) {
      const int w = match_net_of(E, g);
      int32_t* rows = w ? E.leaf_rows2 : E.leaf_rows;
      for (int j = 0; j < c; ++j) rows[tot[w]++] = g * E.d.Pmax + j;
    }
    const int st = ti[TI_STATE];
    if (ti[TI_ACTIVE] && (st == ST_NEED_ROOT || st == ST_SEARCH_INIT || st == ST_SEARCHING)) running++;
  }
  E.leaf_total[0] = tot[0];
  E.leaf_total[1] = running;
  E.leaf_total[4] = tot[1];
}
#else
__global__ void __launch_bounds__(1024) k_compact_match(AzState E, int n) {
  __shared__ int s_sum[2][1024];
  __shared__ int s_run[32];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int g0 = t * per, g1 = min(n, g0 + per);
  int local[2] = {0, 0}, running = 0;
  for (int g = g0; g < g1; ++g) {
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int c = ti[TI_ACTIVE] ? ti[TI_NLEAVES] : 0;
    E.leaf_count[g] = c;
    if (c) local[match_net_of(E, g)] += c;
    const int st = ti[TI_STATE];
    if (ti[TI_ACTIVE] && (st == ST_NEED_ROOT || st == ST_SEARCH_INIT || st == ST_SEARCHING)) running++;
  }
  s_sum[0][t] = local[0];
  s_sum[1][t] = local[1];
  running = w_sum_i(running);
  if ((t & 31) == 0) s_run[t >> 5] = running;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of both columns
    const int v0 = t >= off ? s_sum[0][t - off] : 0, v1 = t >= off ? s_sum[1][t - off] : 0;
    __syncthreads();
    s_sum[0][t] += v0;
    s_sum[1][t] += v1;
    __syncthreads();
  }
  int pos[2] = {s_sum[0][t] - local[0], s_sum[1][t] - local[1]};
  for (int g = g0; g < g1; ++g) {
    const int c = E.leaf_count[g];
    if (!c) continue;
    const int w = match_net_of(E, g);
    int32_t* rows = w ? E.leaf_rows2 : E.leaf_rows;
    for (int j = 0; j < c; ++j) rows[pos[w]++] = g * E.d.Pmax + j;
  }
  if (t == 1023) { E.leaf_total[0] = s_sum[0][1023]; E.leaf_total[4] = s_sum[1][1023]; }
  if (t == 0) {
    int r = 0;
    for (int k = 0; k < 32; ++k) r += s_run[k];
    E.leaf_total[1] = r;
  }
}
#endif

// ---- slot-range kernels of the two-half pipeline (AZ_PIPELINE=1, az_engine.cu selfplay_tick_pipelined) ------------------
// The games are split into two halves; while the network evaluates the leaves of one half on the engine stream, the tree
// kernels of the other half run on a second stream.  Every kernel here is built for 7 CTAs per SM (<= 72 registers, 4 warps):
// exactly the registers (9216) and shared memory the dense-x conv kernel leaves free on an SM, so that the tree work can be
// co-resident with the persistent tensor-core kernel instead of waiting for its launch boundaries.
#ifdef AZ_EMU
#define AZ_GLOBAL_R static void
#else
#define AZ_GLOBAL_R __global__ void __launch_bounds__(AZ_WPB * 32, 7)
#endif

AZ_GLOBAL_R k_collect_r(AzState E, int g0, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    game_collect_nc(E, g0 + az_g, S, lc);
    flush_counters(E, lc);
  }
}

AZ_GLOBAL_R k_apply_r(AzState E, int g0, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    game_apply(E, g0 + az_g, lc);
    flush_counters(E, lc);
  }
}

AZ_GLOBAL_R k_advance_r(AzState E, int g0, int nwarps) {
  AZ_WARP_INDEX(nwarps) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    game_advance(E, g0 + az_g, S);
  }
}

// Persistent fused tree pass of the slot range [g0, g0 + ng) (AZ_PIPELINE=2): games are independent and, per game, expand/backup ->
// move/re-root -> leaf collection only depend on the game's own state, so ONE warp runs the requested phases of a game back to
// back and then fetches the next game from a global counter.  The grid is at most one 4-warp CTA per SM: together with the <= 72
// registers and 8 KB of scratch that is exactly what a persistent dense-x conv CTA leaves free on its SM, so this kernel is placed
// beside the tensor-core kernel of the other half of the slots instead of queueing behind it (a 512-CTA launch fills the SMs
// between two conv launches and starves the conv clusters: measured, profiles/r02_bench_go9_c2_n1_pipeline_v1.json).  The dynamic
// distribution also takes the slowest games (late positions: 16 tries, 30-ply paths) off the critical path of a wave.
#define AZ_PH_APPLY 1
#define AZ_PH_ADVANCE 2
#define AZ_PH_COLLECT 4
#ifdef AZ_EMU
static void k_tree_p(AzState E, int g0, int ng, int phases, unsigned int* ctr) {
  for (int i = 0; i < ng; ++i) {
    Sim S;
    AZ_SCRATCH(E.d, S);
    LocalCounters lc = {0, 0, 0, 0, 0, 0};
    if (phases & AZ_PH_APPLY) game_apply(E, g0 + i, lc);
    if (phases & AZ_PH_ADVANCE) game_advance(E, g0 + i, S);
    if (phases & AZ_PH_COLLECT) game_collect_nc(E, g0 + i, S, lc);
    flush_counters(E, lc);
  }
  (void)ctr;
}
#else
__global__ void __launch_bounds__(AZ_WPB * 32, 7) k_tree_p(AzState E, int g0, int ng, int phases, unsigned int* ctr) {
  Sim S;
  AZ_SCRATCH(E.d, S);
  LocalCounters lc = {0, 0, 0, 0, 0, 0};
  for (bool first = true;; first = false) {
    unsigned int i = 0;
    if (ctr) {
      if ((threadIdx.x & 31) == 0) i = atomicAdd(ctr, 1u);
      i = __shfl_sync(0xffffffffu, i, 0);
    } else {  // one game per warp, the grid covers the range (serial tick: a single wave)
      if (!first) break;
      i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    }
    if (i >= (unsigned int)ng) break;
    const int g = g0 + (int)i;
    if (phases & AZ_PH_APPLY) game_apply(E, g, lc);
    if (phases & AZ_PH_ADVANCE) game_advance(E, g, S);
    if (phases & AZ_PH_COLLECT) game_collect_nc(E, g, S, lc);
    __syncwarp();
  }
  flush_counters(E, lc);
}
#endif

// Leaf-row compaction of the slots [g0, g0 + n): rows go to leaf_rows + g0 * Pmax, the totals to tot[0..1].
#ifdef AZ_EMU
static void k_compact_r(AzState E, int g0, int n, int32_t* tot) {
  int total = 0, running = 0;
  int32_t* rows = E.leaf_rows + (size_t)g0 * E.d.Pmax;
  for (int g = g0; g < g0 + n; ++g) {
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int c = ti[TI_ACTIVE] ? ti[TI_NLEAVES] : 0;
    E.leaf_count[g] = c;
    for (int j = 0; j < c; ++j) rows[total++] = g * E.d.Pmax + j;
    const int st = ti[TI_STATE];
    if (ti[TI_ACTIVE] && (st == ST_NEED_ROOT || st == ST_SEARCH_INIT || st == ST_SEARCHING)) running++;
  }
  tot[0] = total;
  tot[1] = running;
}
#else
__global__ void __launch_bounds__(128, 7) k_compact_r(AzState E, int g0, int n, int32_t* tot) {  // one CTA of 128 threads
  __shared__ int s_sum[128];
  __shared__ int s_run[4];
  const int t = threadIdx.x;
  const int per = (n + 127) / 128;
  const int a = g0 + t * per, b = min(g0 + n, a + per);
  int32_t* rows = E.leaf_rows + (size_t)g0 * E.d.Pmax;
  int local = 0, running = 0;
  for (int g = a; g < b; ++g) {
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int c = ti[TI_ACTIVE] ? ti[TI_NLEAVES] : 0;
    E.leaf_count[g] = c;
    local += c;
    const int st = ti[TI_STATE];
    if (ti[TI_ACTIVE] && (st == ST_NEED_ROOT || st == ST_SEARCH_INIT || st == ST_SEARCHING)) running++;
  }
  s_sum[t] = local;
  running = w_sum_i(running);
  if ((t & 31) == 0) s_run[t >> 5] = running;
  __syncthreads();
  for (int off = 1; off < 128; off <<= 1) {  // Hillis-Steele inclusive scan
    int v = t >= off ? s_sum[t - off] : 0;
    __syncthreads();
    s_sum[t] += v;
    __syncthreads();
  }
  int pos = s_sum[t] - local;
  for (int g = a; g < b; ++g) {
    const int c = E.leaf_count[g];
    for (int j = 0; j < c; ++j) rows[pos++] = g * E.d.Pmax + j;
  }
  if (t == 127) tot[0] = s_sum[127];
  if (t == 0) tot[1] = s_run[0] + s_run[1] + s_run[2] + s_run[3];
}
#endif

AZ_GLOBAL k_gather_obs(AzState E, int8_t* dst, long long n) {  // n = total * obs_bytes
  AZ_THREAD_LOOP(i, n) {
    const long long row = i / E.d.obs_bytes, k = i - row * E.d.obs_bytes;
    dst[i] = E.leaf_obs[(size_t)E.leaf_rows[row] * E.d.obs_bytes + k];
  }
}

AZ_GLOBAL k_scatter_eval(AzState E, const float* pri, const float* val, long long n) {  // n = total * A
  AZ_THREAD_LOOP(i, n) {
    const long long row = i / E.d.A, a = i - row * E.d.A;
    const size_t dst = (size_t)E.leaf_rows[row];
    E.priors[dst * E.d.Ap + a] = pri[i];
    if (a == 0) E.values[dst] = val[row];
  }
}

// Minibatch gather from the device replay with one dihedral transformation applied to the whole batch, exactly like
// apply_random_transformation (utils/transformation.py:160): 0 none, 1 h_flip, 2 v_flip, 3/4/5 rotate 90/180/270
// counter-clockwise (torchvision.rotate); the pass probability (Go) is not moved.
AZ_DEV int az_src_cell(int n, int y, int x, int t) {
  switch (t) {
    case 1: return y * n + (n - 1 - x);
    case 2: return (n - 1 - y) * n + x;
    case 3: return x * n + (n - 1 - y);
    case 4: return (n - 1 - y) * n + (n - 1 - x);
    case 5: return (n - 1 - x) * n + y;
    default: return y * n + x;
  }
}

AZ_GLOBAL k_replay_sample(AzState E, const int32_t* idx, int transform, int8_t* out_obs, float* out_pi, float* out_z, long long n) {
  // n = batch * (obs_bytes + A): one thread per output element
  const AzDims& d = E.d;
  const long long per = (long long)d.obs_bytes + d.A;
  AZ_THREAD_LOOP(i, n) {
    const long long b = i / per;
    const int k = (int)(i - b * per);
    const size_t src = (size_t)idx[b];
    if (k < d.obs_bytes) {
      const int plane = k / d.nc, c = k - plane * d.nc;
      const int y = c / d.n, x = c - y * d.n;
      out_obs[(size_t)b * d.obs_bytes + k] = E.rp_obs[src * d.obs_bytes + (size_t)plane * d.nc + az_src_cell(d.n, y, x, transform)];
    } else {
      const int a = k - d.obs_bytes;
      int sa = a;
      if (a < d.nc) { const int y = a / d.n, x = a - y * d.n; sa = az_src_cell(d.n, y, x, transform); }
      out_pi[(size_t)b * d.A + a] = E.rp_pi[src * d.A + sa];
      if (a == 0) out_z[b] = E.rp_z[src];
    }
  }
}
