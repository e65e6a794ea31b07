// az_tree.cuh — PUCT search on one warp per game, over the structure-of-arrays tree in HBM.
//
// Reference semantics: /root/reference/alpha_zero/core/mcts_v2.py
//   best_child + child_U/child_Q  :99-109, :142-185   -> tree_pick      (warp-shuffle argmax)
//   expand                        :188-210            -> tree_expand
//   backup                        :213-232            -> tree_backup
//   add/revert_virtual_loss       :453-482            -> tree_vloss
//   add_dirichlet_noise           :235-262            -> search_enter
//   uct_search / parallel_uct_search loops :378-421 / :568-625 -> game_collect + game_apply
//   generate_search_policy, move choice, re-root :265-298, :630-655 -> search_finish, game_commit, game_advance
// and core/pipeline.py:289-382 (play_and_record_one_game) -> game_advance.
// Float semantics that are reproduced on purpose are listed in SURVEY.md section 9 / DESIGN.md.
#pragma once
#include "az_board.cuh"

#ifdef AZ_EMU
AZ_DEV int az_popc(uint32_t x) { return __builtin_popcount(x); }
#else
AZ_DEV int az_popc(uint32_t x) { return __popc(x); }
#endif

struct TreeView {
  float *N, *W, *P;
  int16_t* cidx;
  int16_t *parent, *pmove, *vloss;
  int8_t* to_play;
  uint8_t* expanded;
  int8_t* nboard;   // per-node position cache (null unless d.node_cache)
  uint8_t* nlegal;
  int16_t* nko;
};

AZ_DEV TreeView tree_view(const AzState& E, int g, int buf) {
  TreeView T;
  const size_t nb = ((size_t)g * 2 + buf) * E.d.cap;
  T.N = E.cN + nb * E.d.Ap;
  T.W = E.cW + nb * E.d.Ap;
  T.P = E.cP + nb * E.d.Ap;
  T.cidx = E.cidx + nb * E.d.Ap;
  T.parent = E.parent + nb;
  T.pmove = E.pmove + nb;
  T.vloss = E.nvloss + nb;
  T.to_play = E.nto_play + nb;
  T.expanded = E.expanded + nb;
  T.nboard = E.nboard ? E.nboard + nb * E.d.ncp : nullptr;
  T.nlegal = E.nlegal ? E.nlegal + nb * E.d.Ap : nullptr;
  T.nko = E.nko ? E.nko + nb : nullptr;
  return T;
}

struct LocalCounters { unsigned long long sims, evals, nodes, depth, descents, errors; };

AZ_DEV void tree_init_node(const AzState& E, TreeView& T, int idx, int parent, int move, int to_play) {
  const int Ap = E.d.Ap;
  W_FOR(a, Ap) {
    T.N[(size_t)idx * Ap + a] = 0.f;
    T.W[(size_t)idx * Ap + a] = 0.f;
    T.P[(size_t)idx * Ap + a] = 0.f;
    T.cidx[(size_t)idx * Ap + a] = -1;
  }
  W_LANE0 {
    T.parent[idx] = (int16_t)parent;
    T.pmove[idx] = (int16_t)move;
    T.to_play[idx] = (int8_t)to_play;
    T.expanded[idx] = 0;
    T.vloss[idx] = 0;
    if (T.nko) T.nko[idx] = -1;  // neither a ko point nor a cached terminal result yet
  }
}

// A terminal child is never expanded and every later visit re-steps the environment to find the same result again
// (mcts_v2.py:604-608).  With the node cache the result is kept in the node's ko field instead (a terminal node has no ko point
// to remember): AZ_TERM_BASE + 2 * reward of the mover.  Late positions hit terminals on most of their 2 * P tries, and the
// select kernel ends with its slowest warp.
#define AZ_TERM_BASE 30000

// Lazy child creation at selection time (mcts_v2.py:182-183).  `n_nodes` is the caller's register copy of the pool size.
AZ_DEV int tree_new_node(const AzState& E, TreeView& T, int parent, int move, int to_play, int& n_nodes, LocalCounters& lc) {
  const int idx = n_nodes;
  if (idx >= E.d.cap) { lc.errors++; return -1; }
  tree_init_node(E, T, idx, parent, move, to_play);
  W_LANE0 T.cidx[(size_t)parent * E.d.Ap + move] = (int16_t)idx;
  w_sync();
  n_nodes = idx + 1;
  lc.nodes++;
  return idx;
}

// argmax_a ( -Q(a) + U(a) ) over legal a with numpy's float semantics; lowest index wins ties.
// `n_node` is the visit count of `node` (its entry in the parent's row, already in a register from the level
// above; unused for the root).  Besides the action the caller gets the child's link (index | expanded flag, or
// -1) and the child's visit count, both taken from the rows this pick loaded anyway: a level of the descent
// costs ONE round of independent global loads.
AZ_DEV int tree_pick(const AzState& E, const TreeView& T, int g, int node, const uint8_t* legal, float n_node, int& child_link, float& n_child) {
  const int Ap = E.d.Ap;
  const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  double pbc;
  int n_i;
  bool use64 = false;
  if (node == 0) {
    n_i = (int)E.root_nw[(size_t)g * 2];
    if (n_i >= E.d.table_len) n_i = E.d.table_len - 1;
    pbc = ti[TI_ROOT_FRESH] ? E.pbc_fresh[n_i] : E.pbc_f32[n_i];
    use64 = ti[TI_ROOT_NOISED] != 0;
  } else {
    n_i = (int)n_node;
    if (n_i >= E.d.table_len) n_i = E.d.table_len - 1;
    pbc = E.pbc_f32[n_i];
  }
  const float pbc32 = (float)pbc;                 // Python float * float32 array -> float32 (mcts_v2.py:102)
  const float s32 = (float)E.sqrt_tab[n_i];       // math.sqrt(N) / float32 array -> float32
  const float* rN = T.N + (size_t)node * Ap;
  const float* rW = T.W + (size_t)node * Ap;
  const float* rP = T.P + (size_t)node * Ap;
  const int16_t* rC = T.cidx + (size_t)node * Ap;
  const double* p64 = E.root_p64 + (size_t)g * Ap;
  // The rows of a node are cold HBM lines, and a descent is a chain of such levels: all loads of a level are issued before the
  // first value is consumed (PICK_CH elements per lane at a time), so a level costs one memory round trip instead of one per
  // operand (ncu source view of the previous loop: 57 % of the kernel's stall samples sat on the first use of each row).
  constexpr int PICK_CH = 3;
  int bi = 1 << 30, bc = -1;
  float bn = 0.f;
  if (!use64) {
    float best = -INFINITY;
    for (int a0 = 0; a0 < E.d.A; a0 += PICK_CH * AZ_WIDTH) {
      float n_[PICK_CH], w_[PICK_CH], p_[PICK_CH];
      int c_[PICK_CH], l_[PICK_CH];
#pragma unroll
      for (int k = 0; k < PICK_CH; ++k) {
        const int a = a0 + k * AZ_WIDTH + AZ_LANE;
        n_[k] = 0.f; w_[k] = 0.f; p_[k] = 0.f; c_[k] = -1; l_[k] = 0;
        if (a < E.d.A) { n_[k] = rN[a]; c_[k] = rC[a]; w_[k] = rW[a]; p_[k] = rP[a]; l_[k] = legal[a]; }
      }
#pragma unroll
      for (int k = 0; k < PICK_CH; ++k) {
        const int a = a0 + k * AZ_WIDTH + AZ_LANE;
        if (a < E.d.A) {
          const float ratio = f_div(s32, f_add(1.0f, n_[k]));
          const float q = f_div(w_[k], n_[k] > 0.f ? n_[k] : 1.0f);
          float sc = f_add(-q, f_mul(f_mul(pbc32, p_[k]), ratio));
          if (l_[k] != 1) sc = -9999.0f;
          if (sc > best) { best = sc; bi = a; bc = c_[k]; bn = n_[k]; }
        }
      }
#ifndef AZ_EMU
      // Hint: while the scores are reduced, pull the rows of the most-visited expanded child towards L2 - PUCT picks it more often
      // than not, and the next level then starts on warm lines.  Results never depend on it.
      if (a0 == 0 && T.nboard != nullptr) {
        uint32_t nk = 0;
        int ck = -1;
#pragma unroll
        for (int k = 0; k < PICK_CH; ++k)
          if ((c_[k] & AZ_CIDX_EXPANDED) && c_[k] >= 0 && __float_as_uint(n_[k]) >= nk) { nk = __float_as_uint(n_[k]); ck = c_[k] & AZ_CIDX_MASK; }
        const uint32_t top = __reduce_max_sync(AZ_FULL, ck >= 0 ? nk : 0u);
        const uint32_t who = __ballot_sync(AZ_FULL, ck >= 0 && nk == top);
        if (who) {
          const int pc = __shfl_sync(AZ_FULL, ck, __ffs(who) - 1);
          const size_t po = (size_t)pc * Ap;
          w_prefetch(T.N + po, Ap * 4);
          w_prefetch(T.W + po, Ap * 4);
          w_prefetch(T.P + po, Ap * 4);
          w_prefetch(T.cidx + po, Ap * 2);
          w_prefetch(T.nlegal + po, Ap);
        }
      }
#endif
    }
    w_argmax_f(best, bi);
  } else {
    double best = -1e300;
    for (int a0 = 0; a0 < E.d.A; a0 += PICK_CH * AZ_WIDTH) {
      float n_[PICK_CH], w_[PICK_CH];
      double p_[PICK_CH];
      int c_[PICK_CH], l_[PICK_CH];
#pragma unroll
      for (int k = 0; k < PICK_CH; ++k) {
        const int a = a0 + k * AZ_WIDTH + AZ_LANE;
        n_[k] = 0.f; w_[k] = 0.f; p_[k] = 0.0; c_[k] = -1; l_[k] = 0;
        if (a < E.d.A) { n_[k] = rN[a]; c_[k] = rC[a]; w_[k] = rW[a]; p_[k] = p64[a]; l_[k] = legal[a]; }
      }
#pragma unroll
      for (int k = 0; k < PICK_CH; ++k) {
        const int a = a0 + k * AZ_WIDTH + AZ_LANE;
        if (a < E.d.A) {
          const float ratio = f_div(s32, f_add(1.0f, n_[k]));
          const float q = f_div(w_[k], n_[k] > 0.f ? n_[k] : 1.0f);
          double sc = d_add((double)(-q), d_mul(d_mul(pbc, p_[k]), (double)ratio));
          if (l_[k] != 1) sc = -9999.0;
          if (sc > best) { best = sc; bi = a; bc = c_[k]; bn = n_[k]; }
        }
      }
    }
    w_argmax(best, bi);
  }
  // the winner is the local best of the lane that owns it (same tie-break locally and globally)
  child_link = w_bcast_i(bc, bi);
  n_child = w_bcast_f(bn, bi);
  return bi;
}

AZ_DEV void root_add_w(const AzState& E, int g, bool fresh, float v) {
  double* nw = E.root_nw + (size_t)g * 2;
  nw[1] = fresh ? (nw[1] + (double)v) : (double)f_add((float)nw[1], v);
}

AZ_DEV void tree_backup(const AzState& E, TreeView& T, int g, int node, float value, LocalCounters& lc) {
  const int Ap = E.d.Ap;
  W_LANE0 {
    float v = value;
    while (node != 0) {
      const size_t k = (size_t)T.parent[node] * Ap + T.pmove[node];
      T.N[k] = f_add(T.N[k], 1.0f);
      T.W[k] = f_add(T.W[k], v);
      node = T.parent[node];
      v = -v;
    }
    E.root_nw[(size_t)g * 2] += 1.0;
    root_add_w(E, g, E.tree_i[(size_t)g * TREE_INTS + TI_ROOT_FRESH] != 0, v);
  }
  w_sync();
  lc.sims++;
}

// sign=+1: add_virtual_loss; sign=-1: revert_virtual_loss.  W only, N untouched, float32 arithmetic.
AZ_DEV void tree_vloss(const AzState& E, TreeView& T, int g, int node, int sign) {
  const int Ap = E.d.Ap;
  W_LANE0 {
    const bool fresh = E.tree_i[(size_t)g * TREE_INTS + TI_ROOT_FRESH] != 0;
    for (;;) {
      bool touch = true;
      if (sign > 0) T.vloss[node] += 1;
      else if (T.vloss[node] > 0) T.vloss[node] -= 1;
      else touch = false;
      if (touch) {
        if (node == 0) root_add_w(E, g, fresh, (float)sign);
        else {
          const size_t k = (size_t)T.parent[node] * Ap + T.pmove[node];
          T.W[k] = f_add(T.W[k], (float)sign);
        }
      }
      if (node == 0) break;
      node = T.parent[node];
    }
  }
  w_sync();
}

// ---- path-parallel statistics updates --------------------------------------------------------------
// During a descent the flat statistics index k = parent*Ap + move and the node id of every step are
// recorded (shared memory, then leaf_pk / leaf_pn in HBM for the apply pass).  Each node of a path is
// touched exactly once by a virtual-loss / backup walk, so the lanes update them independently and get
// bit-identical results to the reference's leaf-to-root loop.  Deeper than AZ_PATH: serial fallback.
AZ_DEV void path_vloss(const AzState& E, TreeView& T, int g, const int32_t* pk, const int16_t* pn, int depth, int sign) {
  const bool fresh = E.tree_i[(size_t)g * TREE_INTS + TI_ROOT_FRESH] != 0;
  W_FOR(i, depth) {
    const int node = pn[i];
    bool touch = true;
    if (sign > 0) T.vloss[node] += 1;
    else if (T.vloss[node] > 0) T.vloss[node] -= 1;
    else touch = false;
    if (touch) T.W[pk[i]] = f_add(T.W[pk[i]], (float)sign);
  }
  W_LANE0 {
    bool touch = true;
    if (sign > 0) T.vloss[0] += 1;
    else if (T.vloss[0] > 0) T.vloss[0] -= 1;
    else touch = false;
    if (touch) root_add_w(E, g, fresh, (float)sign);
  }
  w_sync();
}

// backup(leaf, value) (mcts_v2.py:213-232): element i of the path is (depth-1-i) plies above the leaf.
AZ_DEV void path_backup(const AzState& E, TreeView& T, int g, const int32_t* pk, int depth, float value, LocalCounters& lc) {
  W_FOR(i, depth) {
    const float v = ((depth - 1 - i) & 1) ? -value : value;
    T.N[pk[i]] = f_add(T.N[pk[i]], 1.0f);
    T.W[pk[i]] = f_add(T.W[pk[i]], v);
  }
  W_LANE0 {
    E.root_nw[(size_t)g * 2] += 1.0;
    root_add_w(E, g, E.tree_i[(size_t)g * TREE_INTS + TI_ROOT_FRESH] != 0, (depth & 1) ? -value : value);
  }
  w_sync();
  lc.sims++;
}

AZ_DEV void tree_expand(const AzState& E, TreeView& T, int node, const float* prior) {
  const int Ap = E.d.Ap;
  W_FOR(a, E.d.A) T.P[(size_t)node * Ap + a] = prior[a];
  W_LANE0 {
    T.expanded[node] = 1;
    if (node != 0) {  // mirror the flag into the parent's link so that a descent does not have to load expanded[child]
      const size_t k = (size_t)T.parent[node] * Ap + T.pmove[node];
      T.cidx[k] = (int16_t)(T.cidx[k] | AZ_CIDX_EXPANDED);
    }
  }
  w_sync();
}

// ---- Dirichlet(0.03) on device (self-play mode; parity mode takes host samples) -------------------
AZ_DEV double az_gamma_small(uint64_t seed, uint64_t stream, uint64_t base, double alpha) {
  // Marsaglia-Tsang for alpha+1, boosted by U^(1/alpha)
  const double dd = alpha + 1.0 - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
  double out = 0.0;
  for (int it = 0; it < 64; ++it) {
    const double u1 = az_u01(az_rand64(seed, stream, base + 4 * it)) + 1e-300;
    const double u2 = az_u01(az_rand64(seed, stream, base + 4 * it + 1));
    const double u3 = az_u01(az_rand64(seed, stream, base + 4 * it + 2)) + 1e-300;
    const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    double v = 1.0 + cc * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u3) < 0.5 * x * x + dd - dd * v + dd * log(v)) { out = dd * v; break; }
  }
  const double u4 = az_u01(az_rand64(seed, stream, base + 3)) + 1e-300;
  return out * exp(log(u4) / alpha);
}

// SEARCH_INIT -> SEARCHING: mix the Dirichlet noise into the root prior (float64 result) and check
// the loop bound once (a re-used root may already have enough visits, mcts_v2.py:568).
AZ_DEV void search_finish(const AzState& E, int g);

AZ_DEV void search_enter(const AzState& E, int g) {
  const AzDims& d = E.d;
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  TreeView T = tree_view(E, g, ti[TI_BUF]);
  if (E.s.root_noise) {
    const uint8_t* legal = E.root_legal + (size_t)g * d.Ap;
    double* p64 = E.root_p64 + (size_t)g * d.Ap;
    double* nz = E.noise + (size_t)g * d.Ap;
    if (!E.s.host_noise) {
      const uint64_t ctr = (uint64_t)ti[TI_RNG_CTR];
      double part = 0.0;
      W_FOR(a, d.A) {
        const double gm = az_gamma_small(E.s.seed, ((uint64_t)ti[TI_GAME_UID] << 20) ^ (uint64_t)g, (ctr * 1024 + a) * 512, 0.03);
        nz[a] = gm;
        part += gm;
      }
      const double tot = w_sum_d(part);
      w_sync();
      W_FOR(a, d.A) nz[a] = tot > 0.0 ? nz[a] / tot : 1.0 / d.A;
      W_LANE0 ti[TI_RNG_CTR] += 1;
      w_sync();
    }
    const bool had = ti[TI_ROOT_NOISED] != 0;
    W_FOR(a, d.A) {
      const double noise = legal[a] == 1 ? nz[a] : 0.0;                       // legal * dirichlet  (mcts_v2.py:260)
      const double base = had ? d_mul(p64[a], 0.75) : (double)f_mul(T.P[a], 0.75f);  // child_P * (1 - eps)
      p64[a] = d_add(base, d_mul(noise, 0.25));                              //   + noise * eps   -> float64
    }
    W_LANE0 ti[TI_ROOT_NOISED] = 1;
    w_sync();
  }
  int st = ST_SEARCHING;
  if (E.root_nw[(size_t)g * 2] >= (double)E.s.sims_bound) { st = ST_DONE; search_finish(E, g); }
  W_LANE0 ti[TI_STATE] = st;
  w_sync();
}

// Search policy (mcts_v2.py:265-298), root value, deterministic move.
AZ_DEV void search_finish(const AzState& E, int g) {
  const AzDims& d = E.d;
  const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  TreeView T = tree_view(E, g, ti[TI_BUF]);
  const uint8_t* legal = E.root_legal + (size_t)g * d.Ap;
  double* pi = E.res_pi + (size_t)g * d.Ap;
  const bool warm = ti[TI_WARM] != 0;
  double part = 0.0, bestn = -1.0;
  int bi = 1 << 30;
  W_FOR(a, d.A) {
    const double n = (double)T.N[a];
    double v = legal[a] == 1 ? n : 0.0;
    if (!warm) v = v * v * v * v * v;  // exponent min(5, 1/0.1); exact for integer counts
    if (d.game == 1) v = (double)(float)v;  // Gomoku's mask is int8 -> float32 policy
    pi[a] = v;
    part += v;
    if (n > bestn) { bestn = n; bi = a; }
  }
  const double tot = w_sum_d(part);
  w_argmax(bestn, bi);
  w_sync();
  W_FOR(a, d.A) if (tot > 0.0) pi[a] = pi[a] / tot;
  W_LANE0 {
    const double* nw = E.root_nw + (size_t)g * 2;
    double q = 0.0;
    if (nw[0] > 0.0) q = ti[TI_ROOT_FRESH] ? nw[1] / nw[0] : (double)f_div((float)nw[1], (float)nw[0]);
    E.res_q[(size_t)g * 2] = q;
    E.res_move[g] = bi;
  }
  w_sync();
}

// One leaf-collection pass for slot g (mcts_v2.py:572-611 / :380-411).
AZ_DEV void game_collect(const AzState& E, int g, Sim& S, LocalCounters& lc) {
  const AzDims& d = E.d;
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  const int st = ti[TI_STATE];
  int nleaves = 0;
  if (ti[TI_ACTIVE]) {
    if (st == ST_NEED_ROOT) {
      sim_load(E, g, S);
      sim_write_obs(d, S, E.leaf_obs + (size_t)g * d.Pmax * d.obs_bytes);
      W_LANE0 { E.leaf_node[(size_t)g * d.Pmax] = 0; E.leaf_depth[(size_t)g * d.Pmax] = 0; }
      nleaves = 1;
    } else if (st == ST_SEARCHING) {
      TreeView T = tree_view(E, g, ti[TI_BUF]);
      int n_nodes = ti[TI_NODES];
      int tries = 0;
      if (d.game == 0) {  // group labels of the root position: one full labelling per pass, copied into every descent
        sim_load(E, g, S);
        go_label(d, S, false);
        W_FOR(c, d.nc) S.root_label[c] = S.label[c];
        w_sync();
      }
      while (nleaves < E.s.P && tries < E.s.tries) {
        tries++;
        sim_load(E, g, S);
        if (d.game == 0) {
          W_FOR(c, d.nc) S.label[c] = S.root_label[c];
          S.labels_valid = 1;
          w_sync();
        }
        int node = 0, depth = 0;
        float n_cur = 0.f;
        bool expanded_cur = true;  // the root of a running search is always expanded
        StepOut o;
        o.done = 0; o.reward_x2 = 0; o.winner = 0; o.captured = 0; o.score = 0.f;
        const uint8_t* legal = E.root_legal + (size_t)g * d.Ap;
        bool aborted = false;
        while (expanded_cur) {
          int link;
          float n_child;
          const int a = tree_pick(E, T, g, node, legal, n_cur, link, n_child);
          int child;
          bool child_expanded = false;
          if (link < 0) {
            child = tree_new_node(E, T, node, a, -S.to_play, n_nodes, lc);
            if (child < 0) { aborted = true; break; }
            n_child = 0.f;
          } else {
            child = link & AZ_CIDX_MASK;
            child_expanded = (link & AZ_CIDX_EXPANDED) != 0;
          }
          if (depth < AZ_PATH) {
            W_LANE0 { S.path_k[depth] = node * d.Ap + a; S.path_n[depth] = (int16_t)child; }
          }
          node = child;
          n_cur = n_child;
          depth++;
          o = sim_play(d, S, a);
          legal = S.legal;
          if (o.done) break;
          expanded_cur = child_expanded;
        }
        if (aborted) break;
        w_sync();
        lc.descents++;
        lc.depth += depth;
        if (o.done) {  // terminal: never expanded, back up the game result (mcts_v2.py:604-608)
          const float v = -(0.5f * (float)o.reward_x2);
          if (depth <= AZ_PATH) path_backup(E, T, g, S.path_k, depth, v, lc);
          else tree_backup(E, T, g, node, v, lc);
          continue;
        }
        if (E.s.use_vloss) {
          if (depth <= AZ_PATH) path_vloss(E, T, g, S.path_k, S.path_n, depth, +1);
          else tree_vloss(E, T, g, node, +1);
        }
        const size_t row = (size_t)g * d.Pmax + nleaves;
        W_LANE0 { E.leaf_node[row] = (int16_t)node; E.leaf_depth[row] = depth; }
        if (depth <= AZ_PATH) {
          W_FOR(i, depth) { E.leaf_pk[row * AZ_PATH + i] = S.path_k[i]; E.leaf_pn[row * AZ_PATH + i] = S.path_n[i]; }
        }
        sim_write_obs(d, S, E.leaf_obs + row * d.obs_bytes);
        nleaves++;
      }
      W_LANE0 ti[TI_NODES] = n_nodes;
    }
  }
  W_LANE0 ti[TI_NLEAVES] = nleaves;
  lc.evals += nleaves;
  w_sync();
}

// Observation of a leaf at `depth` in node-cache mode: the k-th most recent board is the scratch board (k = 0), the cached
// board of the path node at depth-k, or - past the root - the slot's own history (envs/base.py:243-261).
AZ_DEV void leaf_write_obs(const AzState& E, const TreeView& T, int g, const Sim& S, int node, int depth, int8_t* out) {
  const AzDims& d = E.d;
  // the num_stack source boards first (scratch, cached boards of the path nodes, the slot's history: cold global lines), all loads
  // of a chunk of cells in flight together, then the 2 * num_stack + 1 planes
  const int8_t* src[8];
  int anc = node;  // node at depth - k while walking up (only used when the path was too deep to be recorded)
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = depth - k;
    src[k] = S.board;
    if (k > 0 && k < d.num_stack) {
      if (j >= 1) {
        if (depth <= AZ_PATH) anc = S.path_n[j - 1];
        else anc = T.parent[anc];
        src[k] = T.nboard + (size_t)anc * d.ncp;
      } else src[k] = E.hist + ((size_t)g * 8 + (size_t)(-j)) * d.ncp;
    }
  }
  const int8_t me = (int8_t)S.to_play, opp = (int8_t)(-S.to_play);
  const int8_t colour = (S.to_play == 1);
  for (int c = AZ_LANE; c < d.nc; c += AZ_WIDTH) {
    int8_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = k < d.num_stack ? src[k][c] : (int8_t)0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < d.num_stack) {
        out[(size_t)(2 * k) * d.nc + c] = (v[k] == me);
        out[(size_t)(2 * k + 1) * d.nc + c] = (v[k] == opp);
      }
    }
    out[(size_t)(2 * d.num_stack) * d.nc + c] = colour;
  }
  w_sync();
}

// game_collect with the per-node position cache (d.node_cache): interior levels of a descent only pick (the legal mask of an
// expanded node is read from the cache); the game is advanced by ONE ply per descent, from the cached position of the leaf's
// parent, and the new node's position is stored for the descents that will pass through it.  Same results as game_collect:
// a node's position is a function of the path, so reading it back equals replaying it.
AZ_DEV void game_collect_nc(const AzState& E, int g, Sim& S, LocalCounters& lc) {
  const AzDims& d = E.d;
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  const int st = ti[TI_STATE];
  int nleaves = 0;
  if (ti[TI_ACTIVE]) {
    if (st == ST_NEED_ROOT) {
      sim_load(E, g, S);
      sim_write_obs(d, S, E.leaf_obs + (size_t)g * d.Pmax * d.obs_bytes);
      W_LANE0 { E.leaf_node[(size_t)g * d.Pmax] = 0; E.leaf_depth[(size_t)g * d.Pmax] = 0; }
      nleaves = 1;
    } else if (st == ST_SEARCHING) {
      TreeView T = tree_view(E, g, ti[TI_BUF]);
      int n_nodes = ti[TI_NODES];
      int tries = 0;
      const int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
      const int r_to_play = ei[EI_TO_PLAY], r_steps = ei[EI_STEPS], r_h1 = ei[EI_H1], r_h2 = ei[EI_H2], r_ko = ei[EI_KO];
      bool have_root_labels = false;
      while (nleaves < E.s.P && tries < E.s.tries) {
        tries++;
        int node = 0, parent = 0, depth = 0, a = -1;
        int prev1 = r_h1, prev2 = r_h2;  // the two moves before `node` (pass detection, go.py:176-192)
        float n_cur = 0.f;
        const uint8_t* legal = E.root_legal + (size_t)g * d.Ap;
        bool aborted = false;
        for (;;) {
          int link;
          float n_child;
          a = tree_pick(E, T, g, node, legal, n_cur, link, n_child);
          int child;
          bool child_expanded = false;
          if (link < 0) {
            child = tree_new_node(E, T, node, a, ((depth + 1) & 1) ? -r_to_play : r_to_play, n_nodes, lc);
            if (child < 0) { aborted = true; break; }
            n_child = 0.f;
          } else {
            child = link & AZ_CIDX_MASK;
            child_expanded = (link & AZ_CIDX_EXPANDED) != 0;
          }
          if (depth < AZ_PATH) {
            W_LANE0 { S.path_k[depth] = node * d.Ap + a; S.path_n[depth] = (int16_t)child; }
          }
          parent = node;
          node = child;
          n_cur = n_child;
          depth++;
          if (!child_expanded) break;
          prev2 = prev1;
          prev1 = a;
          legal = T.nlegal + (size_t)child * d.Ap;
        }
        if (aborted) break;
        w_sync();
        {
          const int tk = T.nko[node];
          if (tk >= AZ_TERM_BASE - 2) {  // a terminal found by an earlier descent: same backup, no board work
            lc.descents++;
            lc.depth += depth;
            const float v = -(0.5f * (float)(tk - AZ_TERM_BASE));
            if (depth <= AZ_PATH) path_backup(E, T, g, S.path_k, depth, v, lc);
            else tree_backup(E, T, g, node, v, lc);
            continue;
          }
        }
        // ---- the position of `parent`, then the one ply that leads to the leaf
        if (parent == 0) {
          sim_load(E, g, S);
          if (d.game == 0) {  // group labels of the root position: one full labelling per pass
            if (!have_root_labels) {
              go_label(d, S, false);
              W_FOR(c, d.nc) S.root_label[c] = S.label[c];
              have_root_labels = true;
            } else {
              W_FOR(c, d.nc) S.label[c] = S.root_label[c];
            }
            S.labels_valid = 1;
            w_sync();
          }
        } else {
          const int8_t* pb = T.nboard + (size_t)parent * d.ncp;
          W_FOR(c, d.ncp) S.board[c] = pb[c];
          S.to_play = ((depth - 1) & 1) ? -r_to_play : r_to_play;
          S.steps = r_steps + depth - 1;
          S.h1 = prev1;
          S.h2 = prev2;
          S.ko = T.nko[parent];
          S.caps_b = S.caps_w = 0;  // capture totals are not part of the observation; the real game keeps its own (env_step)
          S.head = 0;
          S.labels_valid = 0;
          S.libs_valid = 0;
          w_sync();
        }
        const StepOut o = sim_play(d, S, a);
        lc.descents++;
        lc.depth += depth;
        if (o.done) {  // terminal: never expanded, back up the game result (mcts_v2.py:604-608)
          const float v = -(0.5f * (float)o.reward_x2);
          W_LANE0 T.nko[node] = (int16_t)(AZ_TERM_BASE + o.reward_x2);
          if (depth <= AZ_PATH) path_backup(E, T, g, S.path_k, depth, v, lc);
          else tree_backup(E, T, g, node, v, lc);
          continue;
        }
        {  // remember the leaf's position for the descents that will go through it
          int8_t* nb = T.nboard + (size_t)node * d.ncp;
          W_FOR(c, d.ncp) nb[c] = S.board[c];
          uint8_t* nl = T.nlegal + (size_t)node * d.Ap;
          W_FOR(k, d.Ap) nl[k] = S.legal[k];
          W_LANE0 T.nko[node] = (int16_t)S.ko;
        }
        if (E.s.use_vloss) {
          if (depth <= AZ_PATH) path_vloss(E, T, g, S.path_k, S.path_n, depth, +1);
          else tree_vloss(E, T, g, node, +1);
        }
        const size_t row = (size_t)g * d.Pmax + nleaves;
        W_LANE0 { E.leaf_node[row] = (int16_t)node; E.leaf_depth[row] = depth; }
        if (depth <= AZ_PATH) {
          W_FOR(i, depth) { E.leaf_pk[row * AZ_PATH + i] = S.path_k[i]; E.leaf_pn[row * AZ_PATH + i] = S.path_n[i]; }
        }
        leaf_write_obs(E, T, g, S, node, depth, E.leaf_obs + row * d.obs_bytes);
        nleaves++;
      }
      W_LANE0 ti[TI_NODES] = n_nodes;
    }
  }
  W_LANE0 ti[TI_NLEAVES] = nleaves;
  lc.evals += nleaves;
  w_sync();
}

// Consume the evaluator's output for slot g (mcts_v2.py:364-368, :613-625).
AZ_DEV void game_apply(const AzState& E, int g, LocalCounters& lc) {
  const AzDims& d = E.d;
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  if (!ti[TI_ACTIVE]) return;
  int st = ti[TI_STATE];
  TreeView T = tree_view(E, g, ti[TI_BUF]);
  const float* pri = E.priors + (size_t)g * d.Pmax * d.Ap;
  const float* val = E.values + (size_t)g * d.Pmax;
  if (st == ST_NEED_ROOT) {
    if (ti[TI_NLEAVES] < 1) return;  // no root evaluation collected yet (freshly begun / restarted slot in a fused apply-first pass)
    tree_init_node(E, T, 0, -1, -1, E.env_i[(size_t)g * ENV_INTS + EI_TO_PLAY]);
    W_LANE0 {
      ti[TI_NODES] = 1;
      ti[TI_ROOT_FRESH] = 1;
      ti[TI_ROOT_NOISED] = 0;
      E.root_nw[(size_t)g * 2] = 0.0;
      E.root_nw[(size_t)g * 2 + 1] = 0.0;
    }
    w_sync();
    tree_expand(E, T, 0, pri);
    tree_backup(E, T, g, 0, val[0], lc);
    lc.nodes++;
    search_enter(E, g);
    return;
  }
  if (st != ST_SEARCHING) return;
  const int n = ti[TI_NLEAVES];
  for (int j = 0; j < n; ++j) {
    const size_t row = (size_t)g * d.Pmax + j;
    const int node = E.leaf_node[row];
    const int depth = E.leaf_depth[row];
    const bool fast = depth <= AZ_PATH;
    const int32_t* pk = E.leaf_pk + row * AZ_PATH;
    const int16_t* pn = E.leaf_pn + row * AZ_PATH;
    if (E.s.use_vloss) {
      if (fast) path_vloss(E, T, g, pk, pn, depth, -1);
      else tree_vloss(E, T, g, node, -1);
    }
    if (T.expanded[node]) continue;  // same leaf picked twice in this batch (mcts_v2.py:619-622)
    tree_expand(E, T, node, pri + (size_t)j * d.Ap);
    if (fast) path_backup(E, T, g, pk, depth, val[j], lc);
    else tree_backup(E, T, g, node, val[j], lc);
  }
  if (E.root_nw[(size_t)g * 2] >= (double)E.s.sims_bound) {
    search_finish(E, g);
    W_LANE0 ti[TI_STATE] = ST_DONE;
    w_sync();
  }
}

// Re-root on `move` (mcts_v2.py:643-653): the subtree of the chosen child is compacted breadth-first
// into the slot's other node pool; siblings are dropped.  Returns 1 if a subtree was kept.
// payload of one kept node: old index `old` of the source pool -> index `i` of the destination pool; all loads before the first store
AZ_DEV void commit_copy_node(const TreeView& To, TreeView& Tn, size_t old, int i, int Ap, int ncp) {
  constexpr int KC = 3;
  const bool cache = To.nboard != nullptr;
  const uint8_t ex = To.expanded[old];
  const int8_t tp = To.to_play[old];
  const int16_t ko = cache ? To.nko[old] : (int16_t)-1;
  for (int a00 = 0; a00 < Ap; a00 += KC * AZ_WIDTH) {
    float fn[KC], fw[KC], fp[KC];
    uint8_t lg[KC];
    int8_t bd[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int a = a00 + k * AZ_WIDTH + AZ_LANE;
      fn[k] = fw[k] = fp[k] = 0.f; lg[k] = 0; bd[k] = 0;
      if (a < Ap) {
        fn[k] = To.N[old * Ap + a];
        fw[k] = To.W[old * Ap + a];
        fp[k] = To.P[old * Ap + a];
        if (cache) {
          lg[k] = To.nlegal[old * Ap + a];
          if (a < ncp) bd[k] = To.nboard[old * ncp + a];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int a = a00 + k * AZ_WIDTH + AZ_LANE;
      if (a < Ap) {
        Tn.N[(size_t)i * Ap + a] = fn[k];
        Tn.W[(size_t)i * Ap + a] = fw[k];
        Tn.P[(size_t)i * Ap + a] = fp[k];
        if (cache) {
          Tn.nlegal[(size_t)i * Ap + a] = lg[k];
          if (a < ncp) Tn.nboard[(size_t)i * ncp + a] = bd[k];
        }
      }
    }
  }
  W_LANE0 {
    Tn.expanded[i] = ex;
    Tn.to_play[i] = tp;
    Tn.vloss[i] = 0;
    if (cache) Tn.nko[i] = ko;
  }
}

AZ_DEV void commit_prefetch_node(const TreeView& To, size_t pf, int Ap, int ncp) {
  w_prefetch(To.N + pf * Ap, Ap * 4);
  w_prefetch(To.W + pf * Ap, Ap * 4);
  w_prefetch(To.P + pf * Ap, Ap * 4);
  if (To.nboard != nullptr) { w_prefetch(To.nboard + pf * ncp, ncp); w_prefetch(To.nlegal + pf * Ap, Ap); }
}

// Deferred payload of the re-roots of one k_advance launch (jobs in E.rr_jobs[0 .. *E.rr_count)): `per` consecutive warps of the
// grid share one job and stride over its nodes 1 .. count-1 (node 0 was copied by the advancing warp).  TI_BUF already names the
// destination pool; the old indices are still in the slot's remap table.
AZ_DEV void reroot_payload(const AzState& E, int warp_global, int total_warps) {
  const int n_jobs = *E.rr_count;
  if (n_jobs <= 0) return;
  const int per = total_warps / n_jobs > 1 ? total_warps / n_jobs : 1;
  const int lanes_of_jobs = total_warps / per;  // jobs served per sweep
  const int r = warp_global % per;
  for (int j = warp_global / per; j < n_jobs; j += lanes_of_jobs) {
    const int g = E.rr_jobs[j];
    const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
    const int buf = ti[TI_BUF], count = ti[TI_NODES];
    TreeView To = tree_view(E, g, buf ^ 1);
    TreeView Tn = tree_view(E, g, buf);
    const int16_t* remap = E.remap + (size_t)g * E.d.cap;
    for (int i = 1 + r; i < count; i += per) commit_copy_node(To, Tn, (size_t)remap[i], i, E.d.Ap, E.d.ncp);
  }
}

AZ_DEV int game_commit(const AzState& E, int g, int move, double* best_child_q, bool defer = false) {
  const AzDims& d = E.d;
  const int Ap = d.Ap;
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  const int buf = ti[TI_BUF];
  TreeView To = tree_view(E, g, buf);
  const int link0 = (move >= 0 && move < d.A && ti[TI_NODES] > 0) ? To.cidx[move] : -1;
  const int child = link0 < 0 ? -1 : (link0 & AZ_CIDX_MASK);
  double bq = 0.0;
  int kept = 0;
  if (child >= 0) {
    const float n_c = To.N[move], w_c = To.W[move];
    bq = (double)(-(n_c > 0.f ? f_div(w_c, n_c) : 0.0f));
    TreeView Tn = tree_view(E, g, buf ^ 1);
    int16_t* remap = E.remap + (size_t)g * d.cap;
    W_LANE0 remap[0] = (int16_t)child;
    w_sync();
    int count = 1;
    constexpr int KC = 3;  // chunks of AZ_WIDTH entries fetched together: one round of loads per node at 9x9 (Ap = 96)
    // ---- pass 1, structure: breadth-first numbering.  Only the link rows are read; a node's links are all requested before the
    //      first one is consumed, and the links of a node further down the queue are prefetched, so the dependent chain per
    //      node is one (mostly L2) round trip instead of one per 32 actions.
    for (int i = 0; i < count; ++i) {
      const int old = remap[i];
      if (i + 4 < count) w_prefetch(To.cidx + (size_t)remap[i + 4] * Ap, Ap * 2);
      int base = count;
      for (int a00 = 0; a00 < Ap; a00 += KC * AZ_WIDTH) {
        int c[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int a = a00 + k * AZ_WIDTH + AZ_LANE;
          c[k] = a < Ap ? (int)To.cidx[(size_t)old * Ap + a] : -1;
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int a = a00 + k * AZ_WIDTH + AZ_LANE;
          if (a00 + k * AZ_WIDTH < Ap) {  // uniform across the warp (Ap is a multiple of the warp width)
            const bool has = c[k] >= 0;
            const uint32_t m = w_ballot(has);
            const int ni = base + az_popc(m & w_lanemask_lt());
            if (has) { remap[ni] = (int16_t)(c[k] & AZ_CIDX_MASK); Tn.parent[ni] = (int16_t)i; Tn.pmove[ni] = (int16_t)a; }
            Tn.cidx[(size_t)i * Ap + a] = has ? (int16_t)(ni | (c[k] & AZ_CIDX_EXPANDED)) : (int16_t)-1;
            base += az_popc(m);
          }
        }
      }
      count = base;
      w_sync();
    }
    W_LANE0 { Tn.parent[0] = -1; Tn.pmove[0] = -1; }
    // ---- pass 2, payload: statistics rows, cached position and flags of every kept node.  No node depends on another one here.
    //      With `defer` only the new root is copied by this warp (search_enter needs its rows right away); nodes 1.. are left to
    //      k_reroot_payload, which spreads them over the whole grid (az_engine.cu launches it right behind k_advance).
    const int n_inline = defer ? 1 : count;
    for (int i = 0; i < n_inline; ++i) {
      if (i + 4 < n_inline) commit_prefetch_node(To, (size_t)remap[i + 4], Ap, d.ncp);
      commit_copy_node(To, Tn, (size_t)remap[i], i, Ap, d.ncp);
    }
    if (defer && count > 1) {
      W_LANE0 {
        const int j = atomic_add_i(E.rr_count, 1);
        E.rr_jobs[j] = g;
      }
    }
    w_sync();
    W_LANE0 {
      ti[TI_BUF] = buf ^ 1;
      ti[TI_NODES] = count;
      ti[TI_ROOT_FRESH] = 0;   // carried N / W are np.float32 scalars (mcts_v2.py:439-443)
      ti[TI_ROOT_NOISED] = 0;
      E.root_nw[(size_t)g * 2] = (double)n_c;
      E.root_nw[(size_t)g * 2 + 1] = (double)w_c;
    }
    kept = 1;
  } else {
    W_LANE0 ti[TI_NODES] = 0;
  }
  W_LANE0 ti[TI_STATE] = ST_IDLE;
  w_sync();
  if (best_child_q) *best_child_q = bq;
  return kept;
}

// ---- device-resident self-play: what play_and_record_one_game does between two searches ---------
AZ_DEV void game_new(const AzState& E, int g, Sim& S) {
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  env_reset(E, g, S);
  W_LANE0 {
    // game id = (games started in this slot) * G + slot: unique, and a function of (seed, slot, history) only, so a run is
    // reproducible no matter in which order the warps get here
    const unsigned long long uid = (unsigned long long)ti[TI_SLOT_GAMES] * (unsigned long long)E.d.G + (unsigned long long)g;
    ti[TI_SLOT_GAMES] += 1;
    ti[TI_GAME_UID] = (int)uid;
    ti[TI_GAME_PLY] = 0;
    ti[TI_MARKED] = 0;
    ti[TI_NODES] = 0;
    ti[TI_NLEAVES] = 0;  // leaves of the previous game (already consumed, or abandoned by a restart) are not this game's
    ti[TI_STATE] = ST_NEED_ROOT;
    ti[TI_WARM] = (0 <= E.s.warm_up_steps) ? 1 : 0;
    // resign lottery (pipeline.py:244-246)
    int disabled = 1;
    if (E.d.game == 0 && E.s.resign_threshold > -1.0f) {
      const double u = az_u01(az_rand64(E.s.seed, 0x5e5160ull + (uint64_t)g, uid));
      if (u > (double)E.s.disable_resign_ratio) disabled = 0;
    }
    ti[TI_RESIGN_DISABLED] = disabled;
  }
  w_sync();
}

AZ_DEV void game_advance(const AzState& E, int g, Sim& S, bool defer_payload = false) {
  const AzDims& d = E.d;
  int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  if (!ti[TI_ACTIVE] || ti[TI_STATE] != ST_DONE) return;
  int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
  const double* pi = E.res_pi + (size_t)g * d.Ap;
  const uint8_t* legal = E.root_legal + (size_t)g * d.Ap;
  const bool warm = ti[TI_WARM] != 0;
  // ---- move choice (mcts_v2.py:630-641): rejection of pass-in-warm-up / illegal == renormalised draw
  int move = E.res_move[g];
  if (!E.s.deterministic) {
    int pick = -1;
    W_LANE0 {
      double tot = 0.0;
      for (int a = 0; a < d.A; ++a)
        if (legal[a] == 1 && !(warm && a == d.pass_move)) tot += pi[a];
      if (tot > 0.0) {
        const double u = az_u01(az_rand64(E.s.seed, 0x30fe0000ull + (uint64_t)g, ((uint64_t)ti[TI_GAME_UID] << 12) + ei[EI_STEPS])) * tot;
        double acc = 0.0;
        for (int a = 0; a < d.A; ++a) {
          if (legal[a] == 1 && !(warm && a == d.pass_move) && pi[a] > 0.0) { acc += pi[a]; pick = a; if (acc > u) break; }
        }
      }
      E.res_move[g] = pick;
    }
    w_sync();
    pick = E.res_move[g];
    if (pick >= 0) move = pick;
    else if (d.pass_move >= 0) move = d.pass_move;  // the reference would spin forever here
  }
  // ---- record (state, pi, to_play) (pipeline.py:323-326)
  const int ply = ti[TI_GAME_PLY];
  if (ply < d.max_len) {
    sim_load(E, g, S);
    sim_write_obs(d, S, E.g_obs + ((size_t)g * d.max_len + ply) * d.obs_bytes);
    float* gp = E.g_pi + ((size_t)g * d.max_len + ply) * d.A;
    W_FOR(a, d.A) gp[a] = (float)pi[a];
    W_LANE0 E.g_to_play[(size_t)g * d.max_len + ply] = (int8_t)ei[EI_TO_PLAY];
  }
  // ---- resignation (pipeline.py:328-341)
  const TreeView T = tree_view(E, g, ti[TI_BUF]);
  const int kid = T.cidx[move];
  const float n_c = T.N[move], w_c = T.W[move];
  const double child_q = kid >= 0 ? (double)(-(n_c > 0.f ? f_div(w_c, n_c) : 0.0f)) : 0.0;
  const double root_q = E.res_q[(size_t)g * 2];
  int play = move;
  if (d.game == 0 && ei[EI_STEPS] > E.s.check_resign_after && root_q < (double)E.s.resign_threshold &&
      child_q < (double)E.s.resign_threshold) {
    W_LANE0 if (ti[TI_MARKED] == 0) ti[TI_MARKED] = ei[EI_TO_PLAY];
    if (!ti[TI_RESIGN_DISABLED]) play = -1;
  }
  W_LANE0 if (ply < d.max_len) E.g_move[(size_t)g * d.max_len + ply] = (int16_t)play;
  w_sync();
  const StepOut o = env_step(E, g, S, play);
  W_LANE0 {
    ti[TI_GAME_PLY] = ply + 1;
    atomic_add_u64(&E.counters[CT_MOVES], 1ull);
  }
  w_sync();
  if (!o.done) {
    W_LANE0 ti[TI_WARM] = (ei[EI_STEPS] <= E.s.warm_up_steps) ? 1 : 0;
    w_sync();
    // evaluation matches search every move from a fresh root (`root_node=None`, pipeline.py:834-840): the tree is dropped
    const int kept = E.s.match ? 0 : game_commit(E, g, play, nullptr, defer_payload);
    if (kept) {
      W_LANE0 ti[TI_STATE] = ST_SEARCH_INIT;
      w_sync();
      search_enter(E, g);
    } else {
      W_LANE0 ti[TI_STATE] = ST_NEED_ROOT;
      w_sync();
    }
    return;
  }
  // ---- game over: z (pipeline.py:349-354), emit samples + game record, recycle the slot
  const int len = (ply + 1 < d.max_len) ? ply + 1 : d.max_len;
  const float reward = 0.5f * (float)o.reward_x2;
  const int last_player = ei[EI_LAST_PLAYER];
  unsigned long long head = 0;
  W_LANE0 {
    head = atomic_add_u64(&E.counters[CT_RING_HEAD], (unsigned long long)len);
    E.res_q[(size_t)g * 2 + 1] = (double)head;
  }
  w_sync();
  head = (unsigned long long)E.res_q[(size_t)g * 2 + 1];
  {
    // the samples of one game are contiguous in the record and in the ring (modulo wrap): at most two flat copies per array
    const size_t s0 = (size_t)(head % (unsigned long long)d.ring_cap);
    const size_t n1 = (size_t)len < (size_t)d.ring_cap - s0 ? (size_t)len : (size_t)d.ring_cap - s0, n2 = (size_t)len - n1;
    const int8_t* so = E.g_obs + (size_t)g * d.max_len * d.obs_bytes;
    const float* sp = E.g_pi + (size_t)g * d.max_len * d.A;
    w_copy_bytes(E.r_obs + s0 * d.obs_bytes, so, n1 * d.obs_bytes);
    w_copy_words(E.r_pi + s0 * d.A, sp, n1 * d.A);
    if (n2) {
      w_copy_bytes(E.r_obs, so + n1 * d.obs_bytes, n2 * d.obs_bytes);
      w_copy_words(E.r_pi, sp + n1 * d.A, n2 * d.A);
    }
    W_FOR(i, len) {
      const size_t slot = (size_t)((head + i) % (unsigned long long)d.ring_cap);
      float z = 0.f;
      if (reward != 0.f) z = (E.g_to_play[(size_t)g * d.max_len + i] == last_player) ? reward : -reward;
      E.r_z[slot] = z;
      E.r_move[slot] = E.g_move[(size_t)g * d.max_len + i];
    }
  }
  W_LANE0 {
    const unsigned long long gi = atomic_add_u64(&E.counters[CT_GAMES_HEAD], 1ull);
    int32_t* gr = E.games_ring + (size_t)(gi % AZ_GAMES_RING) * GR_INTS;
    float sc = o.score;
    gr[GR_SLOT] = g;
    gr[GR_LEN] = len;
    gr[GR_WINNER] = o.winner;
    gr[GR_BY_RESIGN] = play < 0 ? 1 : 0;
    gr[GR_PASSES] = ei[EI_NUM_PASSES];
    const int disabled = ti[TI_RESIGN_DISABLED], marked = ti[TI_MARKED];
    const int is_marked = (d.game == 0 && disabled && marked != 0) ? 1 : 0;
    gr[GR_RESIGN_DISABLED] = disabled;
    gr[GR_MARKED_FOR_RESIGN] = is_marked;
    gr[GR_COULD_WON] = (is_marked && o.winner == marked) ? 1 : 0;
    gr[GR_MARKED_PLAYER] = marked;
    gr[GR_FIRST_LO] = (int32_t)(uint32_t)(head & 0xffffffffull);
    gr[GR_FIRST_HI] = (int32_t)(uint32_t)(head >> 32);
    gr[GR_UID] = ti[TI_GAME_UID];
    memcpy(&gr[GR_SCORE_BITS], &sc, 4);
    atomic_add_u64(&E.counters[CT_GAMES], 1ull);
    atomic_add_u64(&E.counters[CT_SAMPLES], (unsigned long long)len);
  }
  w_sync();
  if (E.s.match && ti[TI_SLOT_GAMES] >= E.s.match_games) {  // the slot has played its games: retire it
    W_LANE0 { ti[TI_ACTIVE] = 0; ti[TI_STATE] = ST_IDLE; ti[TI_NLEAVES] = 0; }
    w_sync();
    return;
  }
  game_new(E, g, S);
}

// Which weight set evaluates the leaves of slot g in the match loop: the one that plays the colour to move at the ROOT (each
// player searches with its own network, also below opponent nodes of its tree: pipeline.py:826-840).
AZ_DEV int match_net_of(const AzState& E, int g) {
  const int32_t* ti = E.tree_i + (size_t)g * TREE_INTS;
  int black_net = E.slot_net[g];
  if (E.s.match_alternate) black_net ^= (ti[TI_SLOT_GAMES] - 1) & 1;
  return E.env_i[(size_t)g * ENV_INTS + EI_TO_PLAY] == 1 ? black_net : 1 - black_net;
}
