// az_state.h — structure-of-arrays state of all game slots, resident in HBM.
//
// Layout (G slots; ncp = N*N rounded up to 32; Ap = actions rounded up to 32; cap = nodes per tree):
//   position   board[G][ncp] int8 (+1 black / -1 white / 0), hist[G][8][ncp] newest first,
//              env_i[G][ENV_INTS] scalars, root_legal[G][Ap] uint8
//   tree       two node pools per slot (ping-pong for re-root compaction, mcts_v2.py:643-653):
//              cN / cW / cP [G][2][cap][Ap] float32 rows (stats of a node live in its PARENT's row,
//              mcts_v2.py:111-127), cidx [G][2][cap][Ap] int16 child links, per-node parent / move /
//              to_play / expanded / virtual-loss count; the root's own N / W (the reference's
//              DummyNode, mcts_v2.py:56-62) in root_nw[G][2] as doubles, root_p64[G][Ap] the float64
//              prior row of a noised root (mcts_v2.py:259-262)
//   leaves     leaf_node[G][Pmax], leaf_obs[G*Pmax][obs_bytes] int8, priors[G*Pmax][Ap] f32, values[G*Pmax]
//   samples    per-slot game record (obs, pi, to_play per ply) + a global ring of finished samples
#pragma once
#include <stdint.h>

enum {
  EI_TO_PLAY = 0,   // +1 black / -1 white (both games internally)
  EI_STEPS,
  EI_LAST_MOVE,     // -2 none, -1 resign
  EI_H1,            // last move in `history` (resign excluded, go.py:101-104); -2 none
  EI_H2,            // second last
  EI_KO,            // -1 none
  EI_DONE,
  EI_WINNER,        // +1 / -1 / 0
  EI_LAST_PLAYER,
  EI_BY_RESIGN,
  EI_CAPS_B,
  EI_CAPS_W,
  EI_NUM_PASSES,
  EI_REWARD_X2,     // last step's reward * 2 as int (exact for -1, 0, 1)
  EI_SCORE_BITS,    // float bits of black - (white + komi) when the game ended by score (Go)
  ENV_INTS = 16
};

enum {
  TI_STATE = 0,
  TI_BUF,           // which node pool is live
  TI_NODES,         // nodes in the live pool
  TI_ROOT_FRESH,    // 1: root N/W are Python floats (double arithmetic); 0: carried np.float32
  TI_ROOT_NOISED,
  TI_NLEAVES,
  TI_RNG_CTR,
  TI_WARM,          // warm_up flag of the running search
  TI_GAME_PLY,      // samples recorded for the current game
  TI_RESIGN_DISABLED,
  TI_MARKED,        // marked_resign_player (0 none)
  TI_GAME_UID,
  TI_ACTIVE,        // slot takes part in the current split-phase search batch
  TI_SLOT_GAMES,    // games started in this slot since az_selfplay_begin (makes the RNG streams independent of warp scheduling)
  TREE_INTS = 16
};

enum {  // per-slot search state machine
  ST_IDLE = 0,       // no search running (tree may hold a re-usable subtree)
  ST_NEED_ROOT = 1,  // root must be evaluated (fresh root, mcts_v2.py:364-368)
  ST_SEARCH_INIT = 2,// root ready; noise to be mixed in, then the loop bound is checked
  ST_SEARCHING = 3,
  ST_DONE = 4        // search finished; result readable, waiting for commit / advance
};

struct AzDims {
  int game, n, nc, ncp, A, Ap, num_stack, planes, obs_bytes;
  int max_steps, num_to_win;
  float komi;
  int G, cap, Pmax;
  int pass_move;     // nc for Go, -1 for Gomoku
  int table_len;     // length of the pb_c / sqrt tables
  int max_len;       // max plies per game (sample storage)
  int ring_cap;
  int node_cache;    // 1: every tree node keeps its position (board, legal mask, ko) so that a descent plays ONE ply, at the leaf
};

struct AzSearchCfg {
  int sims_bound;    // loop runs while root.N < sims_bound (mcts_v2.py:378 / :568)
  int P;             // leaves per batch
  int use_vloss;     // parallel_uct_search semantics
  int tries;         // attempts per batch: 2*P (mcts_v2.py:572) or 1 for the serial search
  int root_noise;
  int deterministic;
  int host_noise;    // 1: Dirichlet samples supplied by the host in `noise`
  int selfplay;      // 1: device-resident loop
  int warm_up_steps, check_resign_after;
  float resign_threshold, disable_resign_ratio;
  uint64_t seed;
  // evaluation matches (pipeline.py:815-867): two weight sets, a fresh tree every move, a fixed number of games per slot
  int match;            // 1: match loop (no subtree reuse, slot_net decides the evaluating network, slots retire after their games)
  int match_games;      // games per slot
  int match_alternate;  // 1: the networks swap colours from one game of a slot to the next
};

struct AzState {
  AzDims d;
  AzSearchCfg s;
  // position
  int8_t* board;
  int8_t* hist;
  int32_t* env_i;
  uint8_t* root_legal;
  // tree
  float *cN, *cW, *cP;
  int16_t* cidx;
  int16_t *parent, *pmove, *nvloss;
  int8_t* nto_play;
  uint8_t* expanded;
  int32_t* tree_i;
  double* root_nw;
  double* root_p64;
  double* noise;       // [G][Ap] host-supplied Dirichlet samples
  int16_t* remap;      // [G][cap]
  int32_t* rr_jobs;    // [G] slots whose re-root left nodes 1.. of the payload to k_reroot_payload
  int32_t* rr_count;   // [1] number of such jobs of the current k_advance launch
  // per-node position cache (d.node_cache): the position AFTER the move that leads to the node, written when the node is first
  // reached as a leaf; interior levels of a descent read the legal mask from here instead of replaying the game from the root
  int8_t* nboard;      // [G][2][cap][ncp]
  uint8_t* nlegal;     // [G][2][cap][Ap]
  int16_t* nko;        // [G][2][cap]
  const double* pbc_fresh;  // log((1+n+cb)/cb)+ci, double arithmetic (fresh root)
  const double* pbc_f32;    // same with the quotient rounded to float32 (re-used root, inner nodes)
  const double* sqrt_tab;   // sqrt(n)
  // leaves / evaluator interface
  int16_t* leaf_node;
  int8_t* leaf_obs;
  float* priors;
  float* values;
  int32_t* leaf_rows2;  // match loop: rows to be evaluated by the second weight set (leaf_total[4] of them)
  uint8_t* slot_net;    // match loop: [G] which weight set (0 / 1) plays BLACK in the slot's first game
  int32_t* leaf_rows;   // compacted list of occupied rows g*Pmax+j
  int32_t* leaf_total;  // [0] number of occupied rows, [1] active searches
  int32_t* leaf_count;  // [G] leaves per slot of the last collect pass
  int32_t* leaf_pk;     // [G*Pmax][AZ_PATH] flat (parent*Ap+move) indices of the leaf's path, root side first
  int16_t* leaf_pn;     // [G*Pmax][AZ_PATH] node ids along the path
  int32_t* leaf_depth;  // [G*Pmax]
  // search results
  double* res_pi;       // [G][Ap]
  double* res_q;        // [G][2] root_Q, best_child_Q
  int32_t* res_move;    // [G]
  // per-slot game record (self-play)
  int8_t* g_obs;        // [G][max_len][obs_bytes]
  float* g_pi;          // [G][max_len][A]
  int8_t* g_to_play;    // [G][max_len]
  int16_t* g_move;      // [G][max_len] move played at each ply (-1 resign)
  // finished-sample ring + finished-game ring
  int8_t* r_obs;
  float* r_pi;
  float* r_z;
  int16_t* r_move;
  int32_t* games_ring;  // [ring_games][GR_INTS]
  unsigned long long* counters;  // see CT_*
  // device-resident replay (learner input path): circular storage like replay.py:35-116
  int8_t* rp_obs;       // [rp_cap][obs_bytes]
  float* rp_pi;         // [rp_cap][A]
  float* rp_z;          // [rp_cap]
  int rp_cap;
};

enum { CT_SIMS = 0, CT_EVALS, CT_MOVES, CT_GAMES, CT_NODES, CT_DEPTH, CT_DESCENTS, CT_SAMPLES, CT_DROPPED, CT_ERRORS,
       CT_RING_HEAD, CT_GAMES_HEAD, CT_COUNT = 16 };

enum { GR_SLOT = 0, GR_LEN, GR_WINNER, GR_BY_RESIGN, GR_SCORE_BITS, GR_PASSES, GR_RESIGN_DISABLED, GR_MARKED_FOR_RESIGN,
       GR_COULD_WON, GR_MARKED_PLAYER, GR_FIRST_LO, GR_FIRST_HI, GR_UID, GR_INTS = 16 };
#define AZ_GAMES_RING 4096
#ifndef AZ_PATH
#define AZ_PATH 64          // recorded path length; deeper leaves fall back to the serial parent walk
#endif
#define AZ_CIDX_EXPANDED 0x4000  // child link flag: the child is already expanded (saves a dependent load per level)
#define AZ_CIDX_MASK 0x3FFF
