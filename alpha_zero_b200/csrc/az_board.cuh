// az_board.cuh — Go and Gomoku rules on one warp, working on a per-warp scratch position.
//
// Reference semantics (file:line under /root/reference/alpha_zero):
//   Go      envs/go.py:88-192 (step, game over), envs/go_engine.py:91-99 (is_koish), :386-441 (suicide,
//           legal mask), :460-507 (play_move: captures, ko), :123-152 + :509-516 (Tromp-Taylor score)
//   Gomoku  envs/gomoku.py:45-135 (step, win scan, game over)
//   obs     envs/base.py:228-266
// The reference keeps Python sets of stones / liberties per group; here groups are recomputed by
// min-label propagation over the board held in shared memory (every lane owns cells L, L+32, ...),
// liberties are counted with shared-memory atomics, and legality is evaluated per empty cell.
#pragma once
#include "az_state.h"
#include "az_warp.cuh"
#if defined(AZ_EMU) && defined(AZ_CHECK_LABELS)
#include <stdio.h>
#include <stdlib.h>
#endif

struct Sim {
  int8_t* board;    // [ncp]
  int8_t* ring;     // [8][ncp]  ring[(head+k)&7] = k-th most recent board
  int16_t* label;   // [ncp]
  int16_t* root_label; // [ncp] group labels of the slot's real position, computed once per kernel and copied per descent
  int32_t* aux;     // [ncp]     liberties per group label / border flags per empty region
  uint8_t* legal;   // [Ap]
  uint8_t* nbm;     // [ncp] neighbour mask per cell: bit0 down (+n), bit1 up (-n), bit2 right (+1), bit3 left (-1)
  int32_t* path_k;  // [AZ_PATH] flat stats index (parent*Ap+move) of every node on the current descent
  int16_t* path_n;  // [AZ_PATH] node ids
  int to_play, steps, h1, h2, ko, head;
  int labels_valid; // label[] describes the stone groups of the current board
  int libs_valid;   // aux[] holds their liberty counts
  int caps_b, caps_w;
};

AZ_DEV size_t sim_bytes(const AzDims& d) { return (size_t)d.ncp * 18 + d.Ap + AZ_PATH * 6; }

AZ_DEV void sim_carve(const AzDims& d, Sim& S, unsigned char* mem) {
  S.path_k = (int32_t*)mem; mem += AZ_PATH * 4;
  S.aux = (int32_t*)mem;  mem += (size_t)d.ncp * 4;
  S.label = (int16_t*)mem; mem += (size_t)d.ncp * 2;
  S.path_n = (int16_t*)mem; mem += AZ_PATH * 2;
  S.root_label = (int16_t*)mem; mem += (size_t)d.ncp * 2;
  S.board = (int8_t*)mem;  mem += d.ncp;
  S.ring = (int8_t*)mem;   mem += (size_t)d.ncp * 8;
  S.legal = (uint8_t*)mem; mem += d.Ap;
  S.nbm = (uint8_t*)mem;
  // which neighbours exist, once per kernel: keeps integer divisions by the board size out of the rule loops
  W_FOR(c, d.ncp) {
    uint8_t m = 0;
    if (c < d.nc) {
      const int r = c / d.n, k = c - r * d.n;
      m = (uint8_t)((r + 1 < d.n ? 1 : 0) | (r > 0 ? 2 : 0) | (k + 1 < d.n ? 4 : 0) | (k > 0 ? 8 : 0));
    }
    S.nbm[c] = m;
  }
  w_sync();
}

struct StepOut { int done; int reward_x2; int winner; int captured; float score; };

#define AZ_NEIGHBOURS(d, c, q, BODY)                                         \
  {                                                                         \
    const int _m = S.nbm[(c)];                                              \
    int q;                                                                  \
    if (_m & 1) { q = (c) + (d).n; BODY }                                   \
    if (_m & 2) { q = (c) - (d).n; BODY }                                   \
    if (_m & 4) { q = (c) + 1; BODY }                                       \
    if (_m & 8) { q = (c)-1; BODY }                                         \
  }

// Load the position of slot g (global) into the scratch.
AZ_DEV void sim_load(const AzState& E, int g, Sim& S) {
  const AzDims& d = E.d;
  const int8_t* b = E.board + (size_t)g * d.ncp;
  const int8_t* h = E.hist + (size_t)g * 8 * d.ncp;
  W_FOR(c, d.ncp) S.board[c] = b[c];
  for (int k = 0; k < d.num_stack; ++k) W_FOR(c, d.ncp) S.ring[k * d.ncp + c] = h[k * d.ncp + c];
  const int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
  S.to_play = ei[EI_TO_PLAY];
  S.steps = ei[EI_STEPS];
  S.h1 = ei[EI_H1];
  S.h2 = ei[EI_H2];
  S.ko = ei[EI_KO];
  S.caps_b = ei[EI_CAPS_B];
  S.caps_w = ei[EI_CAPS_W];
  S.head = 0;
  S.labels_valid = 0;
  S.libs_valid = 0;
  w_sync();
}

AZ_DEV void sim_push_board(const AzDims& d, Sim& S) {
  S.head = (S.head + 7) & 7;
  int8_t* dst = S.ring + S.head * d.ncp;
  W_FOR(c, d.ncp) dst[c] = S.board[c];
  w_sync();
}

// Observation planes [X_t, Y_t, X_t-1, ..., C] for the player to move (envs/base.py:243-259).
AZ_DEV void sim_write_obs(const AzDims& d, const Sim& S, int8_t* out) {
  for (int k = 0; k < d.num_stack; ++k) {
    const int8_t* b = S.ring + ((S.head + k) & 7) * d.ncp;
    int8_t* o0 = out + (size_t)(2 * k) * d.nc;
    int8_t* o1 = o0 + d.nc;
    W_FOR(c, d.nc) {
      int8_t v = b[c];
      o0[c] = (v == S.to_play);
      o1[c] = (v == -S.to_play);
    }
  }
  int8_t* oc = out + (size_t)(2 * d.num_stack) * d.nc;
  int8_t colour = (S.to_play == 1);
  W_FOR(c, d.nc) oc[c] = colour;
  w_sync();
}

// Min-label propagation with pointer jumping: label[c] = lowest cell index of the connected set of
// equal-valued cells containing c.  Stones always; empty regions only when with_empty.  Jacobi style (all
// lanes read label[], write aux[], then copy back) so that no lane reads a cell another lane is writing.
AZ_DEV void go_label(const AzDims& d, Sim& S, bool with_empty) {
  W_FOR(c, d.nc) S.label[c] = (S.board[c] != 0 || with_empty) ? (int16_t)c : (int16_t)-1;
  w_sync();
  bool changed;
  do {
    changed = false;
    W_FOR(c, d.nc) {
      const int l = S.label[c];
      int m = l;
      if (l >= 0) {
        const int8_t v = S.board[c];
        AZ_NEIGHBOURS(d, c, q, {
          if (S.board[q] == v) { const int lq = S.label[q]; if (lq < m) m = lq; }
        })
        const int lm = S.label[m];
        if (lm < m) m = lm;
        if (m < l) changed = true;
      }
      S.aux[c] = m;
    }
    w_sync();
    W_FOR(c, d.nc) S.label[c] = (int16_t)S.aux[c];
    w_sync();
  } while (w_any(changed));
}

// aux[label] = number of distinct empty points adjacent to the group.
AZ_DEV void go_count_liberties(const AzDims& d, Sim& S) {
  W_FOR(c, d.nc) S.aux[c] = 0;
  w_sync();
  W_FOR(e, d.nc) {
    if (S.board[e] != 0) continue;
    int seen[4];
    int ns = 0;
    AZ_NEIGHBOURS(d, e, q, {
      if (S.board[q] != 0) {
        int l = S.label[q];
        bool dup = false;
        for (int t = 0; t < ns; ++t) dup = dup || (seen[t] == l);
        if (!dup) { seen[ns++] = l; atomic_add_i(&S.aux[l], 1); }
      }
    })
  }
  S.libs_valid = 1;
  w_sync();
}

AZ_DEV bool go_is_suicide(const AzDims& d, const Sim& S, int c, int colour) {
  bool ok = false;
  AZ_NEIGHBOURS(d, c, q, {
    int8_t v = S.board[q];
    if (v == 0) ok = true;
    else if (v == colour) { if (S.aux[S.label[q]] >= 2) ok = true; }
    else { if (S.aux[S.label[q]] == 1) ok = true; }
  })
  return !ok;
}

// Legal mask for S.to_play (go_engine.py:417-441): empty, not the ko point, not suicide; pass always.
AZ_DEV void go_legal(const AzDims& d, Sim& S) {
  if (!S.labels_valid) {
    go_label(d, S, false);
    S.labels_valid = 1;
    S.libs_valid = 0;
  }
  if (!S.libs_valid) go_count_liberties(d, S);
  W_FOR(a, d.Ap) {
    uint8_t ok = 0;
    if (a < d.nc) ok = (S.board[a] == 0 && a != S.ko && !go_is_suicide(d, S, a, S.to_play)) ? 1 : 0;
    else if (a == d.nc) ok = 1;
    S.legal[a] = ok;
  }
  w_sync();
}

// Tromp-Taylor area score from black's point of view, komi included (go_engine.py:123-152, 509-516).
AZ_DEV float go_score(const AzDims& d, Sim& S) {
  go_label(d, S, true);
  S.labels_valid = 0;
  S.libs_valid = 0;
  W_FOR(c, d.nc) S.aux[c] = 0;
  w_sync();
  W_FOR(c, d.nc) {
    int8_t v = S.board[c];
    if (v == 0) continue;
    AZ_NEIGHBOURS(d, c, q, {
      if (S.board[q] == 0) atomic_or_i(&S.aux[S.label[q]], v > 0 ? 1 : 2);
    })
  }
  w_sync();
  int b = 0, w = 0;
  W_FOR(c, d.nc) {
    int8_t v = S.board[c];
    if (v > 0) b++;
    else if (v < 0) w++;
    else {
      int f = S.aux[S.label[c]];
      if (f == 1) b++;
      else if (f == 2) w++;
    }
  }
  b = w_sum_i(b);
  w = w_sum_i(w);
  w_sync();
  return (float)((double)b - ((double)w + (double)d.komi));
}

// One ply of Go on the scratch position (action in [0, nc]; nc = pass).  Leaves S.legal valid for the
// next mover unless the game ended.
AZ_DEV StepOut go_play(const AzDims& d, Sim& S, int action) {
  StepOut o;
  o.done = 0; o.reward_x2 = 0; o.winner = 0; o.captured = 0; o.score = 0.f;
  const int mover = S.to_play;
  if (action == d.nc) {
    S.ko = -1;  // go_engine.py:443-449
  } else {
    const int p = action;
    // is_koish BEFORE the stone is placed (go_engine.py:479): all neighbours one colour, none empty
    int first = 0;
    bool uniform = true;
    AZ_NEIGHBOURS(d, p, q, {
      int8_t v = S.board[q];
      if (v == 0) uniform = false;
      else if (first == 0) first = v;
      else if (v != first) uniform = false;
    })
    const int koish = uniform ? first : 0;
    w_sync();
    W_LANE0 S.board[p] = (int8_t)mover;
    w_sync();
    if (!S.labels_valid) {
      go_label(d, S, false);
    } else {
      // incremental: the new stone joins its friendly neighbour groups and the union keeps the smallest label
      // (labels are "lowest cell index of the group", so this is exactly what a full relabel would produce)
      int fl[4];
      int nf = 0, newl = p;
      AZ_NEIGHBOURS(d, p, q, {
        if (S.board[q] == mover) {
          const int l = S.label[q];
          bool dup = false;
          for (int t = 0; t < nf; ++t) dup = dup || (fl[t] == l);
          if (!dup) { fl[nf++] = l; if (l < newl) newl = l; }
        }
      })
      w_sync();
      if (nf == 0) {
        W_LANE0 S.label[p] = (int16_t)p;
      } else {
        W_FOR(c, d.nc) {
          if (c == p) S.label[c] = (int16_t)newl;
          else if (S.board[c] == mover) {
            const int l = S.label[c];
            bool hit = false;
            for (int t = 0; t < nf; ++t) hit = hit || (fl[t] == l);
            if (hit) S.label[c] = (int16_t)newl;
          }
        }
      }
      w_sync();
    }
    go_count_liberties(d, S);
    int capl[4];
    int ncapl = 0;
    AZ_NEIGHBOURS(d, p, q, {
      if (S.board[q] == -mover) {
        int l = S.label[q];
        if (S.aux[l] == 0) {
          bool dup = false;
          for (int t = 0; t < ncapl; ++t) dup = dup || (capl[t] == l);
          if (!dup) capl[ncapl++] = l;
        }
      }
    })
    int cnt = 0, capcell = -1;
    if (ncapl > 0) {
      w_sync();
      W_FOR(c, d.nc) {
        if (S.board[c] == -mover) {
          int l = S.label[c];
          bool hit = false;
          for (int t = 0; t < ncapl; ++t) hit = hit || (capl[t] == l);
          if (hit) { S.board[c] = 0; S.label[c] = -1; cnt++; capcell = c; }
        }
      }
      w_sync();
      cnt = w_sum_i(cnt);
      capcell = w_max_i(capcell);
      go_count_liberties(d, S);  // surviving groups keep their labels; only liberties changed
    }
    S.labels_valid = 1;
#if defined(AZ_EMU) && defined(AZ_CHECK_LABELS)
    {  // test-only cross-check of the incremental labels against a full relabel
      int16_t keep[512];
      memcpy(keep, S.label, (size_t)d.nc * 2);
      go_label(d, S, false);
      if (memcmp(keep, S.label, (size_t)d.nc * 2) != 0) { fprintf(stderr, "incremental group labels diverged\n"); abort(); }
      S.libs_valid = 0;  // go_label used aux[] as its scratch: the liberty counts have to be rebuilt before go_legal reads them
    }
#endif
    S.ko = (cnt == 1 && koish == -mover) ? capcell : -1;  // go_engine.py:491-494
    if (mover == 1) S.caps_b += cnt; else S.caps_w += cnt;
    o.captured = cnt;
  }
  S.to_play = -mover;
  S.steps += 1;
  S.h2 = S.h1;
  S.h1 = action;
  sim_push_board(d, S);
  // go.py:176-192 (resign handled by the caller)
  if (S.steps >= d.max_steps || (S.h1 == d.nc && S.h2 == d.nc)) {
    o.done = 1;
    float sc = go_score(d, S);
    o.score = sc;
    o.winner = sc > 0.f ? 1 : (sc < 0.f ? -1 : 0);
    o.reward_x2 = o.winner == 0 ? 0 : (o.winner == mover ? 2 : -2);  // go.py:152-156
  } else {
    go_legal(d, S);
  }
  return o;
}

// One ply of freestyle Gomoku (gomoku.py:45-135).  Stones are +1 / -1 internally (reference ids 1 / 2).
AZ_DEV StepOut gomoku_play(const AzDims& d, Sim& S, int action) {
  StepOut o;
  o.done = 0; o.reward_x2 = 0; o.winner = 0; o.captured = 0; o.score = 0.f;
  const int mover = S.to_play;
  w_sync();
  W_LANE0 S.board[action] = (int8_t)mover;
  w_sync();
  S.steps += 1;
  bool won = false;
  if (S.steps >= (d.num_to_win - 1) * 2) {  // gomoku.py:89
    const int r0 = action / d.n, c0 = action - r0 * d.n;
    const int dr[4] = {0, 1, 1, -1}, dc[4] = {1, 0, 1, 1};
    for (int k = 0; k < 4 && !won; ++k) {
      int run = 1;
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        int r = r0 + sgn * dr[k], c = c0 + sgn * dc[k];
        while (r >= 0 && r < d.n && c >= 0 && c < d.n && S.board[r * d.n + c] == mover) {
          run++; r += sgn * dr[k]; c += sgn * dc[k];
        }
      }
      won = run >= d.num_to_win;
    }
  }
  S.to_play = -mover;
  S.h2 = S.h1;
  S.h1 = action;
  sim_push_board(d, S);
  if (won) { o.done = 1; o.winner = mover; o.reward_x2 = 2; }
  else if (S.steps >= d.nc) { o.done = 1; }  // board full: draw (base.py:268, gomoku.py:130-135)
  W_FOR(a, d.Ap) S.legal[a] = (a < d.nc && S.board[a] == 0) ? 1 : 0;
  w_sync();
  return o;
}

AZ_DEV StepOut sim_play(const AzDims& d, Sim& S, int action) {
  return d.game == 0 ? go_play(d, S, action) : gomoku_play(d, S, action);
}

AZ_DEV void sim_legal(const AzDims& d, Sim& S) {
  if (d.game == 0) go_legal(d, S);
  else { W_FOR(a, d.Ap) S.legal[a] = (a < d.nc && S.board[a] == 0) ? 1 : 0; w_sync(); }
}

// ---- the real position of slot g ---------------------------------------------------------------
AZ_DEV void env_store(const AzState& E, int g, const Sim& S, bool shift_history) {
  const AzDims& d = E.d;
  int8_t* b = E.board + (size_t)g * d.ncp;
  int8_t* h = E.hist + (size_t)g * 8 * d.ncp;
  W_FOR(c, d.ncp) b[c] = S.board[c];
  if (shift_history) {
    for (int k = 0; k < 8; ++k) {
      const int8_t* src = S.ring + ((S.head + k) & 7) * d.ncp;
      W_FOR(c, d.ncp) h[k * d.ncp + c] = src[c];
    }
  }
  uint8_t* lg = E.root_legal + (size_t)g * d.Ap;
  W_FOR(a, d.Ap) lg[a] = S.legal[a];
  W_LANE0 {
    int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
    ei[EI_TO_PLAY] = S.to_play;
    ei[EI_STEPS] = S.steps;
    ei[EI_H1] = S.h1;
    ei[EI_H2] = S.h2;
    ei[EI_KO] = S.ko;
    ei[EI_CAPS_B] = S.caps_b;
    ei[EI_CAPS_W] = S.caps_w;
  }
  w_sync();
}

AZ_DEV void env_reset(const AzState& E, int g, Sim& S) {
  const AzDims& d = E.d;
  W_FOR(c, d.ncp) S.board[c] = 0;
  for (int k = 0; k < 8; ++k) W_FOR(c, d.ncp) S.ring[k * d.ncp + c] = 0;
  S.to_play = 1; S.steps = 0; S.h1 = -2; S.h2 = -2; S.ko = -1; S.head = 0; S.labels_valid = 0; S.libs_valid = 0;
  S.caps_b = S.caps_w = 0;
  w_sync();
  sim_legal(d, S);
  env_store(E, g, S, true);
  W_LANE0 {
    int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
    ei[EI_LAST_MOVE] = -2; ei[EI_DONE] = 0; ei[EI_WINNER] = 0; ei[EI_LAST_PLAYER] = 0; ei[EI_BY_RESIGN] = 0;
    ei[EI_NUM_PASSES] = 0; ei[EI_REWARD_X2] = 0;
  }
  w_sync();
}

// step() of the real game in slot g; action -1 = resign (go.py:103-119).  Caller has validated it.
AZ_DEV StepOut env_step(const AzState& E, int g, Sim& S, int action) {
  const AzDims& d = E.d;
  sim_load(E, g, S);
  StepOut o;
  const int mover = S.to_play;
  if (action < 0) {
    o.done = 1; o.reward_x2 = -2; o.winner = -mover; o.captured = 0; o.score = 0.f;
    S.ko = -1;
    S.to_play = -mover;
    S.steps += 1;
    sim_push_board(d, S);
  } else {
    o = sim_play(d, S, action);
  }
  if (o.done && d.game == 0) { W_FOR(a, d.Ap) S.legal[a] = 0; w_sync(); }  // go.py:106-110,140-142; Gomoku keeps its mask
  env_store(E, g, S, true);
  W_LANE0 {
    int32_t* ei = E.env_i + (size_t)g * ENV_INTS;
    ei[EI_LAST_MOVE] = action;
    ei[EI_LAST_PLAYER] = mover;
    ei[EI_DONE] = o.done;
    ei[EI_WINNER] = o.winner;
    ei[EI_BY_RESIGN] = action < 0 ? 1 : 0;
    ei[EI_REWARD_X2] = o.reward_x2;
    memcpy(&ei[EI_SCORE_BITS], &o.score, 4);
    if (action == d.pass_move) ei[EI_NUM_PASSES] += 1;
  }
  w_sync();
  return o;
}
