/*
 * az_engine.h — C ABI of the B200 self-play engine (libaz_b200.so).
 *
 * The reference (michaelnny/alpha_zero) has no FFI layer: its boundary is a set of Python call
 * signatures.  Every entry point below names the reference interface it sits behind (file:line under
 * /root/reference/alpha_zero); alpha_zero_b200/ binds them with ctypes and re-exposes the reference's
 * Python signatures unchanged (INTEGRATION.md shows the stub).
 *
 * Conventions: every call returns 0 on success, <0 on error (az_last_error() gives the message, mapped
 * by the Python layer onto the reference's ValueError / RuntimeError texts).  The caller owns all host
 * buffers; the engine owns all device buffers.  One engine is bound to one CUDA device and one host
 * thread.  No callbacks into the host language.  No torch types.
 */
#ifndef AZ_ENGINE_H
#define AZ_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZ_GAME_GO 0      /* envs/go.py:19       GoEnv     (stones +1 black / -1 white, pass = N*N, resign = -1) */
#define AZ_GAME_GOMOKU 1  /* envs/gomoku.py:17   GomokuEnv (ids 1 black / 2 white, no pass, no resign)          */

#define AZ_NET_FP32 0 /* CUDA-core fp32 tower: the parity mode (pi within 1e-3 of the reference CPU path) */
#define AZ_NET_BF16 1 /* tcgen05 bf16 tower, f32 accumulate in TMEM: the throughput mode                   */
#define AZ_NET_BF16X3 2 /* tcgen05 tower on split operands (value = bf16 hi + bf16 lo; a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
                           accumulated in f32): the tensor-core parity mode, pi within 1e-3 of the reference CPU path at
                           three times the MMAs of AZ_NET_BF16 */

#define AZ_OK 0
#define AZ_ERR_INVALID_ACTION -2 /* ValueError('Invalid action...')  envs/go.py:92  */
#define AZ_ERR_ILLEGAL_ACTION -3 /* ValueError('Illegal action...')  envs/go.py:94  */
#define AZ_ERR_GAME_OVER -4      /* RuntimeError('Game is over...')  envs/go.py:90, core/mcts_v2.py:360 */
#define AZ_ERR_BAD_ARG -5
#define AZ_ERR_CUDA -6
#define AZ_ERR_STATE -7
#define AZ_ERR_CAPACITY -8

typedef struct az_engine az_engine;

typedef struct az_config {
  int32_t game;            /* AZ_GAME_*                                                        */
  int32_t board_size;      /* N                                                                */
  int32_t num_stack;       /* history depth of the observation, <= 8 (envs/base.py:58-63)       */
  float komi;              /* Go (envs/go.py:47)                                                */
  int32_t max_steps;       /* Go; 0 -> 2*N*N (envs/go.py:49)                                    */
  int32_t num_to_win;      /* Gomoku (envs/gomoku.py:26)                                        */
  int32_t num_games;       /* concurrent game slots held in HBM                                 */
  int32_t max_simulations; /* sizes the per-game node pool: nodes = max_simulations+3*max_parallel+16 */
  int32_t max_parallel;    /* sizes the leaf batch: num_games*max_parallel rows                 */
  int32_t num_res_blocks;  /* core/network.py:85 AlphaZeroNet(..., num_res_block, num_filters, num_fc_units, gomoku) */
  int32_t num_filters;     /*   0 => engine without a network (external evaluator only)         */
  int32_t num_fc_units;
  int32_t net_precision;   /* AZ_NET_*                                                          */
  int32_t device;          /* CUDA ordinal                                                      */
  uint64_t seed;           /* device RNG stream (Dirichlet noise, move sampling, resign lottery) */
  int32_t sample_ring;     /* capacity (samples) of the finished-game sample ring; 0 -> default  */
  int32_t reserved;
} az_config;

/* Parameters of one search call == the arguments of uct_search / parallel_uct_search
 * (core/mcts_v2.py:301-311, 485-496) and of the `act` closure (core/pipeline.py:125-156). */
typedef struct az_search_params {
  double c_puct_base, c_puct_init;
  int32_t num_simulations;
  int32_t num_parallel;  /* <=1: uct_search semantics (no virtual loss, bound = num_simulations);
                             >1: parallel_uct_search (virtual loss, bound = num_simulations+num_parallel) */
  int32_t root_noise;    /* add_dirichlet_noise, core/mcts_v2.py:235-262 */
  int32_t deterministic; /* argmax(child_N) instead of sampling, core/mcts_v2.py:630-641 */
} az_search_params;

/* Parameters of the device-resident self-play loop == the arguments run_selfplay_actor_loop hands to
 * play_and_record_one_game (core/pipeline.py:166-189, 249-259). */
typedef struct az_selfplay_params {
  az_search_params search;
  int32_t warm_up_steps;            /* warm_up = steps <= warm_up_steps, core/pipeline.py:320 */
  int32_t check_resign_after_steps; /* core/pipeline.py:328-341 */
  float resign_threshold;           /* <= -1 disables resignation, core/pipeline.py:216,244 */
  float disable_resign_ratio;       /* per-game lottery, core/pipeline.py:244-246 */
} az_selfplay_params;

typedef struct az_counters {
  uint64_t simulations; /* root-visit increments (SURVEY.md 8d definition) */
  uint64_t evaluations; /* leaves sent to the network / evaluator (duplicates included, terminals not) */
  uint64_t moves;       /* real-game plies played by the self-play loop */
  uint64_t games;       /* finished games */
  uint64_t nodes;       /* tree nodes created */
  uint64_t depth_sum;   /* sum of leaf depths (mean depth = depth_sum / descents) */
  uint64_t descents;    /* select passes (incl. terminal hits) */
  uint64_t samples;     /* (state, pi, z) samples written to the sample ring */
  uint64_t ring_dropped;/* samples overwritten before the host drained them */
  uint64_t errors;      /* sticky device-side error count (node pool exhausted, ...) */
  uint64_t kernel_launches; /* kernels launched by this engine since creation */
  uint64_t ticks;
} az_counters;

/* Finished-game record == the `stats` dict of play_and_record_one_game (core/pipeline.py:367-380). */
typedef struct az_game_record {
  int32_t slot;
  int32_t game_length;
  int32_t winner;         /* black id, white id, or 0 for a draw */
  int32_t by_resign;      /* result string 'B+R' / 'W+R' */
  float score;            /* black - (white + komi) on the final board (Go) */
  int32_t num_passes;
  int32_t is_resign_disabled, is_marked_for_resign, is_could_won, marked_resign_player;
  int32_t first_sample;   /* index (monotonic) of this game's first sample in the sample ring */
  int32_t reserved;
} az_game_record;

const char* az_last_error(void);
int az_version(void);

/* ---- lifetime -------------------------------------------------------------------------------- */
int az_create(const az_config* cfg, az_engine** out);
int az_destroy(az_engine* e);
int az_get_config(az_engine* e, az_config* out);
int az_num_actions(az_engine* e);            /* envs/base.py:65 action_dim */
int az_obs_bytes(az_engine* e);              /* (2*num_stack+1)*N*N int8, envs/base.py:228-259 */

/* ---- network weights: nn.Module.state_dict() -> engine-owned, BN-folded device buffers ----------
 * tensors[i] are host float32 arrays in the order of AlphaZeroNet.state_dict() (core/network.py:85-156)
 * with the num_batches_tracked entries dropped; numel[i] their element counts.  Replaces
 * network.load_state_dict(...) + network.to(device) in run_selfplay_actor_loop (core/pipeline.py:205-214,
 * 232-239). */
int az_set_weights(az_engine* e, const float* const* tensors, const int64_t* numel, int32_t n_tensors);

/* Second weight set of the engine (which = 1; which = 0 is az_set_weights): the opponent of an evaluation match
 * (run_evaluator_loop keeps `network` and `prev_ckpt_network`, core/pipeline.py:722-757). */
int az_set_weights_for(az_engine* e, int32_t which, const float* const* tensors, const int64_t* numel, int32_t n_tensors);

/* Network forward on host observations (eval_position, core/pipeline.py:91-123):
 * obs int8 [n, 2*num_stack+1, N, N] -> priors float32 [n, A] (softmax over ALL actions), values float32 [n]. */
int az_net_forward(az_engine* e, const int8_t* obs, int32_t n, float* priors, float* values);

/* One 3x3 conv layer of the tower (core/network.py:42-82, 101-119; BatchNorm folded, ReLU) on caller-supplied activations,
 * through exactly the kernel az_net_forward / the self-play loop would launch for it: the per-layer parity check of the
 * tensor-core kernels against a plain convolution.  layer 0 = input conv (in: float32 [n, planes, Hc, Hc] on the canvas the
 * tower runs on: Hc = N for Go, N + 4 for Gomoku with the observation at (2, 2), core/network.py:101); odd layers = first
 * conv of a block; even layers >= 2 = second conv of a block, with `res` (float32 [n, num_filters, Hc, Hc], may be NULL)
 * added before the ReLU.  out: float32 [n, num_filters, Hc, Hc].  Values pass through the tower's storage type (bf16
 * rounding; hi + lo for AZ_NET_BF16X3).  Fails with AZ_ERR_STATE if the kernel wrote into the zero padding of its layout. */
int az_net_conv_layer(az_engine* e, int32_t layer, const float* in, const float* res, int32_t n, float* out);
/* tower description: kernel variant (AZ_TC_MODE; -1 for the fp32 tower), channel count the tower runs on (num_filters
 * rounded up to its tile granularity), algorithmic 2*MAC per evaluation of the whole network. */
int az_net_info(az_engine* e, int32_t* tc_mode, int32_t* padded_filters, double* flops_per_eval);

/* ---- BoardGameEnv contract (envs/base.py:26, go.py:88, gomoku.py:45) on game slots -------------- */
int az_env_reset(az_engine* e, const int32_t* slots, int32_t n);                 /* reset() */
/* step(): actions[i] in [0,A) or -1 (resign, Go). rewards/dones out (reward is for the mover). */
int az_env_step(az_engine* e, const int32_t* slots, const int32_t* actions, int32_t n, float* rewards, int32_t* dones);
int az_env_observation(az_engine* e, int32_t slot, int8_t* out);                /* observation() */
int az_env_legal_actions(az_engine* e, int32_t slot, uint8_t* out);             /* legal_actions */
int az_env_board(az_engine* e, int32_t slot, int8_t* out);                      /* board, reference stone ids */
/* scalars: [to_play, steps, last_move, last_player, winner, done, ko, by_resign, caps_b, caps_w, num_passes] */
int az_env_scalars(az_engine* e, int32_t slot, int32_t* out11);
int az_env_score(az_engine* e, int32_t slot, float* out_score);                 /* go_engine.py:509-516 */
int az_env_copy(az_engine* e, int32_t src_slot, int32_t dst_slot);              /* copy.deepcopy(env) */
int az_env_state_bytes(az_engine* e);
int az_env_export(az_engine* e, int32_t slot, uint8_t* out);                    /* pickling */
int az_env_import(az_engine* e, int32_t slot, const uint8_t* in);
/* Replay of recorded games = the loop of replay_sgf (core/eval_dataset.py:166-215) for n games at once: slot i is reset and
 * plays moves[offsets[i] .. offsets[i+1]) (flat actions); states (host, offsets[n]*az_obs_bytes, may be NULL) receives the
 * observation BEFORE each move at row offsets[i]+t.  A game stops at the first move step() would reject: n_played[i] moves
 * were played, status[i] = 0 or AZ_ERR_GAME_OVER / AZ_ERR_INVALID_ACTION / AZ_ERR_ILLEGAL_ACTION.  The slots keep the final
 * positions (az_env_scalars / az_env_score / az_env_board). */
int az_env_replay(az_engine* e, const int32_t* slots, int32_t n, const int16_t* moves, const int32_t* offsets, int8_t* states,
                  int32_t* n_played, int32_t* status);

/* ---- search, split phase: the evaluator lives outside (a Python eval_func, a fake, torch) --------
 * az_search_begin : start uct_search/parallel_uct_search on `slots`; reuse[i]!=0 keeps the slot's
 *                   re-rooted subtree (root_node argument), 0 starts from a fresh root.
 *                   noise: float64 [n, A] host-drawn Dirichlet samples (np.random.dirichlet) or NULL
 *                   (engine draws its own when params.root_noise).
 * az_search_select: one leaf-collection pass (core/mcts_v2.py:568-611).  Writes the leaf observations
 *                   int8 [total, obs_bytes] in slot-major order, counts[i] leaves for slots[i]; returns
 *                   total in *n_leaves.  Slots whose search is complete contribute 0.
 * az_search_apply : priors float32 [total, A], values float32 [total] in the same order: revert virtual
 *                   loss, expand, back up (core/mcts_v2.py:613-625); updates done flags.
 * az_search_result: child_N float32[A], pi float64[A], root_Q, best-child info of a finished search.
 * az_search_commit: re-root on `move` (core/mcts_v2.py:643-653); returns 1 in *has_next if a subtree is kept.
 */
int az_search_begin(az_engine* e, const int32_t* slots, const int32_t* reuse, int32_t n, const az_search_params* p,
                    int32_t warm_up, const double* noise);
int az_search_select(az_engine* e, int8_t* leaf_obs, int32_t* counts, int32_t* n_leaves, int32_t* n_active);
int az_search_apply(az_engine* e, const float* priors, const float* values, int32_t n_leaves);
int az_search_result(az_engine* e, int32_t slot, float* child_N, float* child_W, double* pi, double* root_q,
                     int32_t* argmax_move);
int az_search_commit(az_engine* e, int32_t slot, int32_t move, double* best_child_q, int32_t* has_next);
/* Same loop with the engine's own network as evaluator; runs until every begun search is complete. */
int az_search_run(az_engine* e);

/* ---- device-resident self-play (play_and_record_one_game x num_games, core/pipeline.py:289-382) --
 * az_selfplay_begin resets every slot and arms the loop; az_selfplay_tick(n) runs n leaf batches
 * (select -> network -> apply -> advance) without host synchronisation. */
int az_selfplay_begin(az_engine* e, const az_selfplay_params* p);
int az_selfplay_tick(az_engine* e, int32_t n_ticks);
/* Change warm-up / resignation knobs of the running loop without resetting the games
 * (var_resign_threshold is re-read before every game, core/pipeline.py:241-246).  When p->search.c_puct_base > 0 the
 * search parameters (num_simulations, num_parallel, c_puct_*, root_noise, deterministic) are replaced as well, effective
 * from the next leaf batch: searches in flight simply run to the new bound (the reference reads FLAGS.num_simulations /
 * num_parallel once per actor, core/pipeline.py:166-189; a restarted actor picks up new ones).  c_puct_base <= 0 keeps
 * the current search parameters. */
int az_selfplay_update(az_engine* e, const az_selfplay_params* p);
/* Abandon the games running in `slots` and start new ones there: nothing is emitted for the abandoned games (no samples,
 * no record) — what happens to the game in flight when one of the reference's actor processes is restarted
 * (training_go.py:318-330 spawns them, core/pipeline.py:227 is their game loop).  Must be called between ticks.  bench.py
 * uses it once to stagger the ages of a freshly begun population so that the timed window sees games finishing. */
int az_selfplay_restart(az_engine* e, const int32_t* slots, int32_t n);
/* ---- evaluation matches on the device: eval_against_prev_ckpt (core/pipeline.py:815-867) and
 * eval_play/eval_agent_go_mass_matches.py:103-146 for every slot at once.  Weight set 0 against weight set 1, a fresh search tree
 * every move (root_node=None), no root noise, warm_up=False, no resignation; each player's search is evaluated by its own network.
 * black_net[g] (0 / 1, NULL = all 0) is the weight set that plays black in slot g's first game; with `alternate` the colours swap
 * from one game of a slot to the next; a slot retires after games_per_slot games.  az_match_tick runs n leaf batches for every
 * running game and reports how many slots are still playing; finished games come back through az_drain_games (record.reserved =
 * games_started_before_in_slot * num_games + slot, so the colours of a record follow from black_net / alternate). */
int az_match_begin(az_engine* e, const az_search_params* p, const uint8_t* black_net, int32_t games_per_slot, int32_t alternate);
int az_match_tick(az_engine* e, int32_t n_ticks, int32_t* n_running);
int az_sync(az_engine* e);
int az_get_counters(az_engine* e, az_counters* out);
/* Drain finished games: up to max_games records and their samples, oldest first.
 * states int8 [*, obs_bytes], pis float32 [*, A], values float32 [*] (z, core/pipeline.py:349-354),
 * moves int16 [*] (the move played at each ply, -1 = resign: env.history for to_sgf, envs/go.py:202). */
int az_drain_games(az_engine* e, az_game_record* records, int32_t max_games, int32_t* n_games, int8_t* states,
                   float* pis, float* values, int16_t* moves, int32_t max_samples, int32_t* n_samples);
/* Page-locked host memory for the buffers az_drain_games / az_gather_* fill (the reference hands samples over as pickled numpy
 * arrays, core/pipeline.py:283; here they arrive by DMA, which runs at full PCIe rate only into pinned pages). */
int az_host_alloc(size_t bytes, void** out);
int az_host_free(void* p);
/* Send block of the sample all-gather (the reference's transport is data_queue.put, core/pipeline.py:283, one
 * actor process per game): like az_drain_games, but the samples of the finished games are packed device to device into
 * caller-owned DEVICE buffers (d_states int8 [max_samples, obs_bytes], d_pis float32 [max_samples, A], d_values float32
 * [max_samples]), ready to be handed to ncclAllGather by the caller (torch.distributed on the same device pointers).  Only
 * the counts and the optional game records come back to the host.  Games taken here are not returned by az_drain_games. */
int az_gather_pack(az_engine* e, az_game_record* records, int32_t max_games, int32_t* n_games, void* d_states, void* d_pis,
                   void* d_values, int32_t max_samples, int32_t* n_samples);
/* The CUDA stream every engine kernel is launched on (bench.py times on it with its own events). */
int az_stream(az_engine* e, void** cuda_stream);
/* Mean device time (ms) of the tower's conv launches per network call over the most recent (up to 64) calls (CUDA events on the
 * engine stream around the conv layers only: no input / heads kernels), and the leaves of the last call: bench.py's roofline. */
int az_last_net_ms(az_engine* e, float* ms, int32_t* n_evals);

/* Per-phase device time of the self-play tick, for the measurement tooling (bench.py `tick_breakdown_ms`): with profiling
 * enabled every tick records CUDA events on the engine stream around its phases.  The call returns the times accumulated
 * since the previous call — ms5 = {select (k_collect + k_compact), network, expand/backup (k_apply), move/re-root
 * (k_advance), whole tick} over *n_ticks ticks — resets the accumulators and sets the switch to `enable`.  Off by default:
 * the timed regions of bench.py run without it. */
int az_tick_profile(az_engine* e, int32_t enable, double* ms5, int32_t* n_ticks);

/* ---- device-resident replay: the learner's input path -------------------------------
 * UniformReplay (core/replay.py:35-116) with the storage in HBM: az_replay_ingest == add_game for every finished
 * game, moved from the sample ring device-to-device (games taken here are no longer returned by az_drain_games);
 * az_replay_add == add_game for host samples (other actors, --load_replay); az_replay_sample == get(indices) +
 * the batch-wide dihedral transformation of utils/transformation.py:160 (0 none, 1 h_flip, 2 v_flip, 3/4/5 rotate
 * 90/180/270) fused into the gather.  Indices come from the caller, so `random_state.randint(0, size, batch)`
 * (core/replay.py:75) on the host reproduces the reference's minibatches.  Outputs go to host buffers or, with
 * outputs_on_device != 0, to caller-owned device pointers (e.g. torch CUDA tensors of the learner). */
int az_replay_create(az_engine* e, int32_t capacity);
int az_replay_ingest(az_engine* e, int32_t* n_games, int32_t* n_samples);
int az_replay_add(az_engine* e, const int8_t* states, const float* pis, const float* values, int32_t n, int32_t n_games);
int az_replay_info(az_engine* e, int64_t* num_samples_added, int64_t* num_games_added, int32_t* size, int32_t* capacity);
int az_replay_sample(az_engine* e, const int32_t* indices, int32_t batch, int32_t transform, int8_t* states, float* pis,
                     float* values, int32_t outputs_on_device);

#ifdef __cplusplus
}
#endif
#endif /* AZ_ENGINE_H */
