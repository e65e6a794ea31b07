"""Engine: a set of game slots resident on one B200, driven through the C ABI (include/az_engine.h).

This is the object underneath the reference-shaped façades (envs/, mcts.py, pipeline.py) and the API
bench.py measures.  numpy arrays in, numpy arrays out; all compute happens in libaz_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AzConfig, AzCounters, AzGameRecord, AzSearchParams, AzSelfplayParams, as_ptr, i32

# AlphaZeroNet.state_dict() order with num_batches_tracked dropped (core/network.py:85-156)


def state_dict_tensors(state_dict):
    """float32 host arrays of a reference-shaped state_dict, in az_set_weights order."""
    out = []
    for k, v in state_dict.items():
        if k.endswith('num_batches_tracked'):
            continue
        a = v.detach().cpu().numpy() if hasattr(v, 'detach') else np.asarray(v)
        out.append(np.ascontiguousarray(a, dtype=np.float32))
    return out


class _PinnedBlock:
    """cudaHostAlloc'ed bytes exposed through the array interface: numpy arrays made from it keep it alive."""

    def __init__(self, binding, nbytes):
        p = C.c_void_p()
        binding.check(binding.dll.az_host_alloc(C.c_size_t(nbytes), C.byref(p)))
        self._free, self._ptr = binding.dll.az_host_free, p
        self.__array_interface__ = {'data': (p.value, False), 'shape': (nbytes,), 'typestr': '|u1', 'version': 3}

    def __del__(self):
        try:
            self._free(self._ptr)
        except Exception:
            pass


class Engine:
    def __init__(self, game, board_size, num_games=1, max_simulations=800, max_parallel=8, komi=7.5, max_steps=0, num_to_win=5,
                 num_stack=8, net=None, precision='fp32', device=0, seed=1, sample_ring=0, binding=None):
        self.b = binding if binding is not None else _lib.load()
        self.game = {'go': _lib.GAME_GO, 'gomoku': _lib.GAME_GOMOKU}[game] if isinstance(game, str) else int(game)
        nb, nf, fc = net if net is not None else (0, 0, 0)
        cfg = AzConfig(self.game, board_size, num_stack, komi, max_steps, num_to_win, num_games, max_simulations, max_parallel,
                       nb, nf, fc, _lib.PRECISIONS[precision], device, seed, sample_ring, 0)
        h = C.c_void_p()
        self.b.check(self.b.dll.az_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.b.check(self.b.dll.az_get_config(self.h, C.byref(cfg)))
        self.cfg = cfg
        self.N = board_size
        self.G = num_games
        self.A = self.b.dll.az_num_actions(self.h)
        self.obs_bytes = self.b.dll.az_obs_bytes(self.h)
        self.planes = 2 * num_stack + 1
        self.max_parallel = max_parallel
        self._active = None
        self._drain_buf = None

    def close(self):
        if getattr(self, 'h', None):
            self.b.dll.az_destroy(self.h)
            self.h = None
            self._drain_buf = None  # the pinned blocks go when the last array the caller still holds goes

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- network ------------------------------------------------------------------------------------
    def set_weights(self, state_dict):
        ts = state_dict_tensors(state_dict)
        ptrs = (C.POINTER(C.c_float) * len(ts))(*[as_ptr(t, C.c_float) for t in ts])
        numel = (C.c_int64 * len(ts))(*[t.size for t in ts])
        self.b.check(self.b.dll.az_set_weights(self.h, ptrs, numel, len(ts)))
        self.weight_bytes = int(sum(t.nbytes for t in ts))

    def set_weights_for(self, which, state_dict):
        """Weight set 0 (== set_weights) or 1 (the opponent of an evaluation match)."""
        ts = state_dict_tensors(state_dict)
        ptrs = (C.POINTER(C.c_float) * len(ts))(*[as_ptr(t, C.c_float) for t in ts])
        numel = (C.c_int64 * len(ts))(*[t.size for t in ts])
        self.b.check(self.b.dll.az_set_weights_for(self.h, int(which), ptrs, numel, len(ts)))

    def net_forward(self, obs):
        obs = np.ascontiguousarray(obs, dtype=np.int8).reshape(-1, self.obs_bytes)
        n = obs.shape[0]
        pri = np.empty((n, self.A), dtype=np.float32)
        val = np.empty(n, dtype=np.float32)
        self.b.check(self.b.dll.az_net_forward(self.h, as_ptr(obs, C.c_int8), n, as_ptr(pri, C.c_float), as_ptr(val, C.c_float)))
        return pri, val

    def net_conv_layer(self, layer, x, res=None):
        """One conv layer of the tower through the kernel the self-play loop launches for it (az_net_conv_layer):
        x float32 [n, cin, Hc, Hc] (+ res float32 [n, filters, Hc, Hc] for the second conv of a block) -> float32 [n, filters, Hc, Hc]."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, _, hc, _ = x.shape
        r = None if res is None else np.ascontiguousarray(res, dtype=np.float32)
        out = np.empty((n, int(self.cfg.num_filters), hc, hc), dtype=np.float32)
        self.b.check(self.b.dll.az_net_conv_layer(self.h, int(layer), as_ptr(x, C.c_float), as_ptr(r, C.c_float) if r is not None else None, n,
                                                  as_ptr(out, C.c_float)))
        return out

    def net_info(self):
        mode, cp = C.c_int32(), C.c_int32()
        fl = C.c_double()
        self.b.check(self.b.dll.az_net_info(self.h, C.byref(mode), C.byref(cp), C.byref(fl)))
        return dict(tc_mode=int(mode.value), padded_filters=int(cp.value), flops_per_eval=float(fl.value))

    # ---- env ----------------------------------------------------------------------------------------
    def env_reset(self, slots):
        s = i32(slots).ravel()
        self.b.check(self.b.dll.az_env_reset(self.h, as_ptr(s, C.c_int32), s.size))

    def env_step(self, slots, actions):
        s, a = i32(slots).ravel(), i32(actions).ravel()
        r = np.zeros(s.size, dtype=np.float32)
        d = np.zeros(s.size, dtype=np.int32)
        self.b.check(self.b.dll.az_env_step(self.h, as_ptr(s, C.c_int32), as_ptr(a, C.c_int32), s.size, as_ptr(r, C.c_float), as_ptr(d, C.c_int32)))
        return r, d

    def env_observation(self, slot):
        out = np.empty(self.obs_bytes, dtype=np.int8)
        self.b.check(self.b.dll.az_env_observation(self.h, int(slot), as_ptr(out, C.c_int8)))
        return out.reshape(self.planes, self.N, self.N)

    def env_legal(self, slot):
        out = np.empty(self.A, dtype=np.uint8)
        self.b.check(self.b.dll.az_env_legal_actions(self.h, int(slot), as_ptr(out, C.c_uint8)))
        return out

    def env_board(self, slot):
        out = np.empty(self.N * self.N, dtype=np.int8)
        self.b.check(self.b.dll.az_env_board(self.h, int(slot), as_ptr(out, C.c_int8)))
        return out.reshape(self.N, self.N)

    def env_scalars(self, slot):
        out = np.zeros(11, dtype=np.int32)
        self.b.check(self.b.dll.az_env_scalars(self.h, int(slot), as_ptr(out, C.c_int32)))
        keys = ('to_play', 'steps', 'last_move', 'last_player', 'winner', 'done', 'ko', 'by_resign', 'caps_b', 'caps_w', 'num_passes')
        return dict(zip(keys, (int(v) for v in out)))

    def env_score(self, slot):
        out = C.c_float()
        self.b.check(self.b.dll.az_env_score(self.h, int(slot), C.byref(out)))
        return float(out.value)

    def env_copy(self, src, dst):
        self.b.check(self.b.dll.az_env_copy(self.h, int(src), int(dst)))

    def env_export(self, slot):
        n = self.b.dll.az_env_state_bytes(self.h)
        out = np.empty(n, dtype=np.uint8)
        self.b.check(self.b.dll.az_env_export(self.h, int(slot), as_ptr(out, C.c_uint8)))
        return out.tobytes()

    def env_import(self, slot, blob):
        buf = np.frombuffer(blob, dtype=np.uint8).copy()
        self.b.check(self.b.dll.az_env_import(self.h, int(slot), as_ptr(buf, C.c_uint8)))

    def env_replay(self, slots, games, want_states=True):
        """Replay `games` (lists of flat actions) on `slots`, one warp per game in one kernel (core/eval_dataset.py:166-215).
        Returns (states [total, planes, N, N] int8 or None, offsets, n_played, status); states[offsets[i] + t] is the observation
        before move t of game i; a game stops at its first rejected move (status < 0)."""
        s = i32(slots).ravel()
        n = len(s)
        if len(games) != n:
            raise ValueError('one move list per slot')
        offsets = np.zeros(n + 1, dtype=np.int32)
        offsets[1:] = np.cumsum([len(g) for g in games])
        flat = np.asarray([a for g in games for a in g], dtype=np.int64)
        if flat.size and (flat.min() < -32768 or flat.max() > 32767):
            raise ValueError('action out of int16 range')
        moves = np.ascontiguousarray(flat, dtype=np.int16) if flat.size else np.zeros(1, dtype=np.int16)
        states = np.zeros((int(offsets[-1]), self.planes, self.N, self.N), dtype=np.int8) if want_states else None
        played = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self.b.check(self.b.dll.az_env_replay(self.h, as_ptr(s, C.c_int32), n, as_ptr(moves, C.c_int16), as_ptr(offsets, C.c_int32),
                                              as_ptr(states, C.c_int8) if want_states and states.size else None,
                                              as_ptr(played, C.c_int32), as_ptr(status, C.c_int32)))
        return states, offsets, played, status

    # ---- search (split phase) ---------------------------------------------------------------------
    def search_begin(self, slots, reuse, c_puct_base, c_puct_init, num_simulations, num_parallel, root_noise=False,
                     warm_up=False, deterministic=False, noise=None):
        s = i32(slots).ravel()
        r = i32(reuse).ravel()
        p = AzSearchParams(c_puct_base, c_puct_init, int(num_simulations), int(num_parallel), int(bool(root_noise)), int(bool(deterministic)))
        nz = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64).reshape(s.size, self.A)
        self.b.check(self.b.dll.az_search_begin(self.h, as_ptr(s, C.c_int32), as_ptr(r, C.c_int32), s.size, C.byref(p), int(bool(warm_up)),
                                                as_ptr(nz, C.c_double) if nz is not None else None))
        self._active = s.copy()
        self._leaf_buf = np.empty((s.size * max(1, int(num_parallel)), self.obs_bytes), dtype=np.int8)

    def search_select(self):
        counts = np.zeros(self._active.size, dtype=np.int32)
        n, act = C.c_int32(), C.c_int32()
        self.b.check(self.b.dll.az_search_select(self.h, as_ptr(self._leaf_buf, C.c_int8), as_ptr(counts, C.c_int32), C.byref(n), C.byref(act)))
        obs = self._leaf_buf[: n.value].reshape(n.value, self.planes, self.N, self.N)
        return obs, counts, int(act.value)

    def search_apply(self, priors, values):
        n = 0 if priors is None else len(priors)
        if n:
            pri = np.ascontiguousarray(priors, dtype=np.float32).reshape(n, self.A)
            val = np.ascontiguousarray(values, dtype=np.float32).reshape(n)
            self.b.check(self.b.dll.az_search_apply(self.h, as_ptr(pri, C.c_float), as_ptr(val, C.c_float), n))
        else:
            self.b.check(self.b.dll.az_search_apply(self.h, None, None, 0))

    def search_run(self):
        self.b.check(self.b.dll.az_search_run(self.h))

    def search_result(self, slot):
        cn = np.empty(self.A, dtype=np.float32)
        cw = np.empty(self.A, dtype=np.float32)
        pi = np.empty(self.A, dtype=np.float64)
        rq = C.c_double()
        mv = C.c_int32()
        self.b.check(self.b.dll.az_search_result(self.h, int(slot), as_ptr(cn, C.c_float), as_ptr(cw, C.c_float), as_ptr(pi, C.c_double), C.byref(rq), C.byref(mv)))
        return dict(child_N=cn, child_W=cw, pi=pi, root_q=float(rq.value), argmax=int(mv.value))

    def search_commit(self, slot, move):
        bq, nx = C.c_double(), C.c_int32()
        self.b.check(self.b.dll.az_search_commit(self.h, int(slot), int(move), C.byref(bq), C.byref(nx)))
        return float(bq.value), bool(nx.value)

    # ---- device-resident self-play -----------------------------------------------------------------
    def selfplay_begin(self, num_simulations, num_parallel, c_puct_base=19652.0, c_puct_init=1.25, warm_up_steps=16,
                       check_resign_after_steps=40, resign_threshold=-1.0, disable_resign_ratio=0.1, root_noise=True, deterministic=False):
        sp = AzSelfplayParams(AzSearchParams(c_puct_base, c_puct_init, int(num_simulations), int(num_parallel), int(root_noise), int(deterministic)),
                              int(warm_up_steps), int(check_resign_after_steps), float(resign_threshold), float(disable_resign_ratio))
        self.b.check(self.b.dll.az_selfplay_begin(self.h, C.byref(sp)))

    def selfplay_update(self, warm_up_steps, check_resign_after_steps, resign_threshold, disable_resign_ratio, search=None):
        """Policy knobs of the running loop; `search` = dict(num_simulations, num_parallel[, c_puct_base, c_puct_init, root_noise,
        deterministic]) also replaces the search parameters from the next leaf batch on (games keep running)."""
        sr = AzSearchParams(0.0, 0.0, 1, 1, 0, 0)  # c_puct_base <= 0: keep the current search parameters
        if search is not None:
            sr = AzSearchParams(float(search.get('c_puct_base', 19652.0)), float(search.get('c_puct_init', 1.25)), int(search['num_simulations']),
                                int(search['num_parallel']), int(bool(search.get('root_noise', True))), int(bool(search.get('deterministic', False))))
        sp = AzSelfplayParams(sr, int(warm_up_steps), int(check_resign_after_steps), float(resign_threshold), float(disable_resign_ratio))
        self.b.check(self.b.dll.az_selfplay_update(self.h, C.byref(sp)))

    def selfplay_restart(self, slots):
        """Abandon the games running in `slots` (nothing emitted) and start new ones there; call between ticks."""
        s = i32(slots).ravel()
        if s.size:
            self.b.check(self.b.dll.az_selfplay_restart(self.h, as_ptr(s, C.c_int32), s.size))

    def match_begin(self, num_simulations, num_parallel, c_puct_base=19652.0, c_puct_init=1.25, black_net=None, games_per_slot=1, alternate=False,
                    deterministic=False):
        """Arm the device-resident match loop: weight set 0 vs weight set 1 in every slot (az_match_begin)."""
        p = AzSearchParams(c_puct_base, c_puct_init, int(num_simulations), int(num_parallel), 0, int(bool(deterministic)))
        bn = None if black_net is None else np.ascontiguousarray(black_net, dtype=np.uint8).reshape(self.G)
        self.b.check(self.b.dll.az_match_begin(self.h, C.byref(p), as_ptr(bn, C.c_uint8) if bn is not None else None, int(games_per_slot), int(bool(alternate))))

    def match_tick(self, n=1):
        run = C.c_int32()
        self.b.check(self.b.dll.az_match_tick(self.h, int(n), C.byref(run)))
        return int(run.value)

    def selfplay_tick(self, n=1):
        self.b.check(self.b.dll.az_selfplay_tick(self.h, int(n)))

    def sync(self):
        self.b.check(self.b.dll.az_sync(self.h))

    def counters(self):
        c = AzCounters()
        self.b.check(self.b.dll.az_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in AzCounters._fields_}

    def _pinned(self, nbytes, dtype):
        """numpy view of a page-locked host block; the block is released when the last view of it is garbage-collected, so arrays
        handed to the caller stay valid after close()."""
        return np.asarray(_PinnedBlock(self.b, nbytes)).view(dtype)

    def _drain_buffers(self, max_samples):
        if self._drain_buf is None or self._drain_buf[0] < max_samples:
            st = self._pinned(max_samples * self.obs_bytes, np.int8).reshape(max_samples, self.obs_bytes)
            pis = self._pinned(max_samples * self.A * 4, np.float32).reshape(max_samples, self.A)
            z = self._pinned(max_samples * 4, np.float32)
            mv = self._pinned(max_samples * 2, np.int16)
            self._drain_buf = (max_samples, st, pis, z, mv)
        return self._drain_buf[1:]

    def drain_games(self, max_games=4096, max_samples=None, copy=False):
        """Finished games, oldest first: (records, states int8 [n, planes, N, N], pis f32 [n, A], z f32 [n]).  The arrays are views
        of this engine's pinned host buffers (the device copies land there by DMA) and are overwritten by the next call;
        copy=True returns private copies.  Samples beyond `max_samples` stay queued on the device for the next call."""
        max_samples = max_samples or min(int(self.cfg.sample_ring), 65536)
        recs = (AzGameRecord * max_games)()
        st, pis, z, mv = self._drain_buffers(max_samples)
        ng, ns = C.c_int32(), C.c_int32()
        self.b.check(self.b.dll.az_drain_games(self.h, recs, max_games, C.byref(ng), as_ptr(st, C.c_int8), as_ptr(pis, C.c_float),
                                               as_ptr(z, C.c_float), as_ptr(mv, C.c_int16), max_samples, C.byref(ns)))
        n = ns.value
        self.last_moves = mv[:n].copy()
        games = [{k: getattr(recs[i], k) for k, _ in AzGameRecord._fields_} for i in range(ng.value)]
        out = (st[:n].reshape(n, self.planes, self.N, self.N), pis[:n], z[:n])
        if copy:
            out = tuple(a.copy() for a in out)
        return (games,) + out

    def gather_pack(self, d_states, d_pis, d_values, max_samples, max_games=4096):
        """Pack the samples of the finished games into caller-owned DEVICE buffers (raw pointers, e.g. tensor.data_ptr()):
        the send block of the sample all-gather.  Returns (records, n_samples); nothing but the counts crosses PCIe."""
        recs = (AzGameRecord * max_games)()
        ng, ns = C.c_int32(), C.c_int32()
        self.b.check(self.b.dll.az_gather_pack(self.h, recs, max_games, C.byref(ng), C.c_void_p(d_states), C.c_void_p(d_pis), C.c_void_p(d_values),
                                               int(max_samples), C.byref(ns)))
        games = [{k: getattr(recs[i], k) for k, _ in AzGameRecord._fields_} for i in range(ng.value)]
        return games, int(ns.value)

    # ---- device-resident replay (learner input path) -----------------------------------------------
    def replay_create(self, capacity):
        self.b.check(self.b.dll.az_replay_create(self.h, int(capacity)))

    def replay_ingest(self):
        ng, ns = C.c_int32(), C.c_int32()
        self.b.check(self.b.dll.az_replay_ingest(self.h, C.byref(ng), C.byref(ns)))
        return int(ng.value), int(ns.value)

    def replay_add(self, states, pis, values, n_games=1):
        st = np.ascontiguousarray(states, dtype=np.int8).reshape(-1, self.obs_bytes)
        pi = np.ascontiguousarray(pis, dtype=np.float32).reshape(st.shape[0], self.A)
        z = np.ascontiguousarray(values, dtype=np.float32).reshape(st.shape[0])
        self.b.check(self.b.dll.az_replay_add(self.h, as_ptr(st, C.c_int8), as_ptr(pi, C.c_float), as_ptr(z, C.c_float), st.shape[0], int(n_games)))

    def replay_info(self):
        a, g = C.c_int64(), C.c_int64()
        sz, cap = C.c_int32(), C.c_int32()
        self.b.check(self.b.dll.az_replay_info(self.h, C.byref(a), C.byref(g), C.byref(sz), C.byref(cap)))
        return dict(num_samples_added=int(a.value), num_games_added=int(g.value), size=int(sz.value), capacity=int(cap.value))

    def replay_sample(self, indices, transform=0, out=None):
        """indices int32 [B] -> (states int8 [B,planes,N,N], pis f32 [B,A], values f32 [B]) on the host; with out=(states_ptr,
        pis_ptr, values_ptr) device pointers the batch is written into the caller's CUDA tensors and True is returned."""
        idx = i32(indices).ravel()
        B = idx.size
        if out is not None:
            self.b.check(self.b.dll.az_replay_sample(self.h, as_ptr(idx, C.c_int32), B, int(transform), C.c_void_p(out[0]), C.c_void_p(out[1]),
                                                     C.c_void_p(out[2]), 1))
            return True
        st = np.empty((B, self.obs_bytes), dtype=np.int8)
        pi = np.empty((B, self.A), dtype=np.float32)
        z = np.empty(B, dtype=np.float32)
        self.b.check(self.b.dll.az_replay_sample(self.h, as_ptr(idx, C.c_int32), B, int(transform), as_ptr(st, C.c_int8), as_ptr(pi, C.c_float),
                                                 as_ptr(z, C.c_float), 0))
        return st.reshape(B, self.planes, self.N, self.N), pi, z

    def stream(self):
        p = C.c_void_p()
        self.b.check(self.b.dll.az_stream(self.h, C.byref(p)))
        return p.value

    def tick_profile(self, enable):
        """Per-phase device time (ms) accumulated since the last call, then switch the per-tick event recording on / off."""
        ms = (C.c_double * 5)()
        n = C.c_int32()
        self.b.check(self.b.dll.az_tick_profile(self.h, int(bool(enable)), ms, C.byref(n)))
        keys = ('select', 'network', 'expand_backup', 'move_reroot', 'tick')
        return {k: float(v) for k, v in zip(keys, ms)}, int(n.value)

    def last_net_ms(self):
        ms, n = C.c_float(), C.c_int32()
        self.b.check(self.b.dll.az_last_net_ms(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)
