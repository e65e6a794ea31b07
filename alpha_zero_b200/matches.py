"""Batched evaluation matches: what eval_against_prev_ckpt (core/pipeline.py:815-867) and
eval_play/eval_agent_go_mass_matches.py:103-146 do one game at a time — black network vs white network, a fresh search tree
every move (`root_node=None`), no root noise — for many games at once on one engine.

All games advance in lock step (same side to move), so each ply needs ONE network: the leaves of every running game are
collected by the tree kernels (az_search_select), evaluated in one batch by the mover's network on its CUDA tower
(az_net_forward of that network's engine) and applied (az_search_apply).  Move choice follows mcts_v2.py:630-641 with the
same numpy calls (`argmax(child_N)` when deterministic, otherwise `np.random.choice` on the T=0.1 / warm-up policy).
"""
import os

import numpy as np

from .engine import Engine
from .mcts import generate_search_policy
from .pipeline import _NetOnEngine, _device_index, _net_geometry, _result_string


def play_matches_on_device(num_games, game, board_size, black, white, device=None, num_simulations=400, num_parallel=8, c_puct_base=19652.0,
                           c_puct_init=1.25, komi=7.5, max_steps=0, num_to_win=5, num_stack=8, deterministic=False, swap_colours=False,
                           slots=None, precision=None, seed=1, binding=None):
    """The same matches with BOTH weight sets resident in one engine and the whole loop on the device (az_match_begin /
    az_match_tick): every leaf batch the tree kernels collect the leaves of all running games, each network evaluates the
    leaves of the games in which its colour is to move, moves are chosen on the device (argmax(child_N) when `deterministic`,
    else sampled from the T = 0.1 policy with the engine's counter-based RNG instead of numpy's), finished games come back
    as records.  `black` / `white` are nn.Modules of the same geometry; with `swap_colours` odd games have them trade colours
    (eval_agent_go_mass_matches.py alternates sides the same way).  Returns one dict per game like play_matches, plus `black_is`
    ('black' / 'white': which of the two arguments held the black stones)."""
    has_pass = game == 'go'
    geo = _net_geometry(black)
    if _net_geometry(white) != geo:
        raise ValueError('both networks must have the same geometry')
    G = min(num_games, int(slots or os.environ.get('AZ_EVAL_SLOTS_DEVICE', '1024')))
    per_slot = (num_games + G - 1) // G
    eng = Engine(game, board_size, num_games=G, max_simulations=num_simulations, max_parallel=max(1, num_parallel), komi=komi,
                 max_steps=max_steps, num_to_win=num_to_win, num_stack=num_stack, net=geo, device=_device_index(device), seed=seed,
                 precision=precision or os.environ.get('AZ_NET_PRECISION', 'bf16'), binding=binding)
    eng.set_weights_for(0, black.state_dict())
    eng.set_weights_for(1, white.state_dict())
    # slot g plays games g, g + G, g + 2G, ...; with swap_colours game k has `white` on the black stones when k is odd
    black_net = np.array([(g & 1) if swap_colours else 0 for g in range(G)], dtype=np.uint8)
    alternate = bool(swap_colours and (G & 1))  # keeps "odd game index <-> swapped" when a slot's consecutive games differ in parity
    eng.match_begin(num_simulations, num_parallel, c_puct_base, c_puct_init, black_net=black_net, games_per_slot=per_slot, alternate=alternate,
                    deterministic=deterministic)
    ticks = max(1, (num_simulations + 2 * max(1, num_parallel) - 1) // max(1, num_parallel))
    black_id, white_id = (1, -1) if has_pass else (1, 2)
    out = {}
    running = G
    while True:
        running = eng.match_tick(ticks)
        recs, st, pis, zs = eng.drain_games()
        mv = eng.last_moves
        for r in recs:
            k = (r['reserved'] // G) * G + r['slot']  # game index: (games started earlier in the slot) * G + slot
            if k >= num_games:
                continue
            s0, ln = r['first_sample'], r['game_length']
            moves = [int(m) for m in mv[s0:s0 + ln]]
            nb = int(black_net[r['slot']]) ^ (((r['reserved'] // G) & 1) if alternate else 0)
            stats = {'game': k, 'game_result': _result_string(r, game), 'game_length': ln, 'moves': moves,
                     'winner': {black_id: 'B', white_id: 'W'}.get(r['winner']), 'black_is': 'white' if nb else 'black'}
            if has_pass:
                stats['num_passes'] = r['num_passes']
            out[k] = stats
        if running == 0 and not recs:
            break
    c = eng.counters()
    eng.close()
    if c['errors'] or c['ring_dropped']:
        raise RuntimeError(f'match loop: device errors {c["errors"]}, dropped samples {c["ring_dropped"]}')
    res = [out[k] for k in sorted(out)]
    res_stats = dict(simulations=c['simulations'], evaluations=c['evaluations'], moves=c['moves'])
    for r in res:
        r['engine_counters'] = res_stats
    return res


def play_matches(num_games, game, board_size, black, white, device=None, num_simulations=400, num_parallel=8, c_puct_base=19652.0,
                 c_puct_init=1.25, komi=7.5, max_steps=0, num_to_win=5, num_stack=8, deterministic=False, warm_up_steps=-1, binding=None):
    """`black` / `white`: nn.Modules with the AlphaZeroNet state_dict layout, or eval_func callables (obs, batched) -> (priors, values)
    (pipeline.py:91-123).  Returns one dict per game: game_result, game_length, winner, num_passes, moves."""
    has_pass = game == 'go'
    A = board_size * board_size + (1 if has_pass else 0)
    eng = Engine(game, board_size, num_games=num_games, max_simulations=num_simulations, max_parallel=max(1, num_parallel), komi=komi,
                 max_steps=max_steps, num_to_win=num_to_win, num_stack=num_stack, net=None, device=_device_index(device), binding=binding)

    def evaluator(p):
        if callable(p) and not hasattr(p, 'state_dict'):
            return p
        holder = _NetOnEngine(p, device, game, board_size)

        def ev(obs, batched=False):
            st = obs if batched else obs[None, ...]
            pri, val = holder.net_forward(st)
            return ([pri[i] for i in range(len(val))], [float(v) for v in val]) if batched else (pri[0], float(val[0]))

        return ev

    evals = {1: evaluator(black), -1: evaluator(white)}
    black_id, white_id = (1, -1) if has_pass else (1, 2)
    moves = [[] for _ in range(num_games)]
    passes = [0] * num_games
    live = list(range(num_games))
    side, ply = 1, 0
    while live:
        ev = evals[side]
        warm = ply <= warm_up_steps
        eng.search_begin(live, [0] * len(live), c_puct_base, c_puct_init, num_simulations, num_parallel, False, warm, deterministic)
        while True:
            obs, counts, active = eng.search_select()
            if active == 0:
                break
            if len(obs):
                pri, val = ev(obs, True)
                eng.search_apply(np.stack(pri), np.asarray(val, dtype=np.float32))
            else:
                eng.search_apply(None, None)
        chosen = []
        for s in live:
            res = eng.search_result(s)
            legal = eng.env_legal(s).astype(np.int64 if has_pass else np.int8)
            if deterministic:
                mv = int(np.argmax(res['child_N']))
            else:
                pi = generate_search_policy(res['child_N'], 1.0 if warm else 0.1, legal)
                mv = None
                while mv is None or (warm and has_pass and mv == A - 1) or legal[mv] != 1:
                    mv = int(np.random.choice(np.arange(A), p=pi))
            chosen.append(mv)
        _, dones = eng.env_step(live, chosen)
        nxt = []
        for s, mv, d in zip(live, chosen, dones):
            moves[s].append(mv)
            if has_pass and mv == A - 1:
                passes[s] += 1
            if not d:
                nxt.append(s)
        live = nxt
        side, ply = -side, ply + 1
    out = []
    for s in range(num_games):
        sc = eng.env_scalars(s)
        rec = dict(winner=sc['winner'], by_resign=sc['by_resign'], score=eng.env_score(s) if has_pass else 0.0)
        stats = {'game': s, 'game_result': _result_string(rec, game), 'game_length': sc['steps'], 'moves': moves[s],
                 'winner': {black_id: 'B', white_id: 'W'}.get(sc['winner'])}
        if has_pass:
            stats['num_passes'] = passes[s]
        out.append(stats)
    eng.close()
    return out
