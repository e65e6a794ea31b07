"""create_mcts_player / play_and_record_one_game / run_selfplay_actor_loop with the reference's signatures
(alpha_zero/core/pipeline.py:83,289,166), so training_go.py / training_gomoku.py can import them unchanged.

Two ways of playing:
  * one env at a time (`create_mcts_player` + `play_and_record_one_game`): the reference's own game loop,
    restated, over the GPU search of mcts.py — used by the evaluator / eval scripts and by the parity tests;
  * `run_selfplay_actor_loop`: ONE actor process per GPU drives thousands of game slots through the
    device-resident loop (Engine.selfplay_tick) and emits one `(game_seq, stats)` queue item per finished
    game, honouring ckpt_event / stop_event / var_ckpt / var_resign_threshold like the reference actor.
"""
import os
import queue
import random
import time
from collections import OrderedDict
from copy import copy

import numpy as np

from .engine import Engine
from .mcts import EngineEvaluator, Node, parallel_uct_search, uct_search  # noqa: F401  (re-exported like the reference)
from .replay import Transition
from .util import CsvWriter, Timer, create_logger, get_time_stamp, make_sgf


def set_seed(seed):
    import torch

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _net_geometry(network):
    sd = network.state_dict()
    filters = sd['conv_block.0.weight'].shape[0]
    blocks = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('res_blocks.'))
    fc = sd['value_head.4.weight'].shape[0]
    return blocks, filters, fc


def _device_index(device):
    idx = getattr(device, 'index', None)
    if idx is None and isinstance(device, str) and ':' in device:
        idx = int(device.split(':')[1])
    return int(idx or 0)


def _env_kind(env):
    """'go' | 'gomoku' for our façades and for the reference's own env objects (GoEnv has a pass move, GomokuEnv has not)."""
    kind = getattr(env, 'game', None)
    if kind in ('go', 'gomoku'):
        return kind
    return 'go' if getattr(env, 'has_pass_move', False) else 'gomoku'


class _NetOnEngine:
    """The nn.Module's parameters copied into an engine-owned network; refreshed when the module changes
    (the actor loop hot-swaps checkpoints with load_state_dict, pipeline.py:232-239)."""

    def __init__(self, network, device, game, board_size, precision=None):
        self.network = network
        blocks, filters, fc = _net_geometry(network)
        precision = precision or os.environ.get('AZ_NET_PRECISION', 'fp32')
        self.engine = Engine(game, board_size, num_games=int(os.environ.get('AZ_EVAL_SLOTS', '64')), max_simulations=8, max_parallel=2,
                             net=(blocks, filters, fc), precision=precision, device=_device_index(device))
        self.version = None
        self.refresh()

    def _version(self):
        return sum(int(p._version) for p in self.network.state_dict().values())

    def refresh(self):
        v = self._version()
        if v != self.version:
            self.engine.set_weights(self.network.state_dict())
            self.version = v

    def net_forward(self, obs):
        self.refresh()
        return self.engine.net_forward(obs)


def create_mcts_player(network, device, num_simulations, num_parallel, root_noise=False, deterministic=False):
    """pipeline.py:83 — returns act(env, root_node, c_puct_base, c_puct_init, warm_up=False) ->
    (move, search_pi, root_Q, best_child_Q, next_root_node).  The network forward runs in the engine's CUDA tower."""
    holder = {}

    def eval_position(state, batched=False):
        ev = holder.get('ev')
        if ev is None:
            raise RuntimeError('evaluator not bound to a game yet')
        return ev(state, batched)

    def act(env, root_node, c_puct_base, c_puct_init, warm_up=False):
        if 'ev' not in holder:
            holder['ev'] = EngineEvaluator(_NetOnEngine(network, device, _env_kind(env), env.board_size))
        if num_parallel > 1:
            return parallel_uct_search(env=env, eval_func=eval_position, root_node=root_node, c_puct_base=c_puct_base, c_puct_init=c_puct_init,
                                       num_simulations=num_simulations, num_parallel=num_parallel, root_noise=root_noise, warm_up=warm_up,
                                       deterministic=deterministic)
        return uct_search(env=env, eval_func=eval_position, root_node=root_node, c_puct_base=c_puct_base, c_puct_init=c_puct_init,
                          num_simulations=num_simulations, root_noise=root_noise, warm_up=warm_up, deterministic=deterministic)

    return act


def play_and_record_one_game(env, mcts_player, resign_disabled, c_puct_base, c_puct_init, warm_up_steps, check_resign_after_steps,
                             resign_threshold, logger):
    """pipeline.py:289-382 — one game through the `mcts_player`, returns (list[Transition], stats)."""
    obs = env.reset()
    states, policies, movers = [], [], []
    root, done, reward = None, False, 0.0
    marked_resign_player, num_passes = None, 0
    while not done:
        move, search_pi, root_q, best_child_q, root = mcts_player(env=env, root_node=root, c_puct_base=c_puct_base, c_puct_init=c_puct_init,
                                                                 warm_up=False if env.steps > warm_up_steps else True)
        states.append(obs)
        policies.append(search_pi)
        movers.append(env.to_play)
        if env.has_resign_move and env.steps > check_resign_after_steps and root_q < resign_threshold and best_child_q < resign_threshold:
            if marked_resign_player is None:
                marked_resign_player = copy(env.to_play)
            if logger is not None:
                logger.debug(f'Search root value: {root_q}, best child value: {best_child_q}')
            if not resign_disabled:
                move = env.resign_move
        obs, reward, done, _ = env.step(move)
        if env.has_pass_move and move == env.pass_move:
            num_passes += 1
    values = [0.0] * len(states)
    if reward != 0.0:
        last = env.last_player
        values = [reward if who == last else -reward for who in movers]
    game_seq = [Transition(state=s, pi_prob=p, value=v) for s, p, v in zip(states, policies, values)]
    stats = {'game_length': len(game_seq), 'game_result': env.get_result_string()}
    if env.has_pass_move:
        stats['num_passes'] = num_passes
    if env.has_resign_move:
        marked = resign_disabled and marked_resign_player is not None
        stats['is_resign_disabled'] = resign_disabled
        stats['is_marked_for_resign'] = marked
        stats['is_could_won'] = bool(marked and env.winner == marked_resign_player)
        stats['marked_resign_player'] = env.get_player_name_by_id(marked_resign_player)
        stats['resign_threshold'] = resign_threshold
    return game_seq, stats


def _result_string(rec, game):
    if rec['by_resign']:
        return 'B+R' if rec['winner'] == 1 else 'W+R'
    if game == 'gomoku':
        return 'B+1.0' if rec['winner'] == 1 else ('W+1.0' if rec['winner'] == 2 else 'DRAW')
    s = rec['score']
    return ('B+%.1f' % s) if s > 0 else (('W+%.1f' % abs(s)) if s < 0 else 'DRAW')


def games_to_queue_items(engine, env, games, states, pis, zs, moves, resign_threshold):
    """Finished-game records of the device loop -> the reference's (game_seq, stats) items (pipeline.py:356-382)."""
    items = []
    pi_dtype = np.float64 if env.has_pass_move else np.float32
    for g in games:
        s0, ln = g['first_sample'], g['game_length']
        seq = [Transition(state=states[s0 + i].copy(), pi_prob=pis[s0 + i].astype(pi_dtype), value=float(zs[s0 + i])) for i in range(ln)]
        stats = {'game_length': ln, 'game_result': _result_string(g, _env_kind(env))}
        if env.has_pass_move:
            stats['num_passes'] = g['num_passes']
        if env.has_resign_move:
            stats['is_resign_disabled'] = bool(g['is_resign_disabled'])
            stats['is_marked_for_resign'] = bool(g['is_marked_for_resign'])
            stats['is_could_won'] = bool(g['is_could_won'])
            stats['marked_resign_player'] = env.get_player_name_by_id(g['marked_resign_player']) if g['marked_resign_player'] else None
            stats['resign_threshold'] = resign_threshold
        mv = [int(m) for m in moves[s0:s0 + ln] if m >= 0]
        colours = ['B' if i % 2 == 0 else 'W' for i in range(len(mv))]
        items.append((seq, stats, list(zip(colours, mv))))
    return items


def run_selfplay_actor_loop(seed, rank, network, device, data_queue, env, num_simulations, num_parallel, c_puct_base, c_puct_init,
                            warm_up_steps, check_resign_after_steps, disable_resign_ratio, save_sgf_dir, save_sgf_interval, logs_dir,
                            load_ckpt, log_level, var_ckpt, var_resign_threshold, ckpt_event, stop_event):
    """pipeline.py:166 — same 22 positional arguments.  One process drives $AZ_ACTOR_GAMES (default 1024) concurrent games
    on `device`; every finished game is logged to actor{rank}.csv and put on `data_queue` exactly like the reference's."""
    import torch

    assert num_simulations > 1
    set_seed(int(seed + rank))
    logger = create_logger(log_level)
    writer = CsvWriter(os.path.join(logs_dir, f'actor{rank}.csv'))
    timer = Timer()
    played_games = training_steps = 0
    last_ckpt = None
    should_save_sgf = bool(save_sgf_dir) and os.path.isdir(save_sgf_dir)

    if load_ckpt is not None and os.path.exists(load_ckpt):
        loaded = torch.load(load_ckpt, map_location='cpu')
        network.load_state_dict(loaded['network'])
        training_steps = loaded['training_steps']
        logger.debug(f'Actor{rank} loaded state from checkpoint "{load_ckpt}"')
    network.eval()

    blocks, filters, fc = _net_geometry(network)
    games = int(os.environ.get('AZ_ACTOR_GAMES', '1024'))
    par = max(1, int(num_parallel))
    kind = _env_kind(env)  # `env` may be our façade or the reference's own GoEnv / GomokuEnv: only its settings are read
    engine = Engine(kind, env.board_size, num_games=games, max_simulations=num_simulations, max_parallel=par,
                    komi=getattr(env, 'komi', 0.0), max_steps=getattr(env, 'max_steps', 0), num_to_win=getattr(env, 'num_to_win', 5),
                    num_stack=env.num_stack, net=(blocks, filters, fc), precision=os.environ.get('AZ_NET_PRECISION', 'bf16'),
                    device=_device_index(device), seed=int(seed + rank))
    engine.set_weights(network.state_dict())
    resign_threshold = var_resign_threshold.value if env.has_resign_move else -1
    engine.selfplay_begin(num_simulations, par, c_puct_base, c_puct_init, warm_up_steps, check_resign_after_steps, resign_threshold,
                          disable_resign_ratio, root_noise=True, deterministic=False)
    ticks_per_round = max(1, (num_simulations + 2 * par - 1) // par)
    t_last = time.time()

    while not stop_event.is_set():
        if ckpt_event.is_set():  # the learner is writing a checkpoint: idle like the reference actor (pipeline.py:229)
            time.sleep(0.01)
            continue
        new_ckpt = var_ckpt.value.decode('utf-8') if isinstance(var_ckpt.value, bytes) else str(var_ckpt.value)
        if new_ckpt != '' and new_ckpt != last_ckpt and os.path.exists(new_ckpt):
            loaded = torch.load(new_ckpt, map_location='cpu')
            network.load_state_dict(loaded['network'])
            training_steps = loaded['training_steps']
            engine.set_weights(network.state_dict())
            last_ckpt = new_ckpt
            # The reference swaps weights only between games and throws away the game that was in flight while the learner
            # trained (pipeline.py:232-239, 264-267), so no game ever mixes two weight sets and `training_steps` names the one
            # it was played with.  Same here: the games in flight are abandoned (nothing is emitted for them) and every slot
            # starts a new game on the new weights; what had already finished under the old weights is dropped as well.
            engine.selfplay_restart(np.arange(games, dtype=np.int32))
            while engine.drain_games()[0]:
                pass
            logger.debug(f'Actor{rank} switched to checkpoint "{new_ckpt}"')
        if env.has_resign_move and var_resign_threshold.value != resign_threshold:
            resign_threshold = var_resign_threshold.value
            engine.selfplay_update(warm_up_steps, check_resign_after_steps, resign_threshold, disable_resign_ratio)

        engine.selfplay_tick(ticks_per_round)
        finished, states, pis, zs = engine.drain_games()
        if stop_event.is_set():
            break
        if ckpt_event.is_set() or not finished:
            # games that finished while the learner was busy are discarded, as in the reference (pipeline.py:266-267: `if
            # ckpt_event.is_set(): continue` after the game): the learner would drop them anyway, its training_steps has moved on
            # (pipeline.py:492).  The loop does not tick while the event is set, so at most one round's finishers are lost.
            continue
        now = time.time()
        per_game = (now - t_last) / len(finished)
        t_last = now
        for seq, stats, history in games_to_queue_items(engine, env, finished, states, pis, zs, engine.last_moves, resign_threshold):
            played_games += 1
            timer.history.append(per_game)
            stats['time_per_game'] = round(timer.mean_time(), 4)
            stats['training_steps'] = training_steps
            writer.write(OrderedDict((k, v) for k, v in {'datetime': get_time_stamp(), **stats}.items()))
            if should_save_sgf and played_games % save_sgf_interval == 0:
                sgf = make_sgf(env.board_size, history, stats['game_result'], ruleset='Chinese' if env.has_pass_move else '',
                               komi=getattr(env, 'komi', ''), date=get_time_stamp())
                with open(os.path.join(save_sgf_dir, f'actor{rank}_{get_time_stamp(True)}_{played_games}.sgf'), 'w') as f:
                    f.write(sgf)
            # a bounded queue (maxsize = num_actors, training_go.py:279) fills up when the learner is busy or has finished: keep
            # looking at the stop signal instead of blocking in put() for ever
            while not stop_event.is_set():
                try:
                    data_queue.put((seq, stats), timeout=0.5)
                    break
                except queue.Full:
                    continue

    logger.debug(f'Actor{rank} received stop signal.')
    writer.close()
    engine.close()
    # the learner has stopped reading by now: do not let this process wait at exit for its queue feeder thread to push a last,
    # never-to-be-read game through the pipe (the driver joins the actors, training_go.py:384-388)
    if hasattr(data_queue, 'cancel_join_thread'):
        data_queue.cancel_join_thread()
