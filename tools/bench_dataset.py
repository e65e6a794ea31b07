"""Throughput of the evaluation-dataset on-ramp: az_env_replay (one warp per recorded game, observation
written before every move, host buffers in and out) vs the CPU oracle replaying the same games one by one, plus the batched
evaluation matches of alpha_zero_b200/matches.py on real towers.  Input: the reference's 2412 self-play records
(tests/golden/go9_selfplay.npz), tiled to 9648 games."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from alpha_zero_b200.engine import Engine

g = np.load('tests/golden/go9_selfplay.npz')
off = g['offsets']
games = [g['moves'][off[i]:off[i] + g['played'][i]].tolist() for i in range(len(off) - 1)] * 4
eng = Engine('go', 9, num_games=len(games), max_simulations=1, max_parallel=1, komi=7.5, net=None)
slots = list(range(len(games)))
for _ in range(2):
    eng.env_replay(slots, games)
t0 = time.perf_counter()
iters = 5
for _ in range(iters):
    states, offsets, played, status = eng.env_replay(slots, games)
dt = (time.perf_counter() - t0) / iters
pos = int(offsets[-1])
assert (status == 0).all() and int(played.sum()) == pos
t0 = time.perf_counter()
for _ in range(iters):
    eng.env_replay(slots, games, want_states=False)
dt_dev = (time.perf_counter() - t0) / iters
eng.close()

from oracle.boards import GoBoard  # CPU checker, timed as the reference-style baseline (one game at a time)

env = GoBoard(9, 7.5, 8, 162)
t0 = time.perf_counter()
cpu_pos = 0
for mv in games[:300]:
    env.reset()
    for a in mv:
        env.observation()
        env.step(a)
        cpu_pos += 1
dt_cpu = time.perf_counter() - t0
print(json.dumps({'metric': 'replayed_positions_per_sec', 'games': len(games), 'positions': pos, 'value': pos / dt, 'ms': dt * 1e3,
                  'without_observations': pos / dt_dev, 'd2h_bytes': int(states.nbytes), 'cpu_port_1core': cpu_pos / dt_cpu,
                  'note': 'end to end through az_env_replay: H2D moves, one kernel, D2H observations'}))

# batched evaluation matches, two random 10x128 towers, 64 games, 100 simulations
from alpha_zero_b200.matches import play_matches
from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm

torch.manual_seed(1)
os.environ['AZ_NET_PRECISION'] = 'bf16'
black = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 82, 10, 128, 128, False)).eval()
white = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 82, 10, 128, 128, False)).eval()
np.random.seed(0)
t0 = time.perf_counter()
res = play_matches(64, 'go', 9, black, white, torch.device('cuda:0'), num_simulations=100, num_parallel=8, max_steps=60)
dt = time.perf_counter() - t0
moves = sum(r['game_length'] for r in res)
print(json.dumps({'metric': 'evaluation_match_simulations_per_sec', 'games': len(res), 'moves': moves, 'value': moves * 100 / dt, 'seconds': dt,
                  'black_wins': sum(r['winner'] == 'B' for r in res), 'note': 'lock-step games, leaves of all games in one batch per network call, host hop'}))

# the same matches with both weight sets in ONE engine and the loop on the device (az_match_begin / az_match_tick): the trained
# checkpoint against a random-init net when the checkpoint fixture is present, 400 simulations, games capped at 40 plies
from alpha_zero_b200.matches import play_matches_on_device

ck = os.path.join('tests', 'golden', 'ckpt_go9_154000.npz')
if os.path.exists(ck):
    w = np.load(ck)
    black.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files if k != 'versions'})
for n_games, slots, sims in ((1024, 1024, 400), (4096, 4096, 400)):
    t0 = time.perf_counter()
    res = play_matches_on_device(n_games, 'go', 9, black, white, torch.device('cuda:0'), num_simulations=sims, num_parallel=8, swap_colours=True, slots=slots,
                                 precision='bf16', max_steps=40)
    dt = time.perf_counter() - t0
    c = res[0]['engine_counters']
    first_wins = sum((r['winner'] == 'B') == (r['black_is'] == 'black') and r['winner'] is not None for r in res)
    print(json.dumps({'metric': 'evaluation_match_simulations_per_sec', 'impl': 'device loop, two weight sets in one engine', 'games': len(res), 'slots': slots,
                      'moves': c['moves'], 'value': c['simulations'] / dt, 'seconds': dt, 'games_per_sec': len(res) / dt,
                      'mean_game_length': sum(r['game_length'] for r in res) / len(res), 'first_network_wins': first_wins,
                      'note': 'wall clock incl. engine creation and weight upload; games capped at 40 plies; fresh tree every move, no noise, T=0.1 sampling on the device'}))
