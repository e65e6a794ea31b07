#!/usr/bin/env python
"""bench.py — MCTS simulations/s (and self-play moves / games per s) of the self-play hot path.

    python bench.py --gpus 1 --steps 5 --warmup 3                  # this framework on N B200s
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the reference CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): 9x9 Go, 4096 concurrent games per GPU, 400 simulations/move,
num_parallel 8, AlphaZeroNet 10 blocks x 128 filters (fc 128), random-init weights (seed 123, BN
statistics randomised), synthetic = self-play games produced by the engine itself.  One *step* = 51
leaf batches (ticks) = (400+8)/8, i.e. about one move of every game: select -> network ->
expand/backup -> move / re-root / recycle, all on the device.  A *simulation* is one root-visit
increment; the count comes from the engine's device counters.

Population: a self-play fleet in steady state holds games of every age, and finished games leave the
device every step.  A cold start (all games at move 0) never sees a game end inside a short window, so
before the warm-up steps the bench plays a cheap prologue (num_simulations = num_parallel, two ticks
per move) and restarts slot g after (g mod L) prologue moves (az_selfplay_restart): the timed steps then
run on games aged 0..L-1 plies, with finished games leaving the device inside the e2e region.
`--cold-start` skips the prologue.

Timing: W untimed warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events on
the engine's own stream, max over ranks.  Every tick streams a ~2.5 GB working set (activations of
up to 32768 leaves), far above the 126 MB L2, so no explicit L2 flush is needed ("inputs larger
than L2").  `e2e` repeats the measurement through the public Python API with host buffers: weights
go host->device every step (pinned source); the finished games' (state, pi, z) samples are packed on
the device, all-gathered over NCCL (N > 1) and copied device->host on rank 0 (where a learner would
live; at N = 1 that is the rank's own drain into pinned memory), and the counters come back every step.

Extras on one GPU (each a short run on a fresh engine, skipped with --no-extras): `ckpt_workload` = the same workload
on the reference's trained checkpoint with its resignation rule (-0.88, 10 % of the games with resignation disabled:
short games, many game ends per simulation); `parity_tower` = the same workload on the split-bf16 tensor-core tower, the
precision whose pi matches the reference CPU path within 1e-3; `oracle_replay` = finished games of the timed run replayed
through the oracle's board engine (checker only).  `--strong` shards a TOTAL of 4096 games over the ranks (configs[2]).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (game, board, games/GPU, sims, parallel, blocks, filters, fc, warm_up_steps, check_resign_after)
    'go9_c2': ('go', 9, 4096, 400, 8, 10, 128, 128, 16, 40),
    'gomoku13_c4': ('gomoku', 13, 1024, 200, 8, 6, 64, 64, 16, 0),
    'go19_c5': ('go', 19, 128, 800, 8, 19, 256, 256, 30, 80),
    'go9_tiny': ('go', 9, 256, 64, 8, 2, 64, 64, 8, 20),
}
# total games of the configs that shard ONE population over the GPUs (BASELINE.json configs[2..4]): --strong
TOTAL_GAMES = {'go9_c2': 4096, 'gomoku13_c4': 8192, 'go19_c5': 1024, 'go9_tiny': 256}
# prologue length L (plies) of the age stagger: about the length of a random-init game of the workload
STAGGER = {'go9_c2': 128, 'gomoku13_c4': 64, 'go19_c5': 256, 'go9_tiny': 96}
CKPT_FILE = os.path.join(ROOT, 'tests', 'golden', 'ckpt_go9_154000.npz')  # network tensors of checkpoints/go/9x9/training_steps_154000.ckpt


def stagger_population(eng, G, L, par, warm, chk, sims, resign=(-1.0, 1.0)):
    """Age the freshly begun population: L cheap moves (num_simulations = num_parallel), slot g restarted after g mod L of them,
    then the workload's search parameters are switched in and whatever finished during the prologue is discarded."""
    import numpy as np

    slots = np.arange(G, dtype=np.int32)
    ticks_fast = (par + par + par - 1) // par
    for m in range(L):
        eng.selfplay_tick(ticks_fast)
        eng.selfplay_restart(slots[slots % L == m])
    eng.selfplay_update(warm, chk, resign[0], resign[1], search=dict(num_simulations=sims, num_parallel=par))
    while True:
        games, st, pis, zs = eng.drain_games()
        if not games:
            break


def make_net(game, n, nb, nf, fc):
    import torch

    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm

    torch.manual_seed(123)
    a = n * n + (1 if game == 'go' else 0)
    return randomize_batchnorm(AlphaZeroNet((17, n, n), a, nb, nf, fc, game == 'gomoku')).eval()


def load_ckpt_state_dict():
    import numpy as np
    import torch

    w = np.load(CKPT_FILE)
    return {k: torch.from_numpy(w[k]) for k in w.files if k != 'versions'}


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, pw = [], 0, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's self-play path restated (oracle port), one single-threaded actor per host core
_CPU_STATE = {}


def _cpu_init(wl):
    """Per-process setup of a reference-style actor: single-threaded torch, the workload's network, the oracle evaluator."""
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['MKL_NUM_THREADS'] = '1'
    import torch

    torch.set_num_threads(1)
    from oracle import net as onet

    game, n, _, sims, par, nb, nf, fc, warm, _ = WORKLOADS[wl]
    _CPU_STATE['ev'] = onet.make_eval_func(make_net(game, n, nb, nf, fc).state_dict(), game == 'gomoku')
    _CPU_STATE['env'] = None


def cpu_worker(args):
    """One reference-style actor: the oracle's restatement of play_and_record_one_game (pipeline.py:289-382), single-threaded torch
    CPU net, until the time budget is spent.  The first call ages the actor's game by (seed-dependent) 0..L-1 cheap plies, the
    same stagger the GPU population gets, so that short and long samples see the same mixture of game phases.
    Returns (simulations, moves, seconds)."""
    seed, seconds, wl, n_actors = args
    import numpy as np

    from oracle.boards import GoBoard, GomokuBoard
    from oracle.search import search

    game, n, _, sims, par, nb, nf, fc, warm, _ = WORKLOADS[wl]
    if 'ev' not in _CPU_STATE:
        _cpu_init(wl)
    ev = _CPU_STATE['ev']
    if _CPU_STATE.get('env') is None:
        np.random.seed(seed)
        env = GoBoard(n) if game == 'go' else GomokuBoard(n)
        root = None
        age = ((seed - 1) * STAGGER[wl]) // max(1, n_actors)  # actors spread evenly over 0..L-1 plies
        for _ in range(age):
            if env.is_game_over():
                break
            mv, _, _, _, root, _ = search(env, ev, root, 19652.0, 1.25, par, par, True, env.steps <= warm, False)
            env.step(mv)
        _CPU_STATE['env'], _CPU_STATE['root'] = env, (None if env.is_game_over() else root)
    env, root = _CPU_STATE['env'], _CPU_STATE['root']
    total, moves = 0.0, 0
    t0 = time.time()
    while time.time() - t0 < seconds:
        if env.is_game_over():
            env.reset()
            root = None
        before = float(root.tree.root_N) if root is not None else 0.0
        mv, pi, rq, cq, root, child_N = search(env, ev, root, 19652.0, 1.25, sims, par, True, env.steps <= warm, False)
        total += float(child_N.sum()) + 1.0 - before
        moves += 1
        env.step(mv)
    _CPU_STATE['root'] = root
    return total, moves, time.time() - t0


class CpuPool:
    """One pool of actor processes for the whole run (process start, torch import and the age stagger are paid once, outside the
    timed samples)."""

    def __init__(self, wl, cores=None):
        import multiprocessing as mp

        self.wl = wl
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context('spawn').Pool(self.cores, initializer=_cpu_init, initargs=(wl,))

    def run(self, seconds):
        res = self.pool.map(cpu_worker, [(1 + i, seconds, self.wl, self.cores) for i in range(self.cores)], chunksize=1)
        wall = max(r[2] for r in res)
        return sum(r[0] for r in res) / wall, sum(r[1] for r in res) / wall, self.cores, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_measure(wl, steps, warmup, per_step):
    """The ONE procedure behind both CPU numbers (`cpu_baseline` of the default run and the `--impl reference` arm): staggered
    actors, `warmup` untimed 1.5 s samples, then `steps` samples of `per_step` seconds; mean simulations/s and moves/s."""
    pool = CpuPool(wl)
    try:
        for _ in range(max(1, warmup)):
            pool.run(1.5)
        vals, mv, cores = [], [], pool.cores
        for _ in range(steps):
            s, m, cores, wall = pool.run(per_step)
            vals.append(s)
            mv.append(m)
        return sum(vals) / len(vals), sum(mv) / len(mv), cores
    finally:
        pool.close()


def reference_arm(a):
    """--impl reference: the reference's CPU self-play path (oracle port: numpy PUCT + Python board engine +
    torch-CPU fp32 net, one single-threaded actor per host core like training_go.py:12,19,319) on the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = float(os.environ.get('AZ_REF_SECONDS', '12'))
    t0 = time.time()
    v, mvs, cores = cpu_measure(a.workload, a.steps, a.warmup, per_step)
    game, n, G, sims, par, nb, nf, fc, _, _ = WORKLOADS[a.workload]
    line = {
        'impl': 'reference', 'metric': 'mcts_simulations_per_sec', 'value': v, 'unit': 'simulations/s', 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1000.0 * (time.time() - t0) / max(1, a.steps), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{a.workload}: {n}x{n} {game}, {sims} sims/move, num_parallel {par}, net {nb}x{nf} fc{fc}; one single-threaded actor process per host core, '
                               f'game ages staggered over 0..{STAGGER[a.workload] - 1} plies'},
        'moves_per_sec': mvs,
        'cpu_baseline': {'value': v, 'unit': 'simulations/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{cores} actor processes x {per_step:.0f} s of self-play per step x {a.steps} steps (oracle port of the reference CPU path, torch {__import__("torch").__version__} CPU fp32)'},
        'e2e': {'value': v, 'unit': 'simulations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def oracle_replay(game, n, games, states, pis, zs, moves, limit):
    """Checker (oracle/ is test infrastructure): finished games of the measured run replayed through the oracle's board engine —
    every recorded observation, legality of every move, the result and the z labels (pipeline.py:349-354)."""
    import numpy as np

    from oracle.boards import GoBoard, GomokuBoard

    checked = 0
    for rec in games[:limit]:
        env = GoBoard(n) if game == 'go' else GomokuBoard(n)
        s0, ln = rec['first_sample'], rec['game_length']
        obs = env.reset()
        movers, reward, done = [], 0.0, False
        for i in range(ln):
            if done or not np.array_equal(states[s0 + i], obs):
                return {'games': checked, 'ok': False, 'first_bad': f'game in slot {rec["slot"]} ply {i}: observation differs'}
            mv = int(moves[s0 + i])
            if mv >= 0 and env.legal_actions[mv] != 1:
                return {'games': checked, 'ok': False, 'first_bad': f'game in slot {rec["slot"]} ply {i}: illegal move {mv}'}
            movers.append(env.to_play)
            obs, reward, done, _ = env.step(mv)
        want = np.zeros(ln, dtype=np.float32) if reward == 0.0 else np.array([reward if p == env.last_player else -reward for p in movers], dtype=np.float32)
        if not done or (0 if env.winner is None else env.winner) != rec['winner'] or not np.array_equal(zs[s0:s0 + ln], want):
            return {'games': checked, 'ok': False, 'first_bad': f'game in slot {rec["slot"]}: result / z labels differ'}
        checked += 1
    return {'games': checked, 'ok': True}


def short_run(game, n, G, sims, par, net_dims, sd, precision, warm, chk, resign, L, ticks, warmup, steps, device, seed, stream_of):
    """A short measurement on a fresh engine (extras): staggered population, `warmup` + `steps` steps, device-timed."""
    import torch

    from alpha_zero_b200.engine import Engine

    eng = Engine(game, n, num_games=G, max_simulations=sims, max_parallel=par, net=net_dims, precision=precision, device=device, seed=seed)
    eng.set_weights(sd)
    # the resign lottery of a game is drawn when it starts (pipeline.py:244-246): the prologue runs under the workload's threshold
    # and ratio, with the resign CHECK pushed out of reach, so that the aged population has the workload's mix of resign-enabled games
    eng.selfplay_begin(par, par, warm_up_steps=warm, check_resign_after_steps=1 << 20, resign_threshold=resign[0], disable_resign_ratio=resign[1])
    stagger_population(eng, G, L, par, warm, chk, sims, resign)
    stream = stream_of(eng)
    for _ in range(warmup):
        eng.selfplay_tick(ticks)
    while eng.drain_games()[0]:
        pass
    eng.sync()
    c0 = eng.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    lens, resigned, n_games = 0, 0, 0
    for _ in range(steps):
        eng.selfplay_tick(ticks)
    ev1.record(stream)
    eng.sync()
    ms = ev0.elapsed_time(ev1)
    c1 = eng.counters()
    while True:
        games, st, pis, zs = eng.drain_games()
        if not games:
            break
        n_games += len(games)
        lens += sum(g['game_length'] for g in games)
        resigned += sum(1 for g in games if g['by_resign'])
    out = {'value': (c1['simulations'] - c0['simulations']) / (ms * 1e-3), 'unit': 'simulations/s', 'steps': steps, 'ms_per_step': ms / steps,
           'moves_per_sec': (c1['moves'] - c0['moves']) / (ms * 1e-3), 'games_per_sec': (c1['games'] - c0['games']) / (ms * 1e-3),
           'games_finished': n_games, 'mean_game_length': lens / max(1, n_games), 'resigned_fraction': resigned / max(1, n_games),
           'mean_leaf_depth': (c1['depth_sum'] - c0['depth_sum']) / max(1, c1['descents'] - c0['descents']), 'device_errors': c1['errors'],
           'ring_dropped': c1['ring_dropped']}
    eng.close()
    return out


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='go9_c2', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'bf16x3', 'fp32'])
    ap.add_argument('--games', type=int, default=0, help='override games per GPU')
    ap.add_argument('--strong', action='store_true', help="shard the config's TOTAL game count over the GPUs (BASELINE configs[2..4]) instead of keeping games per GPU fixed")
    ap.add_argument('--weights', default='random', choices=['random', 'ckpt'], help='ckpt: go9_c2 on the reference checkpoint 154000 with its resignation rule')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip ckpt_workload / parity_tower / oracle_replay')
    ap.add_argument('--cold-start', action='store_true', help='all games start at move 0 (no age stagger)')
    a = ap.parse_args()
    if a.impl == 'reference':
        return reference_arm(a)

    import numpy as np  # noqa: F401
    import torch

    from alpha_zero_b200.engine import Engine

    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    game, n, G, sims, par, nb, nf, fc, warm, chk = WORKLOADS[a.workload]
    if a.strong:
        from alpha_zero_b200.gather import shard_slots

        lo, hi = shard_slots(TOTAL_GAMES[a.workload], rank, world)
        G = hi - lo
    if a.games:
        G = a.games
    ticks = (sims + par + par - 1) // par  # leaf batches per move
    resign = (-1.0, 1.0)
    if a.weights == 'ckpt':
        assert a.workload == 'go9_c2', 'the shipped checkpoint is a 9x9 Go 10x128 net'
        sd = load_ckpt_state_dict()
        resign = (-0.88, 0.1)  # training_go.py:122,139 defaults
    else:
        sd = make_net(game, n, nb, nf, fc).state_dict()
    pinned = {k: v.pin_memory() for k, v in sd.items() if not k.endswith('num_batches_tracked')}
    eng = Engine(game, n, num_games=G, max_simulations=sims, max_parallel=par, net=(nb, nf, fc), precision=a.precision, device=local, seed=1 + rank)
    eng.set_weights(pinned)
    info = eng.net_info()

    def stream_of(e):
        return torch.cuda.ExternalStream(e.stream(), device=torch.device('cuda', local))

    stream = stream_of(eng)
    L = 0 if a.cold_start else min(STAGGER[a.workload], max(1, G))
    if L:
        eng.selfplay_begin(par, par, warm_up_steps=warm, check_resign_after_steps=1 << 20, resign_threshold=resign[0], disable_resign_ratio=resign[1])
        stagger_population(eng, G, L, par, warm, chk, sims, resign)
    else:
        eng.selfplay_begin(sims, par, warm_up_steps=warm, check_resign_after_steps=chk, resign_threshold=resign[0], disable_resign_ratio=resign[1])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gatherer = None
    if dist is not None:
        from alpha_zero_b200.gather import DeviceSampleGatherer

        gatherer = DeviceSampleGatherer(eng, capacity=max(4096, 4 * G), device=f'cuda:{local}')

    # ---- warm-up --------------------------------------------------------------------------------
    for _ in range(a.warmup):
        eng.selfplay_tick(ticks)
    eng.sync()
    barrier()

    # ---- device-resident measurement: `value` --------------------------------------------------
    c0 = eng.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record(stream)
        for _ in range(a.steps):
            eng.selfplay_tick(ticks)
        ev1.record(stream)
        eng.sync()
        barrier()
    ms = ev0.elapsed_time(ev1)
    c1 = eng.counters()
    tower_ms, tower_evals = eng.last_net_ms()

    # ---- end to end through the public API with host buffers: `e2e` ----------------------------
    d2h_bytes, gathered, e2e_games, e2e_len = 0, 0, 0, 0
    kept = []  # (records, states, pis, zs, moves) of this rank's first drained games, for the oracle replay
    barrier()
    t0 = time.perf_counter()
    ce0 = eng.counters()
    for _ in range(a.steps):
        eng.set_weights(pinned)  # host -> device: the step's network parameters (ckpt hot-swap, pipeline.py:232-239)
        eng.selfplay_tick(ticks)
        if gatherer is None:
            games, st, pis, zs = eng.drain_games()  # device -> pinned host: finished games' (state, pi, z)
            d2h_bytes += st.nbytes + pis.nbytes + zs.nbytes
            gathered += len(zs)
            if not kept and games:
                kept.append((games, st.copy(), pis.copy(), zs.copy(), eng.last_moves.copy()))
        else:
            games, got = gatherer.push()  # pack on the device, NCCL all-gather of exactly the produced rows
            gathered += len(got)
            if rank == 0:  # the learner's rank takes everybody's samples to the host
                S, P, Z = got.to_host()
                d2h_bytes += S.nbytes + P.nbytes + Z.nbytes
        eng.counters()  # device -> host: the step's result counters
        d2h_bytes += 12 * 8
        e2e_games += len(games)
        e2e_len += sum(g['game_length'] for g in games)
    barrier()
    e2e_s = time.perf_counter() - t0
    ce1 = eng.counters()

    # ---- per-phase split of the tick (one extra, untimed step with CUDA events around every phase) ----
    eng.tick_profile(True)
    eng.selfplay_tick(ticks)
    eng.sync()
    phase_ms, phase_ticks = eng.tick_profile(False)

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_max = allmax(ms)
    e2e_max = allmax(e2e_s)
    d = {k: allsum(float(c1[k] - c0[k])) for k in ('simulations', 'evaluations', 'moves', 'games', 'descents', 'depth_sum')}
    e2e_sims = allsum(float(ce1['simulations'] - ce0['simulations']))
    e2e_moves = allsum(float(ce1['moves'] - ce0['moves']))
    e2e_games_all, e2e_len_all = allsum(float(e2e_games)), allsum(float(e2e_len))
    launches = c1['kernel_launches'] - c0['kernel_launches']
    errors = allsum(float(ce1['errors']))
    dropped = allsum(float(ce1['ring_dropped']))
    G_all = allsum(float(G))
    eng.close()
    del eng

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hw = n * n if game == 'go' else (n + 4) * (n + 4)
        n_conv = 1 + 2 * nb
        # algorithmic 2*MAC of the tower per leaf, layer by layer: the input layer has 17 input planes, not nf
        tower_flops = 2.0 * hw * 9 * 17 * nf + (n_conv - 1) * 2.0 * hw * 9 * nf * nf
        # tower_ms = mean over the last 64 ticks of the timed region; leaves per tick = this rank's evaluations of the region / its ticks
        tower_evals = int(round((c1['evaluations'] - c0['evaluations']) / max(1, a.steps * ticks))) or tower_evals
        achieved = (tower_evals * tower_flops / (tower_ms * 1e-3) / 1e12) if tower_ms > 0 else None
        is_tc = a.precision in ('bf16', 'bf16x3')
        peak = peaks.get('bf16_tflops_sustained') if is_tc else None
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peak else 'fallback 1400 TF/s (of fallback)'
        peak = peak or 1400.0
        # DRAM traffic of the dominant kernel per launch: ncu --set full of THIS build's kernel (profiles/r02_conv_traffic.json, written
        # from the committed raw page by tools/ncu_traffic.py), mean over the tower's plain and residual launches, scaled to this tick's leaves
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'r02_conv_traffic.json')))
            ent = tj.get(f'{a.workload}/{a.precision}/mode{info["tc_mode"]}')
            if ent:
                traffic = (nb * ent['bytes_residual'] + (nb + 1) * ent['bytes_plain']) / n_conv * (tower_evals / float(ent['leaves']))
                traffic_src = ent['source']
        except Exception:
            pass
        value = d['simulations'] / (ms_max * 1e-3)
        mean_len = e2e_len_all / max(1.0, e2e_games_all)
        kern = {5: 'k_conv_tc_x<pair>', 6: 'k_conv_tc_x<single>', 4: 'k_conv_tc_halo<pair>', 2: 'k_conv_tc_halo<single>', 0: 'k_conv_tc', -1: 'k_conv_f32'}.get(info['tc_mode'], 'k_conv_tc_halo')
        line = {
            'metric': 'mcts_simulations_per_sec', 'value': value, 'unit': 'simulations/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_max / a.steps, 'higher_is_better': True, 'scaling': 'strong' if a.strong else 'weak', 'vs_baseline': None,
            'dtype': a.precision if a.precision != 'fp32' else 'f32', 'data': 'synthetic',
            'config': {'workload': f'{a.workload}: {n}x{n} {game}, ' + (f'{int(G_all)} concurrent games in total sharded over {world} GPU(s) ({G} on rank 0)' if a.strong else f'{G} concurrent games per GPU')
                                   + f', {sims} sims/move, num_parallel {par}, net {nb}x{nf} fc{fc} '
                                   + ('(reference checkpoint go/9x9 training_steps_154000, resign threshold -0.88, 10% of games with resignation disabled)' if a.weights == 'ckpt' else '(random init seed 123, resignation disabled)')
                                   + f'; step = {ticks} leaf batches; ' + (f'game ages staggered over 0..{L - 1} plies by a cheap prologue' if L else 'cold start: all games at move 0')
                                   + '; working set >> L2 (no flush needed)',
                       'games_per_gpu': G, 'tower': f'{a.precision} {kern} (AZ_TC_MODE {info["tc_mode"]}), {info["padded_filters"]} channels',
                       'parallelism': f'games sharded, {world} process(es), device-packed NCCL all-gather of samples only'},
            'evals_per_sec': d['evaluations'] / (ms_max * 1e-3), 'moves_per_sec': d['moves'] / (ms_max * 1e-3),
            'games_per_sec': d['games'] / (ms_max * 1e-3), 'games_finished_in_window': d['games'],
            # a window of a few seconds sees the finishing waves of the stagger prologue; the stationary rate of the fleet is moves/s / mean length
            'games_per_sec_steady': (d['moves'] / (ms_max * 1e-3)) / mean_len if e2e_games_all else None,
            'mean_leaf_depth': d['depth_sum'] / max(1.0, d['descents']), 'device_errors': errors, 'ring_dropped': dropped,
            'e2e': {'value': e2e_sims / e2e_max, 'unit': 'simulations/s', 'h2d_bytes_per_step': int(sum(v.numel() * v.element_size() for v in pinned.values())),
                    'd2h_bytes_per_step': int(d2h_bytes / max(1, a.steps)), 'samples_all_gathered': gathered,
                    'games_drained': e2e_games_all, 'games_per_sec': e2e_games_all / e2e_max, 'moves_per_sec': e2e_moves / e2e_max,
                    'mean_game_length': mean_len},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'kernel': kern, 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': (achieved / peak) if achieved else None, 'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu dram__bytes_read+write)',
                         'traffic_source': traffic_src,
                         'algorithmic_bytes_per_launch': 2.0 * tower_evals * hw * nf * 2, 'peak_source': peak_src,
                         'note': f'algorithmic 2*MAC of the {n_conv} tower layers ({tower_flops / 1e6:.1f} MFLOP/leaf: input layer 17 planes, the rest {nf}x{nf}) x {tower_evals} leaves (mean per tick) / '
                                 f'mean tower time per tick over the last 64 ticks of the timed region ({tower_ms:.3f} ms for the {n_conv} conv launches, CUDA events on the engine stream around them)'
                                 + ('; the split tower issues 3x these MMAs' if a.precision == 'bf16x3' else '')},
            'clocks': clk.summary(),
            'tick_breakdown_ms': dict({k: v / max(1, phase_ticks) for k, v in phase_ms.items()}, ticks=phase_ticks,
                                      note='mean device time per tick of one extra step, CUDA events around each phase (az_tick_profile); '
                                           'select = k_collect + k_compact, network = input + tower + heads, expand_backup = k_apply, move_reroot = k_advance'),
        }
        if world == 1 and not a.no_extras:
            if kept:
                games, st, pis, zs, mv = kept[0]
                line['oracle_replay'] = oracle_replay(game, n, games, st, pis, zs, mv, limit=int(os.environ.get('AZ_REPLAY_GAMES', '32')))
            if a.workload == 'go9_c2' and a.precision == 'bf16' and a.weights == 'random':
                xs = max(2, min(a.steps, 3))
                if os.path.exists(CKPT_FILE):
                    line['ckpt_workload'] = dict(short_run(game, n, G, sims, par, (nb, nf, fc), load_ckpt_state_dict(), 'bf16', warm, chk, (-0.88, 0.1), L or 1, ticks, 2, xs, local, 1, stream_of),
                                                 workload='same config on the reference checkpoint go/9x9 training_steps_154000 with its resignation rule (-0.88, 10% disabled)')
                line['parity_tower'] = dict(short_run(game, n, G, sims, par, (nb, nf, fc), sd, 'bf16x3', warm, chk, resign, L or 1, ticks, 2, xs, local, 1, stream_of),
                                            dtype='bf16x3', note='same workload on the split-bf16 tcgen05 tower (pi within 1e-3 of the reference CPU path: tests/test_gpu_net_layers.py, smoke)')
        if world == 1 and not a.no_cpu_baseline:
            per_step = float(os.environ.get('AZ_REF_SECONDS', '12'))
            v, mvs, cores = cpu_measure(a.workload, 2, 1, per_step)
            line['cpu_baseline'] = {'value': v, 'unit': 'simulations/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'{cores} single-threaded actor processes, 2 samples of {per_step:.0f} s of the same workload on the same age-staggered mixture as '
                                              '`--impl reference` (oracle port, torch CPU fp32 net)'}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
