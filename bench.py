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
run on games aged 0..L-1 plies, with finished games drained and all-gathered inside the e2e region.
`--cold-start` skips the prologue (the round's earlier numbers were taken that way).

Timing: W untimed warm-up steps, then K steps bracketed by barrier + synchronize, CUDA events on
the engine's own stream, max over ranks.  Every tick streams a ~2.5 GB working set (activations of
up to 32768 leaves), far above the 126 MB L2, so no explicit L2 flush is needed ("inputs larger
than L2").  `e2e` repeats the measurement through the public Python API with host buffers: weights
go host->device every step (pinned source), finished games' (state, pi, z) samples and the counters
come back device->host every step.
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
# prologue length L (plies) of the age stagger: about the length of a random-init game of the workload
STAGGER = {'go9_c2': 128, 'gomoku13_c4': 64, 'go19_c5': 256, 'go9_tiny': 96}


def stagger_population(eng, G, L, par, warm, chk, sims):
    """Age the freshly begun population: L cheap moves (num_simulations = num_parallel), slot g restarted after g mod L of them,
    then the workload's search parameters are switched in and whatever finished during the prologue is discarded."""
    import numpy as np

    slots = np.arange(G, dtype=np.int32)
    ticks_fast = (par + par + par - 1) // par
    for m in range(L):
        eng.selfplay_tick(ticks_fast)
        eng.selfplay_restart(slots[slots % L == m])
    eng.selfplay_update(warm, chk, -1.0, 1.0, search=dict(num_simulations=sims, num_parallel=par))
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
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
def cpu_worker(args):
    """One reference-style actor: the oracle's restatement of play_and_record_one_game (pipeline.py:289-382),
    single-threaded torch CPU net, until the time budget is spent.  Returns (simulations, moves, seconds)."""
    seed, seconds, wl = args
    import numpy as np

    from oracle.boards import GoBoard, GomokuBoard
    from oracle.search import search

    game, n, _, sims, par, nb, nf, fc, warm, _ = WORKLOADS[wl]
    if 'ev' not in _CPU_STATE:
        _cpu_init(wl)
    ev = _CPU_STATE['ev']
    if _CPU_STATE.get('env') is None:
        np.random.seed(seed)
        _CPU_STATE['env'] = GoBoard(n) if game == 'go' else GomokuBoard(n)
        _CPU_STATE['root'] = None
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


class CpuPool:
    """One pool of actor processes for the whole run (process start + torch import are paid once, outside the timed samples)."""

    def __init__(self, wl, cores=None):
        import multiprocessing as mp

        self.wl = wl
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context('spawn').Pool(self.cores, initializer=_cpu_init, initargs=(wl,))

    def run(self, seconds):
        res = self.pool.map(cpu_worker, [(1 + i, seconds, self.wl) for i in range(self.cores)])
        wall = max(r[2] for r in res)
        return sum(r[0] for r in res) / wall, sum(r[1] for r in res) / wall, self.cores, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def run_cpu(wl, seconds, cores=None):
    pool = CpuPool(wl, cores)
    try:
        pool.run(1.0)  # untimed: first-touch of the conv kernels in every process
        return pool.run(seconds)
    finally:
        pool.close()


def reference_arm(a):
    """--impl reference: the reference's CPU self-play path (oracle port: numpy PUCT + Python board engine +
    torch-CPU fp32 net, one single-threaded actor per host core like training_go.py:12,19,319) on the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = float(os.environ.get('AZ_REF_SECONDS', '12'))
    pool = CpuPool(a.workload)
    for _ in range(max(1, a.warmup)):
        pool.run(1.5)
    vals, mv = [], []
    t0 = time.time()
    for _ in range(a.steps):
        s, m, cores, wall = pool.run(per_step)
        vals.append(s)
        mv.append(m)
    pool.close()
    v = sum(vals) / len(vals)
    game, n, G, sims, par, nb, nf, fc, _, _ = WORKLOADS[a.workload]
    line = {
        'impl': 'reference', 'metric': 'mcts_simulations_per_sec', 'value': v, 'unit': 'simulations/s', 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1000.0 * (time.time() - t0) / max(1, a.steps), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{a.workload}: {n}x{n} {game}, {sims} sims/move, num_parallel {par}, net {nb}x{nf} fc{fc}; one single-threaded actor process per host core'},
        'moves_per_sec': sum(mv) / len(mv),
        'cpu_baseline': {'value': v, 'unit': 'simulations/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{cores} actor processes x {per_step:.0f} s of self-play per step (oracle port of the reference CPU path, torch {__import__("torch").__version__} CPU fp32)'},
        'e2e': {'value': v, 'unit': 'simulations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='go9_c2', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--games', type=int, default=0, help='override games per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cold-start', action='store_true', help='all games start at move 0 (no age stagger)')
    a = ap.parse_args()
    if a.impl == 'reference':
        return reference_arm(a)

    import numpy as np
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
    if a.games:
        G = a.games
    ticks = (sims + par + par - 1) // par  # leaf batches per move
    net = make_net(game, n, nb, nf, fc)
    sd = net.state_dict()
    pinned = {k: v.pin_memory() for k, v in sd.items() if not k.endswith('num_batches_tracked')}
    eng = Engine(game, n, num_games=G, max_simulations=sims, max_parallel=par, net=(nb, nf, fc), precision=a.precision, device=local, seed=1 + rank)
    eng.set_weights(pinned)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device('cuda', local))
    L = 0 if a.cold_start else min(STAGGER[a.workload], max(1, G))
    if L:
        eng.selfplay_begin(par, par, warm_up_steps=warm, check_resign_after_steps=chk, resign_threshold=-1.0, disable_resign_ratio=1.0)
        stagger_population(eng, G, L, par, warm, chk, sims)
    else:
        eng.selfplay_begin(sims, par, warm_up_steps=warm, check_resign_after_steps=chk, resign_threshold=-1.0, disable_resign_ratio=1.0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gatherer = None
    if dist is not None:
        from alpha_zero_b200.gather import SampleGatherer

        gatherer = SampleGatherer(capacity=G * 2, device=f'cuda:{local}')

    def gather_samples(states, pis, zs):
        """NCCL all-gather of the (state, pi, z) samples produced this step: fixed-capacity blocks + counts; what
        does not fit the block waits for the next step."""
        if gatherer is None:
            return len(zs)
        return len(gatherer.push(states, pis, zs)[2])

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
    barrier()
    t0 = time.perf_counter()
    ce0 = eng.counters()
    for _ in range(a.steps):
        eng.set_weights(pinned)  # host -> device: the step's network parameters (ckpt hot-swap, pipeline.py:232-239)
        eng.selfplay_tick(ticks)
        games, st, pis, zs = eng.drain_games()  # device -> host: finished games' (state, pi, z)
        ce = eng.counters()  # device -> host: the step's result counters
        d2h_bytes += st.nbytes + pis.nbytes + zs.nbytes + 12 * 8
        e2e_games += len(games)
        e2e_len += sum(g['game_length'] for g in games)
        gathered += gather_samples(st, pis, zs)
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
    e2e_games_all, e2e_len_all = allsum(float(e2e_games)), allsum(float(e2e_len))
    launches = c1['kernel_launches'] - c0['kernel_launches']
    errors = allsum(float(c1['errors']))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hw = n * n if game == 'go' else (n + 4) * (n + 4)
        conv_flops_per_eval = 2.0 * hw * 9 * nf * nf  # one res-tower 3x3 conv layer, 2*MAC per leaf
        n_conv = 1 + 2 * nb
        launch_ms = tower_ms / n_conv if tower_ms > 0 else None
        achieved = (tower_evals * conv_flops_per_eval / (launch_ms * 1e-3) / 1e12) if launch_ms else None
        peak = peaks.get('bf16_tflops_sustained') if a.precision == 'bf16' else None
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peak else 'fallback 1400 TF/s (of fallback)'
        peak = peak or 1400.0
        # DRAM traffic of the dominant kernel per launch: from the committed `ncu --set full` captures (halo kernel: profiles/
        # r01_ncu_k_conv_tc_halo_pair_raw.csv, 16384 leaves: dram read+write 1209.1 MB with / 790.0 MB without the residual add),
        # mean over the tower's 10 residual + 11 plain launches, scaled linearly to this tick's leaf count
        traffic = None
        tc_mode = os.environ.get('AZ_TC_MODE', '5')
        if a.workload == 'go9_c2' and a.precision == 'bf16' and tc_mode == '4':
            traffic = (10 * 1209.1e6 + 11 * 790.0e6) / 21 * (tower_evals / 16384.0)
        if a.workload == 'go9_c2' and a.precision == 'bf16' and tc_mode == '5':
            # profiles/r01_ncu_k_conv_tc_x_dense_raw.csv, 16384 leaves: 1088.6 MB with / 711.8 MB without the residual add
            traffic = (10 * 1088.6e6 + 11 * 711.8e6) / 21 * (tower_evals / 16384.0)
        value = d['simulations'] / (ms_max * 1e-3)
        line = {
            'metric': 'mcts_simulations_per_sec', 'value': value, 'unit': 'simulations/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_max / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': a.precision if a.precision != 'fp32' else 'f32', 'data': 'synthetic',
            'config': {'workload': f'{a.workload}: {n}x{n} {game}, {G} concurrent games per GPU, {sims} sims/move, num_parallel {par}, net {nb}x{nf} fc{fc} '
                                   f'(random init seed 123); step = {ticks} leaf batches; ' + (f'game ages staggered over 0..{L - 1} plies by a cheap prologue' if L else 'cold start: all games at move 0')
                                   + '; working set >> L2 (no flush needed)',
                       'games_per_gpu': G, 'parallelism': f'games sharded, {world} process(es), NCCL all-gather of samples only'},
            'evals_per_sec': d['evaluations'] / (ms_max * 1e-3), 'moves_per_sec': d['moves'] / (ms_max * 1e-3),
            'games_per_sec': d['games'] / (ms_max * 1e-3), 'games_finished_in_window': d['games'], 'mean_leaf_depth': d['depth_sum'] / max(1.0, d['descents']), 'device_errors': errors,
            'e2e': {'value': e2e_sims / e2e_max, 'unit': 'simulations/s', 'h2d_bytes_per_step': int(eng.weight_bytes),
                    'd2h_bytes_per_step': int(d2h_bytes / max(1, a.steps)), 'samples_all_gathered': gathered,
                    'games_drained': e2e_games_all, 'games_per_sec': e2e_games_all / e2e_max,
                    'mean_game_length': e2e_len_all / max(1.0, e2e_games_all)},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'kernel': ({'5': 'k_conv_tc_x', '6': 'k_conv_tc_x', '0': 'k_conv_tc'}.get(tc_mode, 'k_conv_tc_halo') if nf <= 128 else 'k_conv_tc') if a.precision == 'bf16' else 'k_conv_f32', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': (achieved / peak) if achieved else None, 'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu dram__bytes_read+write)',
                         'algorithmic_bytes_per_launch': 2.0 * tower_evals * (n * n if game == 'go' else (n + 4) ** 2) * nf * 2, 'peak_source': peak_src,
                         'note': f'algorithmic 2*MAC of one 3x3 conv layer ({conv_flops_per_eval / 1e6:.1f} MFLOP/leaf) x {tower_evals} leaves / mean launch time over the '
                                 f'{n_conv} tower launches of the last tick ({tower_ms:.3f} ms, CUDA events on the engine stream)'},
            'clocks': clk.summary(),
            'tick_breakdown_ms': dict({k: v / max(1, phase_ticks) for k, v in phase_ms.items()}, ticks=phase_ticks,
                                      note='mean device time per tick of one extra step, CUDA events around each phase (az_tick_profile); '
                                           'select = k_collect + k_compact, network = input + tower + heads, expand_backup = k_apply, move_reroot = k_advance'),
        }
        if world == 1 and not a.no_cpu_baseline:
            secs = float(os.environ.get('AZ_CPU_SECONDS', '15'))
            v, mvs, cores, wall = run_cpu(a.workload, secs)
            line['cpu_baseline'] = {'value': v, 'unit': 'simulations/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'{cores} single-threaded actor processes x {secs:.0f} s of the same workload (oracle port, torch CPU fp32 net)'}
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
