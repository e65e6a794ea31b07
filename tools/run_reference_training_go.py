#!/usr/bin/env python
"""TEST INFRASTRUCTURE: run the UNMODIFIED reference driver /root/reference/alpha_zero/training_go.py with the one-line import
swap of INTEGRATION.md applied from outside (the file is read-only and stays byte for byte what the reference ships):

    alpha_zero.core.pipeline.run_selfplay_actor_loop  ->  alpha_zero_b200.pipeline.run_selfplay_actor_loop

Everything else is the reference's own code: flags, process wiring (`spawn`, mp.Queue, Manager values, Events), the learner
(UniformReplay, compute_losses, SGD, checkpoint writer, resign controller) and the evaluator (its own MCTS + Elo).  The actor
process is ours: one process driving $AZ_ACTOR_GAMES concurrent games on an engine.

    python tools/run_reference_training_go.py --emu [--driver training_gomoku.py] [--out DIR] [driver flags...]

`--emu` makes the engine the HOST EMULATION build (tests/emu/libaz_emu.so: same game / tree / host-ABI code compiled for the
host, a hash of the observation standing in for the network) — the only way to run this in the build container, which has no GPU;
on the GPU box there is no /root/reference to run.  Without `--emu` the actor opens libaz_b200.so on its CUDA device.
The missing third-party packages of the image (gym, sgf, snappy) come from tests/shims, as for the golden generators.
"""
import os
import runpy
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def main():
    argv = sys.argv[1:]
    emu = '--emu' in argv
    if emu:
        argv.remove('--emu')
    driver = 'training_go.py'
    if '--driver' in argv:  # training_go.py (default) | training_gomoku.py | training_go_jumbo.py: same wiring, other env / defaults
        i = argv.index('--driver')
        driver = argv[i + 1]
        del argv[i:i + 2]
    out = None
    if '--out' in argv:
        i = argv.index('--out')
        out = argv[i + 1]
        del argv[i:i + 2]
    out = out or tempfile.mkdtemp(prefix='az_training_go_')
    for sub in ('ckpt', 'logs', 'sgf'):
        os.makedirs(os.path.join(out, sub), exist_ok=True)
    if emu:
        os.environ['AZ_TEST_EMU_BINDING'] = '1'  # read by tests/shims/gym/__init__.py in every spawned child
    os.environ.setdefault('AZ_ACTOR_GAMES', '64')
    os.environ.setdefault('AZ_NET_PRECISION', 'fp32')
    for p in (ROOT, REF, os.path.join(ROOT, 'tests', 'shims'), os.path.join(ROOT, 'tests', 'emu')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['PYTHONPATH'] = os.pathsep.join([os.path.join(ROOT, 'tests', 'shims'), os.path.join(ROOT, 'tests', 'emu'), REF, ROOT, os.environ.get('PYTHONPATH', '')])
    defaults = [
        '--num_actors=1', '--num_res_blocks=1', '--num_filters=16', '--num_fc_units=16', '--num_simulations=8', '--num_parallel=4',
        # the learner asserts min_games, games_per_ckpt, ckpt_interval, log_interval >= 100 (core/pipeline.py:424-429)
        '--min_games=100', '--games_per_ckpt=100', '--ckpt_interval=100', '--max_training_steps=300', '--batch_size=32', '--log_interval=100',
        '--replay_capacity=20000', '--init_resign_threshold=-1', '--eval_games_dir=/nonexistent', '--save_sgf_interval=5',
        f'--ckpt_dir={out}/ckpt', f'--logs_dir={out}/logs', f'--save_sgf_dir={out}/sgf', '--log_level=DEBUG',
    ]
    sys.argv = [os.path.join(REF, 'alpha_zero', driver)] + defaults + argv
    # go_engine.py reads BOARD_SIZE when it is first imported (go_engine.py:31); training_go.py sets it from its flag before ITS
    # imports (training_go.py:204-208), and so must whoever imports the reference earlier
    board = [a.split('=')[1] for a in sys.argv if a.startswith('--board_size=')]
    os.environ['BOARD_SIZE'] = board[-1] if board else ('19' if 'jumbo' in driver else '9')
    import gym  # noqa: F401  (the shim; installs the emulation binding when AZ_TEST_EMU_BINDING is set)
    import alpha_zero.core.pipeline as ref_pipeline
    from alpha_zero_b200.pipeline import run_selfplay_actor_loop

    ref_pipeline.run_selfplay_actor_loop = run_selfplay_actor_loop  # THE import swap
    print(f'[import swap] {driver} will import run_selfplay_actor_loop from {run_selfplay_actor_loop.__module__}; output in {out}', flush=True)
    runpy.run_path(sys.argv[0], run_name='__main__')


if __name__ == '__main__':
    main()
