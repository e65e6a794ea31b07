"""The UNMODIFIED reference driver (alpha_zero/training_go.py: flags, spawn wiring, learner, evaluator) with the one-line import
swap of INTEGRATION.md applied from outside, the actor's engine being the host-emulation build (tools/run_reference_training_go.py
--emu).  Three checkpoints are trained, the actor switches to each, the run shuts down cleanly.

Opt-in (AZ_RUN_TRAINING_DRIVER=1): the reference's learner sleeps 90 s before it returns (core/pipeline.py:623-626), so the run
takes 2.5 minutes; the log of such a run is kept in profiles/r02_training_go_import_swap_emu.log.  Needs /root/reference."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(os.environ.get('AZ_RUN_TRAINING_DRIVER') != '1' or not os.path.isdir('/root/reference'),
                    reason='opt-in (AZ_RUN_TRAINING_DRIVER=1) and needs the reference checkout')
def test_unmodified_training_go_with_import_swap(tmp_path):
    out = str(tmp_path / 'run')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_reference_training_go.py'), '--emu', '--out', out], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=900)
    log = r.stdout + r.stderr
    assert r.returncode == 0, log[-3000:]
    assert 'Traceback' not in log, log[-3000:]
    assert sorted(os.listdir(os.path.join(out, 'ckpt'))) == ['training_steps_100.ckpt', 'training_steps_200.ckpt', 'training_steps_300.ckpt']
    for step in (100, 200, 300):
        assert f'Actor0 switched to checkpoint "{out}/ckpt/training_steps_{step}.ckpt"' in log
    assert 'Actor0 received stop signal.' in log
    rows = open(os.path.join(out, 'logs', 'actor0.csv')).read().strip().splitlines()
    assert len(rows) > 300 and 'training_steps' in rows[0]
    steps = {line.split(',')[-1] for line in rows[1:]}
    assert {'0', '100', '200'} <= steps  # every emitted game names the weight set it was played with
    assert len(open(os.path.join(out, 'logs', 'training.csv')).read().strip().splitlines()) == 4
