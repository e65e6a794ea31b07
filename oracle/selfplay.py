"""ORACLE (test infrastructure, never shipped): CPU restatement of the self-play game loop.

play_one_game follows /root/reference/alpha_zero/core/pipeline.py:289-382 (search per ply with subtree
re-use, resign marking, z assignment, stats dict); used to pin the pipeline semantics against
tests/golden/pipeline_*.npz and as the timed CPU arm of bench.py (`--impl reference`, cpu_baseline).
"""
from oracle.search import search


def play_one_game(env, eval_func, num_simulations, num_parallel, resign_disabled, c_puct_base, c_puct_init, warm_up_steps,
                  check_resign_after_steps, resign_threshold, counters=None):
    obs = env.reset()
    states, pis, to_plays = [], [], []
    root, done, reward = None, False, 0.0
    marked, num_passes = None, 0
    while not done:
        n_before = float(root.tree.root_N) if root is not None else 0.0
        move, pi, root_q, child_q, root, child_N = search(
            env, eval_func, root, c_puct_base, c_puct_init, num_simulations, num_parallel, root_noise=True,
            warm_up=not env.steps > warm_up_steps, deterministic=False)
        if counters is not None:  # simulations = root visits gained by this search
            counters['sims'] = counters.get('sims', 0) + float(child_N.sum()) + 1.0 - n_before
            counters['moves'] = counters.get('moves', 0) + 1
        states.append(obs)
        pis.append(pi)
        to_plays.append(env.to_play)
        if env.has_resign_move and env.steps > check_resign_after_steps and root_q < resign_threshold and child_q < resign_threshold:
            if marked is None:
                marked = env.to_play
            if not resign_disabled:
                move = env.resign_move
        obs, reward, done, _ = env.step(move)
        if env.has_pass_move and move == env.pass_move:
            num_passes += 1
    values = [0.0] * len(states)
    if reward != 0.0:  # pipeline.py:349-354
        values = [reward if p == env.last_player else -reward for p in to_plays]
    stats = {'game_length': len(states), 'game_result': env.get_result_string()}
    if env.has_pass_move:
        stats['num_passes'] = num_passes
    if env.has_resign_move:
        is_marked = resign_disabled and marked is not None
        stats['is_resign_disabled'] = resign_disabled
        stats['is_marked_for_resign'] = is_marked
        stats['is_could_won'] = bool(is_marked and env.winner == marked)
        stats['marked_resign_player'] = env.get_player_name_by_id(marked)
        stats['resign_threshold'] = resign_threshold
    return list(zip(states, pis, values)), stats
