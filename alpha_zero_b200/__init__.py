"""alpha_zero_b200 — B200-native AlphaZero self-play engine behind the michaelnny/alpha_zero API.

    from alpha_zero_b200.envs.go import GoEnv             # BoardGameEnv contract (envs/base.py:26)
    from alpha_zero_b200.envs.gomoku import GomokuEnv
    from alpha_zero_b200.mcts import Node, uct_search, parallel_uct_search      # core/mcts_v2.py:65,301,485
    from alpha_zero_b200.pipeline import create_mcts_player, play_and_record_one_game, run_selfplay_actor_loop

Everything computes in alpha_zero_b200/libaz_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/az_engine.h).  There is no CPU fallback.
"""
__version__ = '0.1.0'
