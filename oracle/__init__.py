"""ORACLE — test infrastructure only.

A CPU restatement (numpy / plain Python / torch-fp32 for the net) of the reference self-play hot path
(michaelnny/alpha_zero: envs/go.py, envs/go_engine.py, envs/gomoku.py, envs/base.py, core/mcts_v2.py,
core/network.py, core/pipeline.py:83-382), pinned against golden vectors produced by running the
unmodified reference in the build container (tests/golden/make_golden.py).

Nothing under alpha_zero_b200/ may import this package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs do, and only as the checker / the CPU arm.
"""
