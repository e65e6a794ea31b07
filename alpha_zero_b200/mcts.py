"""Node / uct_search / parallel_uct_search with the reference's signatures (alpha_zero/core/mcts_v2.py:65,301,485).

The tree lives in HBM next to the env's game slot; select / expand / backup run in csrc/az_tree.cuh.  What
stays on the host is exactly what has to consume numpy's global RNG in the reference's order, so that a
seeded run reproduces the reference move for move: the Dirichlet draw (mcts_v2.py:259-260), the
search-policy arithmetic (:265-298, same numpy expressions, same dtypes) and the move sampling (:640-641).
`eval_func` is the reference's callback contract (pipeline.py:91-123); pass an `EngineEvaluator`
(pipeline.create_mcts_player does) to keep the leaves on the device instead.
"""
import numpy as np

from .envs.base import BoardGameEnv


class Node:
    """Handle on the (re-rooted) search tree of an env slot.  Opaque for the pipeline, which only hands it
    back (pipeline.py:308-321); the statistics of the reference's Node are exposed read-only."""

    def __init__(self, to_play=None, num_actions=None, move=None, parent=None, _env=None, _gen=None):
        self.to_play = to_play
        self.num_actions = num_actions
        self.move = move
        self.parent = parent
        self.is_expanded = _env is not None
        self._env, self._gen = _env, _gen
        self.N = self.W = 0.0

    @property
    def Q(self):
        return self.W / self.N if self.N > 0 else 0.0

    @property
    def has_parent(self):
        return isinstance(self.parent, Node)


def generate_search_policy(child_N, temperature, legal_actions):
    """mcts_v2.py:265-298 with the same numpy expressions (so dtypes and rounding are the reference's)."""
    if not isinstance(temperature, float) or not 0 < temperature <= 1.0:
        raise ValueError(f'Expect `temperature` to be float type in the range (0.0, 1.0], got {temperature}')
    child_N = legal_actions * child_N
    child_N = np.power(child_N, max(1.0, min(5.0, 1.0 / temperature)))
    sums = np.sum(child_N)
    if sums > 0:
        child_N /= sums
    return child_N


class EngineEvaluator:
    """Marker for evaluators that run inside the engine (leaves never leave the device)."""

    def __init__(self, engine_net):
        self.engine_net = engine_net

    def __call__(self, obs, batched=False):
        st = obs if batched else obs[None, ...]
        pri, val = self.engine_net.net_forward(st)
        pis = [pri[i] for i in range(pri.shape[0])]
        vs = [float(v) for v in val]
        return (pis, vs) if batched else (pis[0], vs[0])


def _search(env, eval_func, root_node, c_puct_base, c_puct_init, num_simulations, num_parallel, root_noise, warm_up, deterministic):
    if not isinstance(env, BoardGameEnv):
        raise ValueError(f'Expect `env` to be a valid BoardGameEnv instance, got {env}')
    if not 1 <= num_simulations:
        raise ValueError(f'Expect `num_simulations` to a positive integer, got {num_simulations}')
    if env.is_game_over():
        raise RuntimeError('Game is over.')
    eng, slot = env.engine, env.slot
    reuse = False
    if root_node is not None:
        if not isinstance(root_node, Node) or root_node._env is not env or root_node._gen != env._tree_gen:
            raise ValueError('`root_node` does not belong to the current position of `env`')
        assert root_node.to_play == env.to_play
        reuse = True
    root_legal = env.legal_actions
    noise = None
    if root_noise:  # same call, same argument as mcts_v2.py:259-260 -> same RNG stream
        alphas = np.ones_like(root_legal) * 0.03
        noise = np.random.dirichlet(alphas)
    eng.search_begin([slot], [1 if reuse else 0], float(c_puct_base), float(c_puct_init), int(num_simulations), int(num_parallel),
                     bool(root_noise), bool(warm_up), bool(deterministic), noise=noise)
    need_root = not reuse
    while True:
        obs, counts, active = eng.search_select()
        if active == 0:
            break
        if len(obs) == 0:
            eng.search_apply(None, None)
            continue
        if need_root or num_parallel <= 1:
            # root evaluation (mcts_v2.py:364, :555) and every leaf of the serial search (:414) are unbatched calls
            p, v = eval_func(obs[0], False)
            pri, val = [p], [v]
            need_root = False
        else:
            pri, val = eval_func(obs, True)  # mcts_v2.py:614
        eng.search_apply(np.stack(pri), np.asarray(val, dtype=np.float32))
    res = eng.search_result(slot)
    search_pi = generate_search_policy(res['child_N'], 1.0 if warm_up else 0.1, root_legal)
    move = None
    if deterministic:
        move = int(np.argmax(res['child_N']))
    else:
        while move is None or (warm_up and env.has_pass_move and move == env.pass_move) or root_legal[move] != 1:
            move = int(np.random.choice(np.arange(search_pi.shape[0]), p=search_pi))
    best_child_q, has_next = eng.search_commit(slot, move)
    next_root = None
    env._tree_gen += 1
    if has_next:
        next_root = Node(to_play=env.opponent_player, num_actions=env.action_dim, move=None, parent=None, _env=env, _gen=env._tree_gen)
        next_root.N, next_root.W = float(res['child_N'][move]), float(res['child_W'][move])
    else:
        best_child_q = 0.0
    root_q = res['root_q']
    return move, search_pi, root_q, best_child_q, next_root


def uct_search(env, eval_func, root_node, c_puct_base, c_puct_init, num_simulations=800, root_noise=False, warm_up=False, deterministic=False):
    """mcts_v2.py:301 — same arguments, same 5-tuple (move, search_pi, root_Q, best_child_Q, next_root_node)."""
    return _search(env, eval_func, root_node, c_puct_base, c_puct_init, num_simulations, 1, root_noise, warm_up, deterministic)


def parallel_uct_search(env, eval_func, root_node, c_puct_base, c_puct_init, num_simulations, num_parallel, root_noise=False, warm_up=False,
                        deterministic=False):
    """mcts_v2.py:485 — virtual-loss batched search; same arguments, same 5-tuple."""
    # num_parallel == 1 degenerates to the serial search here (the reference's act() never calls it that way, pipeline.py:132)
    return _search(env, eval_func, root_node, c_puct_base, c_puct_init, num_simulations, int(num_parallel), root_noise, warm_up, deterministic)
