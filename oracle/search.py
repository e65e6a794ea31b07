"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference PUCT search.

Follows /root/reference/alpha_zero/core/mcts_v2.py (file:line cited per function) but stores the tree
structure-of-arrays — one row per node in growing 2-D arrays, statistics of a node kept in its
parent's row — i.e. the same layout the CUDA engine uses in HBM, so this file is also the
executable specification of the device kernels.  Pinned by tests/test_oracle_search.py against
traces recorded from the reference (tests/golden/mcts_*.npz).

Numerics that must not be "fixed" (SURVEY.md section 9):
  * rows are float32; the visit count of the root is a Python float for a fresh root and an
    np.float32 for a re-used one (mcts_v2.py:56-62, 439-443), which changes the rounding of
    (1 + N + c_base) / c_base (mcts_v2.py:101);
  * the root's prior row becomes float64 once Dirichlet noise is mixed in (mcts_v2.py:259-262);
  * virtual loss is +1 on W only, applied and reverted in float32, in leaf order (mcts_v2.py:453-482).
"""
import math

import numpy as np


class Tree:
    """Struct-of-arrays search tree; node 0.. are rows. `root` is the row of the current root."""

    def __init__(self, num_actions, root_to_play):
        self.A = num_actions
        cap = 64
        self.N = np.zeros((cap, num_actions), dtype=np.float32)
        self.W = np.zeros((cap, num_actions), dtype=np.float32)
        self.P = np.zeros((cap, num_actions), dtype=np.float32)
        self.child = -np.ones((cap, num_actions), dtype=np.int32)
        self.parent = -np.ones(cap, dtype=np.int32)
        self.move = -np.ones(cap, dtype=np.int32)
        self.to_play = np.zeros(cap, dtype=np.int32)
        self.expanded = np.zeros(cap, dtype=bool)
        self.vloss = np.zeros(cap, dtype=np.int32)
        self.size = 1
        self.root = 0
        self.to_play[0] = root_to_play
        # the DummyNode slot (mcts_v2.py:56-62): Python floats until a carried np.float32 is stored
        self.root_N = 0.0
        self.root_W = 0.0
        self.root_P64 = None  # float64 prior row of a noised root

    def _grow(self):
        for name in ('N', 'W', 'P', 'child', 'parent', 'move', 'to_play', 'expanded', 'vloss'):
            arr = getattr(self, name)
            fill = -1 if name in ('child', 'parent', 'move') else 0
            extra = np.full_like(arr, fill)
            setattr(self, name, np.concatenate([arr, extra], axis=0))

    def new_child(self, node, action, to_play):
        if self.size == self.N.shape[0]:
            self._grow()
        i = self.size
        self.size += 1
        self.parent[i] = node
        self.move[i] = action
        self.to_play[i] = to_play
        self.child[node, action] = i
        return i

    # stats of a node live in the parent's row (mcts_v2.py:111-127)
    def get_N(self, node):
        return self.root_N if node == self.root else self.N[self.parent[node], self.move[node]]

    def add_N(self, node, d):
        if node == self.root:
            self.root_N = self.root_N + d
        else:
            self.N[self.parent[node], self.move[node]] += d

    def add_W(self, node, d):
        if node == self.root:
            self.root_W = self.root_W + d
        else:
            self.W[self.parent[node], self.move[node]] += d

    def root_Q(self):  # mcts_v2.py:129-135
        return self.root_W / self.root_N if self.root_N > 0 else 0.0


class NodeRef:
    """Opaque handle returned as `next_root_node` (the pipeline only passes it back, pipeline.py:308-321)."""

    def __init__(self, tree):
        self.tree = tree


def _pick(tree, node, legal, c_base, c_init):
    """best_child (mcts_v2.py:142-185) with Node.child_U / child_Q (:99-109) spelled out."""
    n_here = tree.get_N(node)
    pb_c = math.log((1 + n_here + c_base) / c_base) + c_init
    prior = tree.root_P64 if (node == tree.root and tree.root_P64 is not None) else tree.P[node]
    visits = tree.N[node]
    u = pb_c * prior * (math.sqrt(n_here) / (1 + visits))
    q = tree.W[node] / np.where(visits > 0, visits, 1)
    score = np.where(legal == 1, -q + u, -9999)
    return int(np.argmax(score))


def _backup(tree, node, value):
    """mcts_v2.py:213-232."""
    assert isinstance(value, float)
    while node >= 0:
        tree.add_N(node, 1)
        tree.add_W(node, value)
        node = tree.parent[node] if node != tree.root else -1
        value = -1 * value


def _vloss(tree, node, sign):
    """add_virtual_loss / revert_virtual_loss (mcts_v2.py:453-482)."""
    while node >= 0:
        if sign > 0:
            tree.vloss[node] += 1
            tree.add_W(node, +1)
        elif tree.vloss[node] > 0:
            tree.vloss[node] -= 1
            tree.add_W(node, -1)
        node = tree.parent[node] if node != tree.root else -1


def _expand(tree, node, prior):
    """mcts_v2.py:188-210 — all actions, illegal ones included, nothing renormalised."""
    assert not tree.expanded[node]
    assert isinstance(prior, np.ndarray) and prior.ndim == 1 and prior.dtype in (np.float32, np.float64)
    tree.P[node] = prior
    tree.expanded[node] = True


def _descend(tree, env, c_base, c_init):
    """Phase 1 of both searches (mcts_v2.py:380-404 / :577-600) on a private copy of the env."""
    sim = env.copy()
    node = tree.root
    obs, reward, done = None, 0.0, sim.is_game_over()
    while tree.expanded[node]:
        a = _pick(tree, node, sim.legal_actions, c_base, c_init)
        nxt = tree.child[node, a]
        if nxt < 0:
            nxt = tree.new_child(node, a, sim.opponent_player)
        node = nxt
        obs, reward, done, _ = sim.step(a)
        if done:
            break
    return node, obs, reward, done


def search(env, eval_func, root_ref, c_puct_base, c_puct_init, num_simulations, num_parallel=1, root_noise=False,
           warm_up=False, deterministic=False, noise_fn=None, choice_fn=None):
    """uct_search (mcts_v2.py:301-450) when num_parallel <= 1, parallel_uct_search (:485-657) otherwise.

    Returns (move, search_pi, root_Q, best_child_Q, next_root) like the reference.  `noise_fn` /
    `choice_fn` default to the same global numpy RNG calls the reference makes, so a seeded run
    consumes the identical random stream.
    """
    if not 1 <= num_simulations:
        raise ValueError(f'Expect `num_simulations` to a positive integer, got {num_simulations}')
    if env.is_game_over():
        raise RuntimeError('Game is over.')
    noise_fn = noise_fn or np.random.dirichlet
    choice_fn = choice_fn or (lambda n, p: np.random.choice(np.arange(n), p=p))
    A = env.action_dim

    if root_ref is None:  # mcts_v2.py:364-368
        prior, value = eval_func(env.observation(), False)
        tree = Tree(A, env.to_play)
        _expand(tree, 0, prior)
        _backup(tree, 0, value)
    else:
        tree = root_ref.tree
    assert tree.to_play[tree.root] == env.to_play
    root_legal = env.legal_actions

    if root_noise:  # mcts_v2.py:235-262
        alphas = np.ones_like(root_legal) * 0.03
        noise = root_legal * noise_fn(alphas)
        base = tree.root_P64 if tree.root_P64 is not None else tree.P[tree.root]
        tree.root_P64 = base * (1 - 0.25) + noise * 0.25

    if num_parallel <= 1:
        while tree.root_N < num_simulations:  # mcts_v2.py:378
            node, obs, reward, done = _descend(tree, env, c_puct_base, c_puct_init)
            if done:
                _backup(tree, node, -reward)
                continue
            prior, value = eval_func(obs, False)
            _expand(tree, node, prior)
            _backup(tree, node, value)
    else:
        while tree.root_N < num_simulations + num_parallel:  # mcts_v2.py:568
            leaves, tries = [], 0
            while len(leaves) < num_parallel and tries < num_parallel * 2:
                tries += 1
                node, obs, reward, done = _descend(tree, env, c_puct_base, c_puct_init)
                if done:
                    _backup(tree, node, -reward)
                    continue
                _vloss(tree, node, +1)
                leaves.append((node, obs))
            if leaves:
                priors, values = eval_func(np.stack([o for _, o in leaves], axis=0), True)
                for (node, _), prior, value in zip(leaves, priors, values):
                    _vloss(tree, node, -1)
                    if tree.expanded[node]:  # picked twice in one batch (mcts_v2.py:619-622)
                        continue
                    _expand(tree, node, prior)
                    _backup(tree, node, value)

    # search policy (mcts_v2.py:265-298): exponent 1 in warm-up, else min(5, 1/0.1) = 5
    visits = root_legal * tree.N[tree.root]
    visits = np.power(visits, max(1.0, min(5.0, 1.0 / (1.0 if warm_up else 0.1))))
    total = np.sum(visits)
    search_pi = visits / total if total > 0 else visits

    if deterministic:
        move = int(np.argmax(tree.N[tree.root]))
    else:  # mcts_v2.py:433-434 / 640-641
        move = None
        while move is None or (warm_up and env.has_pass_move and move == env.pass_move) or root_legal[move] != 1:
            move = int(choice_fn(search_pi.shape[0], search_pi))

    root_q = tree.root_Q()
    next_ref, best_child_q = None, 0.0
    kid = tree.child[tree.root, move]
    if kid >= 0:  # re-root (mcts_v2.py:436-446): carried N / W are float32 scalars from the parent's row
        n_c, w_c = tree.N[tree.root, move], tree.W[tree.root, move]
        tree.root = int(kid)
        tree.root_N, tree.root_W = n_c, w_c
        tree.root_P64 = None
        next_ref = NodeRef(tree)
        best_child_q = -(w_c / n_c if n_c > 0 else 0.0)
    child_N_row = np.array(tree.N[tree.parent[tree.root]] if kid >= 0 else tree.N[tree.root])
    return move, search_pi, root_q, best_child_q, next_ref, child_N_row
