"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference board engines.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned against the reference itself: tests/golden/*.npz were produced by running /root/reference
(tests/golden/make_golden.py) and tests/test_oracle_boards.py replays them through this file.

This is an independent flat-array formulation (flood fill on a 1-D board), not the reference's
set/dict LibertyTracker, but it must produce identical legal masks, captures, ko points, scores and
observations.  Citations are file:line under /root/reference/alpha_zero.

  Go rules          envs/go.py:88-192, envs/go_engine.py:91-152,386-516
  Gomoku rules      envs/gomoku.py:45-146,233-303
  observation       envs/base.py:228-266
"""
from collections import deque

import numpy as np


def _neighbour_table(n):
    nb = []
    for p in range(n * n):
        r, c = divmod(p, n)
        cur = []
        if r + 1 < n:
            cur.append(p + n)
        if r > 0:
            cur.append(p - n)
        if c + 1 < n:
            cur.append(p + 1)
        if c > 0:
            cur.append(p - 1)
        nb.append(tuple(cur))
    return tuple(nb)


_NB_CACHE = {}


def neighbours(n):
    if n not in _NB_CACHE:
        _NB_CACHE[n] = _neighbour_table(n)
    return _NB_CACHE[n]


class _BoardBase:
    """State common to both games; mirrors the attribute names the search and pipeline read
    (mcts_v2.py:356-404, pipeline.py:300-379)."""

    has_pass_move = False
    has_resign_move = False

    def __init__(self, n, num_stack, black, white):
        self.board_size = n
        self.num_stack = num_stack
        self.black_player = black
        self.white_player = white
        self.cells = n * n
        self.action_dim = self.cells + (1 if self.has_pass_move else 0)
        self.pass_move = self.cells if self.has_pass_move else None
        self.resign_move = -1 if self.has_resign_move else None
        self._reset_common()

    def _reset_common(self):
        self.board = np.zeros(self.cells, dtype=np.int8)
        self.to_play = self.black_player
        self.steps = 0
        self.winner = None
        self.last_player = None
        self.last_move = None
        self.history = []  # non-resign moves, in order (base.py:224-226)
        # newest board first, num_stack deep, zero boards at the start (base.py:261-266)
        self.recent = deque([np.zeros(self.cells, dtype=np.int8) for _ in range(self.num_stack)], maxlen=self.num_stack)

    @property
    def opponent_player(self):
        return self.white_player if self.to_play == self.black_player else self.black_player

    def observation(self):
        """[X_t, Y_t, X_t-1, Y_t-1, ..., C] using the CURRENT mover for every time step (base.py:243-259)."""
        n = self.board_size
        planes = np.zeros((2 * self.num_stack + 1, n, n), dtype=np.int8)
        me, opp = self.to_play, self.opponent_player
        for k, b in enumerate(self.recent):
            b2 = b.reshape(n, n)
            planes[2 * k] = b2 == me
            planes[2 * k + 1] = b2 == opp
        if self.to_play == self.black_player:
            planes[-1] = 1
        return planes

    def get_player_name_by_id(self, pid):
        if pid == self.black_player:
            return 'B'
        if pid == self.white_player:
            return 'W'
        return None

    def _check_action(self, action):
        if self.is_game_over():
            raise RuntimeError('Game is over, call reset before using step method.')
        if action is not None and action != self.resign_move and not 0 <= int(action) <= self.action_dim - 1:
            raise ValueError(f'Invalid action. The action {action} is out of bound.')
        if action is not None and action != self.resign_move and self.legal_actions[int(action)] != 1:
            raise ValueError(f'Illegal action {action}.')

    def _copy_common(self, other):
        other.board = self.board.copy()
        other.to_play = self.to_play
        other.steps = self.steps
        other.winner = self.winner
        other.last_player = self.last_player
        other.last_move = self.last_move
        other.history = list(self.history)
        other.recent = deque((b for b in self.recent), maxlen=self.num_stack)  # boards are never mutated in place
        other.legal_actions = self.legal_actions.copy()


class GoBoard(_BoardBase):
    """Go with simple ko, no suicide, pass, resign, Tromp-Taylor area score (go.py, go_engine.py)."""

    has_pass_move = True
    has_resign_move = True
    BLACK, WHITE = 1, -1  # go_engine.py:37

    def __init__(self, board_size=9, komi=7.5, num_stack=8, max_steps=None):
        self.komi = komi
        self.max_steps = 2 * board_size * board_size if max_steps is None else max_steps
        self.nb = neighbours(board_size)
        super().__init__(board_size, num_stack, self.BLACK, self.WHITE)
        self.reset()

    def reset(self):
        self._reset_common()
        self.ko = None
        self.caps = [0, 0]
        self.legal_actions = self._legal_mask()
        return self.observation()

    def copy(self):
        o = GoBoard.__new__(GoBoard)
        o.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ('board', 'history', 'recent', 'legal_actions', 'caps')})
        self._copy_common(o)
        o.caps = list(self.caps)
        return o

    # -- group machinery -------------------------------------------------------------------------
    def _group(self, start):
        """(stones, liberties) of the chain containing `start` (plays the role of go_engine.py:77-89)."""
        colour = self.board[start]
        stones = {start}
        libs = set()
        stack = [start]
        while stack:
            p = stack.pop()
            for q in self.nb[p]:
                v = self.board[q]
                if v == colour:
                    if q not in stones:
                        stones.add(q)
                        stack.append(q)
                elif v == 0:
                    libs.add(q)
        return stones, libs

    def _would_be_suicide(self, p, colour):
        """go_engine.py:386-402: no empty neighbour, captures nothing, and the friendly neighbours have no
        liberty other than p."""
        for q in self.nb[p]:
            if self.board[q] == 0:
                return False
        for q in self.nb[p]:
            _, libs = self._group(q)
            if self.board[q] == colour:
                if len(libs - {p}) > 0:
                    return False
            elif len(libs) == 1:
                return False
        return True

    def _legal_mask(self):
        """go_engine.py:417-441.  dtype int64 because the reference concatenates an int8 array with the Python list [1]."""
        legal = np.zeros(self.action_dim, dtype=np.int64)
        for p in range(self.cells):
            if self.board[p] != 0 or p == self.ko:
                continue
            if not self._would_be_suicide(p, self.to_play):
                legal[p] = 1
        legal[self.cells] = 1
        return legal

    def _koish_colour(self, p):
        """go_engine.py:91-99: empty point whose neighbours are all one colour."""
        if self.board[p] != 0:
            return None
        cols = {int(self.board[q]) for q in self.nb[p]}
        if len(cols) == 1 and 0 not in cols:
            return cols.pop()
        return None

    # -- stepping --------------------------------------------------------------------------------
    def step(self, action):
        self._check_action(action)
        self.last_move = int(action)
        self.last_player = self.to_play
        self.steps += 1

        if action == self.resign_move:  # go.py:103-119
            self.ko = None
            self.to_play = -self.to_play
            self.recent.appendleft(self.board.copy())
            self.legal_actions = np.zeros(self.action_dim, dtype=np.int8)
            self.winner = self.black_player if self.last_player == self.white_player else self.white_player
            return self.observation(), -1, True, {}

        self.history.append(int(action))
        mover = self.to_play
        if action == self.pass_move:  # go_engine.py:443-449
            self.ko = None
        else:
            p = int(action)
            koish = self._koish_colour(p)
            self.board = self.board.copy()
            self.board[p] = mover
            captured = set()
            for q in self.nb[p]:
                if self.board[q] == -mover and q not in captured:
                    stones, libs = self._group(q)
                    if not libs:
                        captured |= stones
            for q in captured:
                self.board[q] = 0
            # go_engine.py:491-494
            self.ko = next(iter(captured)) if (len(captured) == 1 and koish == -mover) else None
            self.caps[0 if mover == self.BLACK else 1] += len(captured)
        self.to_play = -mover
        self.legal_actions = self._legal_mask()
        self.recent.appendleft(self.board.copy())

        reward = 0.0
        done = self.is_game_over()
        if done:  # go.py:140-156
            self.legal_actions = np.zeros(self.action_dim, dtype=np.int8)
            s = self.score()
            self.winner = self.black_player if s > 0 else (self.white_player if s < 0 else None)
            if self.winner is not None:
                reward = 1.0 if self.last_player == self.winner else -1.0
        return self.observation(), reward, done, {}

    def is_game_over(self):  # go.py:176-192
        if self.last_move == self.resign_move and self.last_move is not None:
            return True
        if self.steps >= self.max_steps:
            return True
        h = self.history
        return len(h) >= 2 and h[-1] == self.pass_move and h[-2] == self.pass_move

    # -- scoring ---------------------------------------------------------------------------------
    def area(self):
        """Tromp-Taylor area (go_engine.py:123-152): empty regions touching exactly one colour go to it."""
        black = int(np.count_nonzero(self.board == self.BLACK))
        white = int(np.count_nonzero(self.board == self.WHITE))
        seen = np.zeros(self.cells, dtype=bool)
        for p in range(self.cells):
            if self.board[p] != 0 or seen[p]:
                continue
            region, stack, touch_b, touch_w = 0, [p], False, False
            seen[p] = True
            while stack:
                x = stack.pop()
                region += 1
                for q in self.nb[x]:
                    v = self.board[q]
                    if v == 0:
                        if not seen[q]:
                            seen[q] = True
                            stack.append(q)
                    elif v == self.BLACK:
                        touch_b = True
                    else:
                        touch_w = True
            if touch_b and not touch_w:
                black += region
            elif touch_w and not touch_b:
                white += region
        return black, white

    def score(self):
        b, w = self.area()
        return b - (w + self.komi)

    def get_result_string(self):  # go.py:194-200, go_engine.py:526-534
        if self.last_move == self.resign_move and self.last_move is not None:
            return 'B+R' if self.winner == self.black_player else 'W+R'
        s = self.score()
        if s > 0:
            return 'B+' + '%.1f' % s
        if s < 0:
            return 'W+' + '%.1f' % abs(s)
        return 'DRAW'


class GomokuBoard(_BoardBase):
    """Freestyle Gomoku (gomoku.py): ids black=1 / white=2, no pass, no resign."""

    def __init__(self, board_size=15, num_to_win=5, num_stack=8):
        self.num_to_win = num_to_win
        super().__init__(board_size, num_stack, 1, 2)
        self.reset()

    def reset(self):
        self._reset_common()
        self.legal_actions = np.ones(self.action_dim, dtype=np.int8)
        return self.observation()

    def copy(self):
        o = GomokuBoard.__new__(GomokuBoard)
        o.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ('board', 'history', 'recent', 'legal_actions')})
        self._copy_common(o)
        return o

    def _run_through(self, p, colour, dr, dc):
        """Stones of `colour` in a straight line through p, both directions (gomoku.py:183-303)."""
        n = self.board_size
        r0, c0 = divmod(p, n)
        total = 1
        for sgn in (1, -1):
            r, c = r0 + sgn * dr, c0 + sgn * dc
            while 0 <= r < n and 0 <= c < n and self.board[r * n + c] == colour:
                total += 1
                r += sgn * dr
                c += sgn * dc
        return total

    def _mover_won(self):  # gomoku.py:85-129
        if self.steps < (self.num_to_win - 1) * 2:
            return False
        for dr, dc in ((0, 1), (1, 0), (1, 1), (-1, 1)):
            if self._run_through(self.last_move, self.to_play, dr, dc) >= self.num_to_win:
                return True
        return False

    def step(self, action):
        self._check_action(action)
        self.last_move = int(action)
        self.last_player = self.to_play
        self.steps += 1
        self.history.append(int(action))
        self.legal_actions[action] = 0
        self.board = self.board.copy()
        self.board[action] = self.to_play
        self.recent.appendleft(self.board.copy())
        reward = 0.0
        if self._mover_won():
            reward = 1.0
            self.winner = self.to_play
        done = self.is_game_over()
        self.to_play = self.opponent_player
        return self.observation(), reward, done, {}

    def is_game_over(self):  # gomoku.py:130-135
        return self.winner is not None or bool(np.all(self.board != 0))

    def get_result_string(self):  # gomoku.py:137-146
        if not self.is_game_over():
            return ''
        if self.winner == self.black_player:
            return 'B+1.0'
        if self.winner == self.white_player:
            return 'W+1.0'
        return 'DRAW'
