"""BoardGameEnv (reference: alpha_zero/envs/base.py:26) as a handle on a GPU-resident game slot.

Same attributes and methods as the reference class — `reset/step/observation/legal_actions/is_game_over/
to_play/opponent_player/steps/winner/last_player/last_move/history/...` — but `step` runs the rules kernel
(csrc/az_board.cuh) through the C ABI and the arrays are read back from HBM.  Objects can be deep-copied
(new slot + az_env_copy) and pickled (az_env_export); the engine is created lazily per process.
"""
import sys
from collections import namedtuple
from io import StringIO

import numpy as np

from . import _pool
from .coords import CoordsConvertor

PlayerMove = namedtuple('PlayerMove', ['color', 'move'])


class _Space:
    def __init__(self, shape=None, n=None):
        self.shape, self.n = shape, n


class BoardGameEnv:
    game = None  # 'go' | 'gomoku'

    def __init__(self, board_size=15, num_stack=8, black_player_id=1, white_player_id=2, has_pass_move=False, has_resign_move=False,
                 id='', komi=0.0, max_steps=0, num_to_win=5):
        self.id = id
        self.board_size = board_size
        self.num_stack = num_stack
        self.black_player = black_player_id
        self.white_player = white_player_id
        self.has_pass_move = has_pass_move
        self.has_resign_move = has_resign_move
        self.action_dim = board_size**2 + (1 if has_pass_move else 0)
        self.observation_space = _Space(shape=(num_stack * 2 + 1, board_size, board_size))
        self.action_space = _Space(n=self.action_dim)
        self.pass_move = self.action_dim - 1 if has_pass_move else None
        self.resign_move = -1 if has_resign_move else None
        self.gtp_columns = 'ABCDEFGHJKLMNOPQRSTUVWXYZ'
        self.gtp_rows = [str(i) for i in range(board_size, -1, -1)]
        self.cc = CoordsConvertor(board_size)
        self._key = (self.game, board_size, float(komi), int(max_steps), int(num_to_win), int(num_stack))
        self._pool = None
        self._slot = None
        self._pending = None  # exported state waiting for a slot (after unpickling)
        self._tree_gen = 0
        self.history = []
        self._cache = {}

    # ---- slot management --------------------------------------------------------------------------
    @property
    def engine(self):
        self._attach()
        return self._pool.engine

    @property
    def slot(self):
        self._attach()
        return self._slot

    def _attach(self):
        if self._slot is None:
            self._pool = _pool.get_pool(self._key)
            self._slot = self._pool.acquire()
            if self._pending is not None:
                self._pool.engine.env_import(self._slot, self._pending)
                self._pending = None
            else:
                self._pool.engine.env_reset([self._slot])
            self._cache = {}

    def __del__(self):
        try:
            if self._slot is not None and self._pool is not None:
                self._pool.release(self._slot)
        except Exception:
            pass

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ('_slot', '_pool', '_cache', 'history', '_pending')})
        new.history = list(self.history)
        new._slot, new._pool, new._pending, new._cache = None, None, None, {}
        new._tree_gen = 0
        if self._slot is not None:
            new._attach()
            self.engine.env_copy(self._slot, new._slot)
            new._cache = {}
        elif self._pending is not None:
            new._pending = self._pending
        return new

    def __getstate__(self):
        st = {k: v for k, v in self.__dict__.items() if k not in ('_slot', '_pool', '_cache')}
        st['_pending'] = self.engine.env_export(self._slot) if self._slot is not None else self._pending
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._slot, self._pool, self._cache = None, None, {}

    # ---- state read back from the device (cached until the next mutation) --------------------------
    def _scalars(self):
        if 's' not in self._cache:
            self._cache['s'] = self.engine.env_scalars(self.slot)
        return self._cache['s']

    def _mutated(self):
        self._cache = {}

    @property
    def to_play(self):
        return self._scalars()['to_play']

    @property
    def steps(self):
        return self._scalars()['steps']

    @property
    def winner(self):
        s = self._scalars()
        return s['winner'] if (s['done'] and s['winner'] != 0) else None

    @property
    def last_player(self):
        s = self._scalars()
        return s['last_player'] if s['last_move'] != -2 else None

    @property
    def last_move(self):
        m = self._scalars()['last_move']
        return None if m == -2 else m

    @property
    def board(self):
        if 'b' not in self._cache:
            self._cache['b'] = self.engine.env_board(self.slot)
        return self._cache['b']

    @property
    def legal_actions(self):
        if 'l' not in self._cache:
            self._cache['l'] = self._legal_dtype(self.engine.env_legal(self.slot))
        return self._cache['l']

    def _legal_dtype(self, mask):
        return mask.astype(np.int8)

    @property
    def opponent_player(self):
        return self.white_player if self.to_play == self.black_player else self.black_player

    # ---- gym-style API ---------------------------------------------------------------------------
    def reset(self, **kwargs):
        self.engine.env_reset([self.slot])
        self._mutated()
        self._tree_gen += 1
        del self.history[:]
        return self.observation()

    def step(self, action):
        if action is None:
            raise ValueError('Invalid action. The action None is out of bound.')
        mover = self.to_play
        r, d = self.engine.env_step([self.slot], [int(action)])  # raises the reference's errors (go.py:90-95)
        self._mutated()
        if action != self.resign_move:
            self.history.append(PlayerMove(color=self.get_player_name_by_id(mover), move=int(action)))
        reward = float(r[0])
        if self.has_resign_move and action == self.resign_move:
            reward = -1  # the reference returns the int -1 here (go.py:119)
        return self.observation(), reward, bool(d[0]), {}

    def observation(self):
        if 'o' not in self._cache:
            self._cache['o'] = self.engine.env_observation(self.slot)
        return self._cache['o'].copy()

    def is_game_over(self):
        return bool(self._scalars()['done'])

    def close(self):
        del self.history[:]

    # ---- helpers with the reference's names (envs/base.py:268-364) ---------------------------------
    def add_to_history(self, player_id, move):
        if move != self.resign_move:
            self.history.append(PlayerMove(color=self.get_player_name_by_id(player_id), move=move))

    def is_board_full(self):
        return bool(np.all(self.board != 0))

    def is_pass_move(self, move):
        return self.has_pass_move and move == self.pass_move

    def is_resign_move(self, move):
        return self.has_resign_move and move == self.has_resign_move

    def is_legal_move(self, move):
        if move is None or move < 0 or move > self.action_dim - 1:
            return False
        return self.legal_actions[move] == 1

    def is_coords_on_board(self, coords):
        x, y = coords
        return max(x, y) < self.board_size and min(x, y) >= 0

    def action_to_coords(self, action):
        return (-1, -1) if action is None else self.cc.from_flat(action)

    def action_to_gtp(self, action):
        try:
            return self.cc.to_gtp(self.cc.from_flat(action))
        except Exception:
            return None

    def coords_to_action(self, coords):
        try:
            return self.cc.to_flat(coords) if self.is_coords_on_board(coords) else None
        except Exception:
            return None

    def gtp_to_action(self, gtpc, check_illegal=True):
        try:
            action = self.cc.to_flat(self.cc.from_gtp(gtpc))
            if action < 0 or action >= self.action_dim:
                return None
            if check_illegal and self.legal_actions[action] != 1:
                return None
            return action
        except Exception:
            return None

    def get_player_name_by_id(self, id):
        return 'B' if id == self.black_player else ('W' if id == self.white_player else None)

    def get_captures(self):
        return {self.black_player: 0, self.white_player: 0}

    def get_result_string(self):
        return ''

    def to_sgf(self):
        return None

    def render(self, mode='terminal'):
        out = StringIO() if mode == 'ansi' else sys.stdout
        out.write(f'{self.id} ({self.board_size}x{self.board_size})\nBlack: X, White: O\n\n')
        out.write(f'Game over: {"Yes" if self.is_game_over() else "No"}, Result: {self.get_result_string()}\n')
        out.write(f'Steps: {self.steps}, Current player: {"X" if self.to_play == self.black_player else "O"}\n\n')
        cols = '     ' + ''.join(f'{self.gtp_columns[c]:3}' for c in range(self.board_size))
        out.write(cols + '\n   +' + '-' * self.board_size * 3 + '+\n')
        board, last = self.board, self.action_to_coords(self.last_move)
        for r in range(self.board_size):
            out.write(f'{self.gtp_rows[r]:2} |')
            for c in range(self.board_size):
                ch = 'X' if board[r, c] == self.black_player else ('O' if board[r, c] == self.white_player else '.')
                out.write((f'({ch})' if (r, c) == last else ch).center(3))
            out.write(f'| {self.gtp_rows[r]:2}\r\n')
        out.write('   +' + '-' * self.board_size * 3 + '+\n' + cols + '\n\n')
        return out
