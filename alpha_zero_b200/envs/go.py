"""GoEnv façade (reference: alpha_zero/envs/go.py:19).  Rules run in csrc/az_board.cuh on the GPU."""
import os

import numpy as np

from ..util import get_time_stamp, make_sgf
from .base import BoardGameEnv

BLACK, WHITE = 1, -1  # go_engine.py:37


def default_board_size():
    """The reference fixes the Go board size per process through $BOARD_SIZE (go_engine.py:31)."""
    return int(os.environ.get('BOARD_SIZE', 19))


class GoEnv(BoardGameEnv):
    game = 'go'
    metadata = {'render.modes': ['terminal'], 'players': ['black', 'white']}

    def __init__(self, komi=7.5, num_stack=8, max_steps=None, board_size=None):
        n = board_size or default_board_size()
        self.komi = komi
        self.max_steps = n * n * 2 if max_steps is None else max_steps
        super().__init__(id='Go', board_size=n, num_stack=num_stack, black_player_id=BLACK, white_player_id=WHITE, has_pass_move=True,
                         has_resign_move=True, komi=komi, max_steps=self.max_steps)

    def _legal_dtype(self, mask):
        # the reference's mask is int64 while the game runs and int8 zeros once it is over (go_engine.py:441, go.py:142)
        return mask.astype(np.int8 if self.is_game_over() else np.int64)

    def get_captures(self):
        s = self._scalars()
        return {self.black_player: s['caps_b'], self.white_player: s['caps_w']}

    def score(self):
        return self.engine.env_score(self.slot)

    def get_result_string(self):
        s = self._scalars()
        if s['by_resign']:
            return 'B+R' if s['winner'] == self.black_player else 'W+R'
        sc = self.score()
        if sc > 0:
            return 'B+' + '%.1f' % sc
        if sc < 0:
            return 'W+' + '%.1f' % abs(sc)
        return 'DRAW'

    def to_sgf(self):
        return make_sgf(self.board_size, self.history, self.get_result_string(), ruleset='Chinese', komi=self.komi, date=get_time_stamp())
