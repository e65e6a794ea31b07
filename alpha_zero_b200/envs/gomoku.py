"""GomokuEnv façade (reference: alpha_zero/envs/gomoku.py:17).  Win scan runs in csrc/az_board.cuh on the GPU."""
from ..util import get_time_stamp, make_sgf
from .base import BoardGameEnv


class GomokuEnv(BoardGameEnv):
    game = 'gomoku'

    def __init__(self, board_size=15, num_to_win=5, num_stack=8):
        self.num_to_win = num_to_win
        super().__init__(id='Freestyle Gomoku', board_size=board_size, num_stack=num_stack, has_pass_move=False, has_resign_move=False,
                         num_to_win=num_to_win)

    def get_result_string(self):
        if not self.is_game_over():
            return ''
        w = self.winner
        return 'B+1.0' if w == self.black_player else ('W+1.0' if w == self.white_player else 'DRAW')

    def to_sgf(self):
        return make_sgf(self.board_size, self.history, self.get_result_string(), ruleset='', komi='', date=get_time_stamp())
