"""Coordinate conversions with the reference's conventions (alpha_zero/envs/coords.py:49-91):
flat index = row * N + col with N*N meaning pass, (row, col) from the upper-left corner, SGF 'cr'
letter pairs, GTP 'A19'-style names that skip the letter I."""

SGF_LETTERS = 'abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ'
GTP_LETTERS = 'ABCDEFGHJKLMNOPQRSTUVWXYZ'


class CoordsConvertor:
    def __init__(self, board_size):
        self.board_size = board_size

    # flat <-> (row, col)
    def from_flat(self, flat):
        n = self.board_size
        return None if flat == n * n else divmod(flat, n)

    def to_flat(self, coord):
        n = self.board_size
        return n * n if coord is None else n * coord[0] + coord[1]

    # SGF
    def from_sgf(self, text):
        if not text or (self.board_size <= 19 and text == 'tt'):
            return None
        return SGF_LETTERS.index(text[1]), SGF_LETTERS.index(text[0])

    def to_sgf(self, coord):
        return '' if coord is None else SGF_LETTERS[coord[1]] + SGF_LETTERS[coord[0]]

    # GTP
    def from_gtp(self, text):
        text = text.upper()
        if text == 'PASS':
            return None
        return self.board_size - int(text[1:]), GTP_LETTERS.index(text[0])

    def to_gtp(self, coord):
        if coord is None:
            return 'pass'
        row, col = coord
        return f'{GTP_LETTERS[col]}{self.board_size - row}'
