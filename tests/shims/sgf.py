"""Stub of the `sgf` package: the hot path never parses SGF through it."""


def parse(_text):
    raise RuntimeError('sgf stub: parsing is not available in this image')
