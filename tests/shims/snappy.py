"""Stub of python-snappy: replay compression is off by default in every reference driver."""


def compress(_b):
    raise RuntimeError('snappy stub')


def uncompress(_b):
    raise RuntimeError('snappy stub')
