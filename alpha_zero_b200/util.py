"""Host-side glue with the reference's names: time stamps, logger, Timer, CsvWriter, SGF writer
(reference: alpha_zero/utils/util.py:15,57,75; utils/csv_writer.py:13; utils/sgf_wrapper.py:38-96)."""
import csv
import datetime
import logging
import os
import sys
import time
from collections import deque

from .envs.coords import CoordsConvertor


def get_time_stamp(file_name=False):
    now = datetime.datetime.now()
    return now.strftime('%Y%m%d_%H%M%S') if file_name else now.strftime('%Y-%m-%d %H:%M:%S')


def create_logger(level='INFO'):
    handler = logging.StreamHandler(stream=sys.stderr)
    handler.setFormatter(logging.Formatter('%(levelname)s %(asctime)s.%(msecs)03d %(process)d %(message)s', '%H:%M:%S'))
    logger = logging.getLogger(f'alpha_zero_b200.{os.getpid()}')
    logger.setLevel(str(level).upper())
    if not logger.handlers:
        logger.addHandler(handler)
    logger.propagate = False
    return logger


class Timer:
    def __init__(self, max_history=100):
        self.history = deque(maxlen=max_history)
        self.t0 = None

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.history.append(time.time() - self.t0)

    def mean_time(self):
        return sum(self.history) / len(self.history) if self.history else 0.0

    def last_time(self):
        return self.history[-1] if self.history else 0.0


class CsvWriter:
    """Append-mode CSV log with a header written once (columns = keys of the first row)."""

    def __init__(self, fname, buffer_size=100, flush_interval=60):
        d = os.path.dirname(fname)
        if d:
            os.makedirs(d, exist_ok=True)
        self.fname, self.buffer_size, self.flush_interval = fname, buffer_size, flush_interval
        self.rows, self.last = [], time.time()
        self.need_header = not os.path.exists(fname) or os.path.getsize(fname) == 0

    def write(self, values):
        self.rows.append(dict(values))
        if len(self.rows) >= self.buffer_size or time.time() - self.last > self.flush_interval:
            self._flush()

    def _flush(self):
        if not self.rows:
            return
        with open(self.fname, 'a', newline='') as f:
            w = csv.DictWriter(f, fieldnames=list(self.rows[0].keys()), extrasaction='ignore')
            if self.need_header:
                w.writeheader()
                self.need_header = False
            w.writerows(self.rows)
        self.rows, self.last = [], time.time()

    def close(self):
        self._flush()


SGF_HEAD = '(;\nCA[UTF-8]\nAP[AlphaZeroMini_sgfgenerator]\nRU[{ruleset}]\nPB[{black}]\nBR[]\nPW[{white}]\nWR[]\nKM[{komi}]\nRE[{result}]\nDT[{date}]\nSZ[{size}]\n\n'


def make_sgf(board_size, move_history, result_string, ruleset='Chinese', komi=7.5, white_name='AlphaZeroMini', black_name='AlphaZeroMini', date=''):
    """Same record format the reference writes (utils/sgf_wrapper.py:38-96): ten moves per line."""
    cc = CoordsConvertor(board_size)
    parts = []
    for i, pm in enumerate(move_history):
        color, move = (pm.color, pm.move) if hasattr(pm, 'color') else pm
        parts.append(f';{color}[{cc.to_sgf(cc.from_flat(move))}]' + ('\n' if (i + 1) % 10 == 0 else ''))
    return SGF_HEAD.format(ruleset=ruleset, black=black_name, white=white_name, komi=komi, result=result_string, date=date, size=board_size) + ''.join(parts) + ')'
