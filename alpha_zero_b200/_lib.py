"""ctypes binding of libaz_b200.so (include/az_engine.h).

The product path loads exactly one library — alpha_zero_b200/libaz_b200.so, the CUDA build — and
raises when it is missing or when no CUDA device can be opened: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libaz_b200.so')

GAME_GO, GAME_GOMOKU = 0, 1
NET_FP32, NET_BF16, NET_BF16X3 = 0, 1, 2
PRECISIONS = {'fp32': NET_FP32, 'bf16': NET_BF16, 'bf16x3': NET_BF16X3}
ERR_INVALID_ACTION, ERR_ILLEGAL_ACTION, ERR_GAME_OVER, ERR_BAD_ARG = -2, -3, -4, -5


class AzConfig(C.Structure):
    _fields_ = [
        ('game', C.c_int32), ('board_size', C.c_int32), ('num_stack', C.c_int32), ('komi', C.c_float),
        ('max_steps', C.c_int32), ('num_to_win', C.c_int32), ('num_games', C.c_int32), ('max_simulations', C.c_int32),
        ('max_parallel', C.c_int32), ('num_res_blocks', C.c_int32), ('num_filters', C.c_int32), ('num_fc_units', C.c_int32),
        ('net_precision', C.c_int32), ('device', C.c_int32), ('seed', C.c_uint64), ('sample_ring', C.c_int32),
        ('reserved', C.c_int32),
    ]


class AzSearchParams(C.Structure):
    _fields_ = [('c_puct_base', C.c_double), ('c_puct_init', C.c_double), ('num_simulations', C.c_int32),
                ('num_parallel', C.c_int32), ('root_noise', C.c_int32), ('deterministic', C.c_int32)]


class AzSelfplayParams(C.Structure):
    _fields_ = [('search', AzSearchParams), ('warm_up_steps', C.c_int32), ('check_resign_after_steps', C.c_int32),
                ('resign_threshold', C.c_float), ('disable_resign_ratio', C.c_float)]


class AzCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ('simulations', 'evaluations', 'moves', 'games', 'nodes', 'depth_sum', 'descents',
                                           'samples', 'ring_dropped', 'errors', 'kernel_launches', 'ticks')]


class AzGameRecord(C.Structure):
    _fields_ = [('slot', C.c_int32), ('game_length', C.c_int32), ('winner', C.c_int32), ('by_resign', C.c_int32),
                ('score', C.c_float), ('num_passes', C.c_int32), ('is_resign_disabled', C.c_int32),
                ('is_marked_for_resign', C.c_int32), ('is_could_won', C.c_int32), ('marked_resign_player', C.c_int32),
                ('first_sample', C.c_int32), ('reserved', C.c_int32)]


# every symbol include/az_engine.h declares; tests/test_abi.py checks the shared library exports each one
SYMBOLS = [
    'az_last_error', 'az_version', 'az_create', 'az_destroy', 'az_get_config', 'az_num_actions', 'az_obs_bytes',
    'az_set_weights', 'az_set_weights_for', 'az_match_begin', 'az_match_tick', 'az_net_forward', 'az_net_conv_layer', 'az_net_info', 'az_env_reset', 'az_env_step', 'az_env_observation', 'az_env_legal_actions',
    'az_env_board', 'az_env_scalars', 'az_env_score', 'az_env_copy', 'az_env_state_bytes', 'az_env_export',
    'az_env_import', 'az_env_replay', 'az_search_begin', 'az_search_select', 'az_search_apply', 'az_search_result', 'az_search_commit',
    'az_search_run', 'az_selfplay_begin', 'az_selfplay_tick', 'az_selfplay_update', 'az_selfplay_restart', 'az_sync', 'az_get_counters', 'az_drain_games',
    'az_gather_pack', 'az_host_alloc', 'az_host_free', 'az_stream', 'az_last_net_ms', 'az_tick_profile', 'az_replay_create', 'az_replay_ingest', 'az_replay_add', 'az_replay_info',
    'az_replay_sample',
]


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class Binding:
    """Typed access to one loaded shared library."""

    def __init__(self, cdll):
        self.dll = cdll
        missing = [s for s in SYMBOLS if not hasattr(cdll, s)]
        if missing:
            raise ImportError(f'shared library lacks symbols: {missing}')
        cdll.az_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            if s != 'az_last_error':
                getattr(cdll, s).restype = C.c_int

    def check(self, rc):
        if rc < 0:
            msg = self.dll.az_last_error().decode('utf-8', 'replace')
            if rc in (ERR_INVALID_ACTION, ERR_ILLEGAL_ACTION, ERR_BAD_ARG):
                raise ValueError(msg)  # the reference raises ValueError for these (envs/go.py:92-95, mcts_v2.py:356-359)
            if rc == ERR_GAME_OVER:
                raise RuntimeError(msg)  # envs/go.py:90, mcts_v2.py:360
            raise EngineError(rc, msg)
        return rc


_BINDING = None


def load():
    """The CUDA library, or an error.  Never falls back to anything else."""
    global _BINDING
    if _BINDING is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a).  alpha_zero_b200 has no CPU fallback.')
        _BINDING = Binding(C.CDLL(LIB_PATH))
    return _BINDING


def as_ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
