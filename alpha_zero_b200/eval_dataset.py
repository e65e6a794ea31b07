"""Evaluation dataset from recorded Go games: core/eval_dataset.py:80-277 (`replay_sgf`,
`build_eval_dataset`) and `eval_on_pro_games` (core/pipeline.py:868-941) on the engine.

The reference replays one SGF at a time through its Python GoEnv.  Here the per-file filters (board size, result, ratings,
duplicates, games per player — all text work, order dependent through MATCHES / GAME_COUNTS like the reference's module globals)
run on the host, then every surviving game is replayed by ONE warp of `k_env_replay` (az_env_replay): the kernel writes the
observation before each move, stops at the first move `step()` would reject, and the host keeps the games that played through.
The third-party `sgf` parser the reference calls (utils/sgf_wrapper.py:105) is replaced by the small reader below (main line =
first variation, properties as lists of strings).
"""
import logging
import os
import re

import numpy as np

from .engine import Engine
from .envs.coords import CoordsConvertor

# module state, as in core/eval_dataset.py:17-27
GAME_COUNTS = {}
MATCHES = []
MISMATCH_GAMES = {'winner_mismatch': 0, 'score_mismatch': 0, 'score_mismatch_le_1': 0, 'score_mismatch_gt_1_le_2': 0,
                  'score_mismatch_gt_2_le_4': 0, 'score_mismatch_gt_4': 0}


def reset_filters():
    GAME_COUNTS.clear()
    MATCHES.clear()
    for k in MISMATCH_GAMES:
        MISMATCH_GAMES[k] = 0


# ---- SGF reader ------------------------------------------------------------------------------------------------------
def parse_sgf_main_line(text):
    """Nodes of the first game's main line (first variation at every branch); each node is {ident: [values]}.
    Raises ValueError on malformed input (the reference drops such files, eval_dataset.py:85-91)."""
    n = len(text)
    i = text.find('(')
    if i < 0:
        raise ValueError('no game tree')
    nodes = []
    # GameTree = '(' Sequence GameTree* ')'.  depth counts open trees; while `side` is set we are inside a variation that is
    # not the first child of its parent and only track brackets and parentheses.
    depth, side, closed = 0, None, set()
    cur = None
    while i < n:
        c = text[i]
        if c == '[' and (side is not None or cur is None):
            i = _skip_value(text, i)
        elif c == '(':
            depth += 1
            if side is None and (depth - 1) in closed:
                side = depth  # the parent tree already had a child: this one is a side variation
            i += 1
        elif c == ')':
            if side == depth:
                side = None
            closed.discard(depth)  # children of the tree being closed are forgotten
            depth -= 1
            closed.add(depth)
            i += 1
            if depth == 0:
                break
        elif side is not None:
            i += 1
        elif c == ';':
            cur = {}
            nodes.append(cur)
            i += 1
        elif c.isalpha():
            j = i
            while j < n and text[j].isalpha():
                j += 1
            ident = ''.join(ch for ch in text[i:j] if ch.isupper())
            i = j
            while i < n and text[i].isspace():
                i += 1
            vals = []
            while i < n and text[i] == '[':
                end = _skip_value(text, i)
                vals.append(_unescape(text[i + 1:end - 1]))
                i = end
                j = i
                while j < n and text[j].isspace():
                    j += 1
                if j < n and text[j] == '[':
                    i = j
            if cur is None or not vals:
                raise ValueError('property outside a node')
            cur.setdefault(ident, []).extend(vals)
        else:
            i += 1
    if depth != 0 or not nodes:
        raise ValueError('unbalanced game tree')
    return nodes


def _skip_value(text, i):
    """`i` at '[': index just past the matching ']' (backslash escapes the next character)."""
    n = len(text)
    i += 1
    while i < n:
        if text[i] == '\\':
            i += 2
        elif text[i] == ']':
            return i + 1
        else:
            i += 1
    raise ValueError('unterminated property value')


def _unescape(s):
    return re.sub(r'\\(.)', r'\1', s, flags=re.S)


def _prop(props, key, default=''):
    v = props.get(key, default)
    if v is None or isinstance(v, str):
        return v
    return v[0] if len(v) == 1 else v  # sgf_wrapper.sgf_prop (utils/sgf_wrapper.py:93-100)


def _get_player_str(player):  # eval_dataset.py:36-39
    player = re.sub(r'\([^)]*\)', '', player)
    player = re.sub(r'[^a-zA-Z0-9 ]', '', player)
    return player.strip()


def _extract_ratings(black_player, white_player, black_rank, white_rank):  # eval_dataset.py:56-77
    ratings = []
    if all(r is not None and r != '' and 'k' not in r and 'd' not in r and 'p' not in r for r in (black_rank, white_rank)):
        for r in (black_rank, white_rank):
            try:
                ratings.append(int(re.sub(r'[^0-9]', '', r)))
            except Exception:
                pass
    elif all('(' in p and ')' in p for p in (black_player, white_player)):
        for p in (black_player, white_player):
            m = re.search(r'\((\d+)\)', p)
            if m:
                ratings.append(int(m.group(1)))
    return ratings


def get_sgf_files(games_dir):  # eval_dataset.py:42-53
    out = []
    if os.path.exists(games_dir):
        for root, _dirs, files in os.walk(games_dir):
            out.extend(os.path.join(root, f) for f in files if f.endswith('.sgf'))
    return out


class _Candidate:
    __slots__ = ('name', 'moves', 'players', 'winner', 'komi', 'num_moves', 'result_str')


def _screen(name, text, board_size, logger, min_elo, max_games_per_player):
    """Everything replay_sgf decides before touching the board (eval_dataset.py:85-160), in its order.  None = dropped."""
    try:
        nodes = parse_sgf_main_line(text)
    except Exception:
        return None
    props = nodes[0]
    sz = _prop(props, 'SZ')
    try:
        if sz is None or sz == '' or int(sz) != board_size:
            logger.debug(f'Game "{name}" board size mismatch')
            return None
    except (TypeError, ValueError):
        return None
    result_str = _prop(props, 'RE')
    if result_str is None or result_str == '' or len(result_str) < 3:
        logger.debug(f'Game "{name}" has no result property')
        return None
    if re.search(r'\+T', result_str):
        logger.debug(f'Game "{name}" with result {result_str} does not have a natural winner')
        return None
    black_player, white_player = _prop(props, 'PB'), _prop(props, 'PW')
    ratings = _extract_ratings(black_player, white_player, _prop(props, 'BR'), _prop(props, 'WR'))
    if ratings and any(v < min_elo for v in ratings):
        logger.info(f'Game "{name}" with player ratings {ratings} is too weak')
        return None
    black_id, white_id = _get_player_str(black_player), _get_player_str(white_player)
    num_moves = len(re.findall(r';[BW]\[[a-z]{0,2}\]', text))
    match_str = f'{black_id}-{white_id}-{num_moves}-{result_str}'
    if match_str in MATCHES:
        logger.info(f'Game "{name}" might be duplicate')
        return None
    MATCHES.append(match_str)
    for pid in (black_id, white_id):
        if pid in GAME_COUNTS:
            if GAME_COUNTS[pid] > max_games_per_player:
                logger.info(f'Too many games from player {pid}')
                return None
            GAME_COUNTS[pid] += 1
        else:
            GAME_COUNTS[pid] = 1
    c = _Candidate()
    c.name, c.num_moves, c.result_str = name, num_moves, result_str
    c.komi = float(_prop(props, 'KM')) if props.get('KM') is not None else 0
    c.winner = 1 if re.match(r'B\+', result_str, re.IGNORECASE) else (-1 if re.match(r'W\+', result_str, re.IGNORECASE) else None)
    cc = CoordsConvertor(board_size)
    c.moves, c.players = [], []
    for node in nodes[1:]:
        if 'TW' in node or 'TB' in node:  # territory markup ends the record (eval_dataset.py:166)
            break
        if 'W' in node:
            player, val = -1, node['W'][0]
        elif 'B' in node:
            player, val = 1, node['B'][0]
        else:
            return None
        try:
            flat = cc.to_flat(cc.from_sgf(val))
        except (ValueError, IndexError):
            return None
        if flat < 0 or flat > board_size * board_size:
            return None
        if player != (1 if len(c.moves) % 2 == 0 else -1):  # env.to_play != next_player: handicap games
            return None
        c.moves.append(flat)
        c.players.append(player)
    return c


def _count_mismatch(c, env_result, logger):  # eval_dataset.py:217-245
    env_result, result = env_result.upper(), c.result_str.upper()
    if re.search(r'\+T', result, re.IGNORECASE) or re.search(r'\+R', result, re.IGNORECASE):
        return
    mismatch = False
    if env_result[:2] != result[:2]:
        mismatch = True
        MISMATCH_GAMES['winner_mismatch'] += 1
    else:
        sgf_score = re.findall(r'[-+]?\d*\.\d+|\d+', result)
        env_score = re.findall(r'[-+]?\d*\.\d+|\d+', env_result)
        sgf_score = float(sgf_score[0]) if sgf_score else sgf_score
        env_score = float(env_score[0]) if env_score else env_score
        if sgf_score != env_score:
            mismatch = True
            MISMATCH_GAMES['score_mismatch'] += 1
            delta = abs(sgf_score - env_score)
            key = ('score_mismatch_le_1' if delta <= 1 else 'score_mismatch_gt_1_le_2' if delta <= 2 else
                   'score_mismatch_gt_2_le_4' if delta <= 4 else 'score_mismatch_gt_4')
            MISMATCH_GAMES[key] += 1
    if mismatch:
        logger.debug(f'Game "{c.name}" has mismatching result, env result: {env_result}, SGF result: {result}')


def replay_sgf_games(games, num_stack, board_size=None, logger=None, skip_n=0, min_elo=2100, max_games_per_player=200, device=None,
                     binding=None, max_slots=2048, max_positions=131072):
    """`games`: iterable of (name, sgf_text).  Returns, in order, one entry per game: None (dropped, as replay_sgf returns None) or
    (states int8 [n, 2*num_stack+1, N, N], moves int64 [n], values float32 [n]) — the history replay_sgf builds, with the
    target move as an index instead of a one-hot row."""
    logger = logger or logging.getLogger('alpha_zero_b200.eval_dataset')
    board_size = int(board_size if board_size is not None else os.environ.get('BOARD_SIZE', 19))
    out, cands = [], []
    for name, text in games:
        c = _screen(name, text, board_size, logger, min_elo, max_games_per_player) if text is not None else None
        out.append(None)
        if c is not None:
            cands.append((len(out) - 1, c))
    if not cands:
        return out
    eng = Engine('go', board_size, num_games=min(max_slots, len(cands)), max_simulations=1, max_parallel=1, komi=0.0, num_stack=num_stack,
                 net=None, device=0 if device is None else device, binding=binding)
    try:
        i = 0
        while i < len(cands):
            wave, pos = [], 0
            while i < len(cands) and len(wave) < eng.G and (not wave or pos + len(cands[i][1].moves) <= max_positions):
                wave.append(cands[i])
                pos += len(cands[i][1].moves)
                i += 1
            slots = list(range(len(wave)))
            states, offsets, played, status = eng.env_replay(slots, [c.moves for _, c in wave])
            for s, (idx, c) in enumerate(wave):
                n = len(c.moves)
                if status[s] != 0 or played[s] != n or n != c.num_moves:  # illegal move / game already over / steps != num_moves
                    continue
                score = eng.env_score(s) - c.komi  # the engine was built with komi 0 (GoEnv(komi=KM), eval_dataset.py:151-155)
                # get_result_string() scores the final position whether or not the game ended (envs/go.py:194-199)
                env_result = 'B+%.1f' % score if score > 0 else ('W+%.1f' % abs(score) if score < 0 else 'DRAW')
                _count_mismatch(c, env_result, logger)
                first = int(offsets[s])
                keep = [t for t in range(n) if t > skip_n]  # `if env.steps > skip_n` (eval_dataset.py:197)
                values = np.array([0.0 if c.winner is None else (1.0 if c.winner == c.players[t] else -1.0) for t in keep], dtype=np.float32)
                out[idx] = (states[first + np.array(keep, dtype=np.int64)] if keep else states[first:first],
                            np.array([c.moves[t] for t in keep], dtype=np.int64), values)
    finally:
        eng.close()
    return out


def build_eval_dataset(games_dir, num_stack, logger=None, board_size=None, device=None, binding=None):
    """core/eval_dataset.py:248-277 — TensorDataset(states f32 [n,17,N,N], target_pi one-hot f32 [n,A], target_v f32 [n])."""
    import torch
    from torch.utils.data import TensorDataset

    logger = logger or logging.getLogger('alpha_zero_b200.eval_dataset')
    logger.info('Building evaluation dataset...')

    def texts():
        for path in get_sgf_files(games_dir):
            try:
                with open(path) as f:
                    yield path, f.read()
            except Exception:
                yield path, None

    board_size = int(board_size if board_size is not None else os.environ.get('BOARD_SIZE', 19))
    hist = [h for h in replay_sgf_games(texts(), num_stack, board_size, logger, device=device, binding=binding) if h is not None]
    states = np.concatenate([h[0] for h in hist], axis=0)
    moves = np.concatenate([h[1] for h in hist], axis=0)
    values = np.concatenate([h[2] for h in hist], axis=0)
    target_pi = np.zeros((len(moves), board_size * board_size + 1), dtype=np.float32)
    target_pi[np.arange(len(moves)), moves] = 1.0
    ds = TensorDataset(torch.from_numpy(states).to(dtype=torch.float32), torch.from_numpy(target_pi), torch.from_numpy(values))
    logger.warning(f'Number of games with mismatched results: {MISMATCH_GAMES}')
    logger.info(f'Finished loading {len(ds)} positions from {len(hist)} games')
    return ds


def eval_on_pro_games(network, device, dataloader, k_list=(1, 3, 5), precision=None):
    """core/pipeline.py:868-941 with the forward pass on the engine's tower: top-k accuracy of the policy on the recorded
    move, policy entropy, value MSE.  `dataloader` yields (states, target_pi, target_v) batches (a DataLoader over
    build_eval_dataset's TensorDataset)."""
    import torch
    from .pipeline import _NetOnEngine

    assert min(k_list) >= 1
    if dataloader is None:
        return {}
    holder, correct, entropy, mse, total = None, {k: 0 for k in k_list}, 0.0, 0.0, 0
    for states, target_pi, target_v in dataloader:
        st = states.cpu().numpy().astype(np.int8)
        if holder is None:
            holder = _NetOnEngine(network, device, 'go', st.shape[-1], precision=precision)
        pri, val = holder.net_forward(st)
        p = torch.from_numpy(np.ascontiguousarray(pri))
        _, pred = torch.topk(p, max(k_list), dim=1)
        hit = pred.eq(torch.argmax(target_pi.cpu(), dim=1).unsqueeze(1))
        for k in k_list:
            correct[k] += int(hit[:, :k].any(dim=1).sum())
        entropy += float(-(p * torch.log(p)).sum(dim=1).sum())
        mse += float(((torch.from_numpy(np.ascontiguousarray(val)) - target_v.cpu()) ** 2).sum())
        total += len(st)
    stats = {'value_mse_error': mse / total, 'policy_entropy': entropy / total}
    for k in k_list:
        stats[f'policy_top_{k}_accuracy'] = correct[k] / total
    return stats
