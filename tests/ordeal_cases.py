"""Action scripts for the Ordeal goldens (examples/ordeal.py): shortest walks found
by BFS over the chapter art, so that the sword, both outcomes of the battle with
the dragonduck, every chapter crossing and the quit action are all exercised."""

import collections

import numpy as np

from pycolab_b200.games import ordeal

N, S, W, E, QUIT = 0, 1, 2, 3, 4
_STEP = {N: (-1, 0), S: (1, 0), W: (0, -1), E: (0, 1)}


def _walk(chapter, start, goal):
  """Shortest action list from `start` to `goal` over cells the player may enter."""
  art = ordeal.ARTS[chapter]
  rows, cols = len(art), len(art[0])
  ok = lambda r, c: 0 <= r < rows and 0 <= c < cols and art[r][c] not in '@#w'
  prev = {start: None}
  queue = collections.deque([start])
  while queue:
    cur = queue.popleft()
    if cur == goal:
      break
    for a, (dr, dc) in _STEP.items():
      nxt = (cur[0] + dr, cur[1] + dc)
      if ok(*nxt) and nxt not in prev:
        prev[nxt] = (cur, a)
        queue.append(nxt)
  assert goal in prev, (chapter, start, goal)
  out, cur = [], goal
  while prev[cur] is not None:
    cur, a = prev[cur]
    out.append(a)
  return out[::-1]


def scripts():
  """{name: action list}."""
  start = (7, 12)                                   # 'P' in GAME_ART_KANSAS
  to_cavern = _walk('kansas', start, (4, 44)) + [E]            # off the east edge, row 4
  sword = _walk('cavern', (4, 0), (3, 8))
  back = _walk('cavern', (3, 8), (4, 0)) + [W]                 # off the west edge
  to_castle = _walk('kansas', (4, 44), (0, 7)) + [N]           # off the north edge, col 7
  fight = [N] * 8
  out = {
      'ordeal_sword_wins': to_cavern + sword + back + to_castle + fight,
      # the duck only catches a player who stops: walk into the north wall and wait
      'ordeal_no_sword_loses': _walk('kansas', start, (0, 7)) + [N] + [N] * 8 + [E, E, S],
      'ordeal_castle_and_back': _walk('kansas', start, (0, 6)) + [N, S] + [S, E, W, QUIT],
      'ordeal_quit': [E, E, S, QUIT],
  }
  for seed in range(3):
    rs = np.random.RandomState(40 + seed)
    out['ordeal_random_%d' % seed] = rs.choice([N, S, W, E], size=500,
                                               p=[.3, .2, .2, .3]).tolist()
  return out
