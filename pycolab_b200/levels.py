"""Seeded synthetic level generators for the benchmark configurations.

The reference ships only small hand-drawn levels (scrolly_maze 10x30 board over
a 45x89 world, warehouse 11x10..11x13, marauders 16x39; see
examples/scrolly_maze.py:45-195, warehouse_manager.py:42-80,
extraterrestrial_marauders.py:40-55).  BASELINE.json's configs ask for 64x64 and
80x80 boards, so these generators produce ASCII art in the SAME vocabulary as
the reference art (same characters, same invariants) that feeds the reference,
the oracle and the B200 engine alike.  All randomness is
`numpy.random.RandomState(seed)` (stream frozen by NumPy policy).
"""

import numpy as np


def _to_art(arr):
  return [bytes(row).decode('ascii') for row in arr]


def scrolly_maze_level(seed, world_shape=(129, 129), board_shape=(64, 64),
                       coin_density=0.08, star_density=0.08, extra_doors=0.06):
  """A scrolly_maze level: (maze_art, board_art, what_lies_beneath_mark).

  Invariants kept from examples/scrolly_maze.py:45-60: solid outer wall (the
  player cannot escape), exactly one each of 'P', 'a', 'b', 'c' and '+', the
  board window anchored at '+' lies inside the world.  The maze is a spanning
  tree carved over the odd-coordinate lattice with a fraction `extra_doors` of
  the remaining interior walls knocked out so patrollers have room to roam.
  """
  rs = np.random.RandomState(seed)
  WH, WW = world_shape
  BH, BW = board_shape
  assert WH % 2 == 1 and WW % 2 == 1 and WH >= BH and WW >= BW
  art = np.full((WH, WW), ord('#'), dtype=np.uint8)
  nr, nc = WH // 2, WW // 2                 # lattice rooms at (2i+1, 2j+1)
  visited = np.zeros((nr, nc), dtype=bool)
  stack = [(int(rs.randint(nr)), int(rs.randint(nc)))]
  visited[stack[0]] = True
  art[2 * stack[0][0] + 1, 2 * stack[0][1] + 1] = ord(' ')
  steps = ((-1, 0), (1, 0), (0, -1), (0, 1))
  while stack:
    r, c = stack[-1]
    options = [(r + dr, c + dc, dr, dc) for dr, dc in steps
               if 0 <= r + dr < nr and 0 <= c + dc < nc
               and not visited[r + dr, c + dc]]
    if not options:
      stack.pop()
      continue
    r2, c2, dr, dc = options[int(rs.randint(len(options)))]
    visited[r2, c2] = True
    art[2 * r + 1 + dr, 2 * c + 1 + dc] = ord(' ')
    art[2 * r2 + 1, 2 * c2 + 1] = ord(' ')
    stack.append((r2, c2))
  # Extra doors: interior walls between two rooms (one odd, one even coord).
  rr, cc = np.meshgrid(np.arange(1, WH - 1), np.arange(1, WW - 1), indexing='ij')
  door = ((rr % 2) != (cc % 2)) & (art[1:-1, 1:-1] == ord('#'))
  door &= rs.random_sample(door.shape) < extra_doors
  art[1:-1, 1:-1][door] = ord(' ')

  # Board window: '+' sits on an (even, even) lattice point, always a wall.
  cr = 2 * int(rs.randint(0, (WH - BH) // 2 + 1))
  cc0 = 2 * int(rs.randint(0, (WW - BW) // 2 + 1))
  cr, cc0 = min(cr, WH - BH), min(cc0, WW - BW)
  cr -= cr % 2
  cc0 -= cc0 % 2

  floor = np.argwhere(art == ord(' '))
  def pick(pred):
    cand = floor[[bool(pred(r, c)) for r, c in floor]]
    r, c = cand[int(rs.randint(len(cand)))]
    return int(r), int(c)
  taken = set()
  def place(ch, pred):
    while True:
      pos = pick(pred)
      if pos not in taken:
        taken.add(pos)
        return pos
  # Player well inside the initial window; patrollers anywhere in the window's
  # neighbourhood (they may start off-board, as in the reference's Maze #0).
  inner = lambda r, c: (cr + BH // 4 <= r < cr + BH - BH // 4 and
                        cc0 + BW // 4 <= c < cc0 + BW - BW // 4)
  near = lambda r, c: (cr - 8 <= r < cr + BH + 8 and cc0 - 8 <= c < cc0 + BW + 8)
  spots = {'P': place('P', inner)}
  for ch in 'abc':
    spots[ch] = place(ch, near)
  coins = rs.random_sample(len(floor)) < coin_density
  for (r, c), is_coin in zip(floor, coins):
    if is_coin and (int(r), int(c)) not in taken:
      art[r, c] = ord('@')
  for ch, (r, c) in spots.items():
    art[r, c] = ord(ch)
  art[cr, cc0] = ord('+')

  stars = np.full((BH, BW), ord(' '), dtype=np.uint8)
  stars[rs.random_sample((BH, BW)) < star_density] = ord('.')
  return _to_art(art), _to_art(stars), '#'


def warehouse_level(seed, shape=(80, 80), num_boxes=10, num_goals=12,
                    wall_density=0.10):
  """A warehouse_manager level (art only; every sprite stands on ' ').

  Invariants from examples/warehouse_manager.py:42-80,203-226: a '.' border of
  width >= 1 around a '#' wall ring (so the BoxSprite's layers['P'][r±1, c±1]
  look-ups stay in bounds), boxes '0'..'9' each at most once, one 'P', goals
  '_' >= boxes so the puzzle is not trivially unsolvable by count.
  """
  rs = np.random.RandomState(seed)
  H, W = shape
  assert 1 <= num_boxes <= 10 and H >= 8 and W >= 8
  art = np.full((H, W), ord('.'), dtype=np.uint8)
  art[1:H - 1, 1:W - 1] = ord('#')
  art[2:H - 2, 2:W - 2] = ord(' ')
  inner = art[2:H - 2, 2:W - 2]
  inner[rs.random_sample(inner.shape) < wall_density] = ord('#')
  floor = np.argwhere(art == ord(' '))
  order = rs.permutation(len(floor))
  need = num_goals + num_boxes + 1
  assert len(floor) >= need
  picks = floor[order[:need]]
  k = 0
  for _ in range(num_goals):
    art[tuple(picks[k])] = ord('_'); k += 1
  for b in '1234567890'[:num_boxes]:
    art[tuple(picks[k])] = ord(b); k += 1
  art[tuple(picks[k])] = ord('P')
  return _to_art(art)


def marauders_level(rows=16, cols=39):
  """The marauders layout (extraterrestrial_marauders.py:40-55) generated
  procedurally: five staggered rows of 'X' every 4th column, four 4x3 bunkers
  on rows 11-13, the player two cells in on the last row."""
  assert rows >= 16 and cols >= 39
  art = np.full((rows, cols), ord(' '), dtype=np.uint8)
  for r in range(5):
    start = 4 if r % 2 == 0 else 5
    for i in range(8):
      art[r, start + 4 * i] = ord('X')
  for b in range(4):
    art[11:14, 4 + 9 * b: 8 + 9 * b] = ord('B')
  art[rows - 1, 2] = ord('P')
  return _to_art(art)


def classic_level(kind):
  """A second, larger level for each `examples/classics` game (the stock art is
  the only one the reference ships; the rules read only the board shape and,
  for four_rooms, the fixed goal cell (4, 3) — four_rooms.py:78)."""
  if kind == 'four_rooms':
    rows, cols = 15, 21
    a = np.full((rows, cols), ord(' '), dtype=np.uint8)
    a[0, :] = a[-1, :] = a[:, 0] = a[:, -1] = ord('#')
    a[:, 10] = ord('#')
    a[7, :] = ord('#')
    for r, c in ((3, 10), (11, 10), (7, 4), (7, 15)):   # doors
      a[r, c] = ord(' ')
    a[2, 7] = ord('P')
    return _to_art(a)
  if kind == 'cliff_walk':
    a = np.full((6, 20), ord('.'), dtype=np.uint8)
    a[5, 0] = ord('P')
    return _to_art(a)
  if kind == 'chain_walk':
    a = np.full((1, 40), ord('.'), dtype=np.uint8)
    a[0, 17] = ord('P')
    return _to_art(a)
  raise ValueError(kind)


def fluvial_level(rows=7, cols=40, seed=0):
  """A second river for `examples/fluvial_natation.py` (the rules flow rows 1..3
  whatever the size, fluvial_natation.py:106-110)."""
  rs = np.random.RandomState(seed)
  a = np.full((rows, cols), ord(' '), dtype=np.uint8)
  a[0, :] = a[-1, :] = ord('=')
  waves = rs.random_sample((rows - 2, cols)) < 0.15
  a[1:-1][waves] = rs.choice([ord(c) for c in '.,`:~'], size=int(waves.sum()))
  a[3, cols // 3] = ord('P')
  return _to_art(a)


def aperture_level():
  """A small `examples/aperture.py` level: an ooze moat between the player and
  the cranachan, special walls '@' on both sides to shoot apertures into."""
  return ['###############',
          '#@           @#',
          '#  A   ..     #',
          '#      ..     #',
          '#@     ..    @#',
          '#      ..  C  #',
          '###@###..##@###',
          '###############']


def shockwave_level(seed, height=12, width=15, safety_density=0.15):
  """A shockwave level in the vocabulary of examples/shockwave.py:40-88: a '^' safe row
  on top, ' ' exposed cells, '+' bunkers, '=' wall runs on non-adjacent interior rows
  (never the bottom row, always leaving gaps), one 'P' on the bottom row.  Same
  ingredients as the reference's `random_level`, drawn from a private
  RandomState(seed) (that function itself no longer runs: it uses `np.bool`)."""
  rs = np.random.RandomState(seed)
  level = np.full((height, width), ord(' '), dtype=np.uint8)
  level[rs.random_sample(level.shape) < safety_density] = ord('+')
  rows = set(range(1, height - 1))
  while rows:
    row = int(rs.choice(sorted(rows)))
    n_walls = int(rs.randint(2, max(3, width - 3)))
    mask = np.zeros((width,), dtype=bool)
    mask[:n_walls] = True
    rs.shuffle(mask)
    level[row, mask] = ord('=')
    rows -= {row - 1, row, row + 1}
  level[-1, int(rs.randint(0, width - 1))] = ord('P')
  level[0] = ord('^')
  return _to_art(level)
