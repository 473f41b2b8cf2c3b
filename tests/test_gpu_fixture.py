"""GPU parity of the general device program (PCL_PROG_FIXTURE, csrc/fixture.cu).

Replays the golden trajectories the reference produced with its own test
fixtures (TestMazeWalker / TestScrolly / TestDrape + post_update directive
injection): diagonal moves, EDGE, confined walkers, arbitrary impassable sets,
two egocentric walkers, margin-less scrolling, two Scrollys in one group,
update-group staging, rewards / termination / change_z_order.  Bit-exact.
"""

import numpy as np
import pytest

import golden_cases as gc
from oracle import games as ogames

pytestmark = pytest.mark.gpu


def _torch():
  import torch
  return torch


def _engine(kw, batch):
  from pycolab_b200 import batched
  from pycolab_b200.games import fixtures
  game = fixtures.make_game(kw['art'], kw['what_lies_beneath'], kw['walkers'],
                            kw['scrollys'], kw['drapes'], kw['update_schedule'],
                            kw['z_order'])
  return batched.BatchedEngine([game], batch=batch, auto_reset=False)


def _rows(eng, motions_by_char, T):
  """[T, B, A] action rows from {char: [T] codes} (same for every env)."""
  order = ''.join(eng.game.groups)
  n = len(order)
  from pycolab_b200 import _lib
  rows = np.zeros((T, eng.batch, n + 2 * _lib.FIXTURE_DIRECTIVES), dtype=np.int64)
  for k, ch in enumerate(order):
    rows[:, :, k] = np.asarray(motions_by_char.get(ch, np.full(T, 8)))[:, None]
  return rows.astype(np.int32)            # no Plot directives: opcode 0


def _sprite_rows(eng, chars, env=0):
  rec = eng.sprites[env].cpu().numpy()
  out = []
  for ch in chars:
    r = rec[eng.sprite_chars.index(ch)]
    out.append([r[0], r[1], r[4] & 1, r[2], r[3]])
  return np.array(out)


@pytest.mark.parametrize('name', gc.names('fixture_walkers_'))
def test_fixture_walkers_golden(name):
  torch = _torch()
  g = gc.load(name)
  kw, cfg = gc.fixture_kwargs(g)
  eng = _engine(kw, batch=3)
  chars = cfg['action_chars']
  T = len(g['actions'])
  rows = _rows(eng, {ch: g['actions'][:, i] for i, ch in enumerate(chars)}, T)
  res = eng.its_showtime()
  for t in range(T + 1):
    torch.cuda.synchronize()
    for e in range(eng.batch):
      np.testing.assert_array_equal(res.board[e].cpu().numpy(), g['boards'][t],
                                    err_msg='%s t=%d' % (name, t))
    np.testing.assert_array_equal(_sprite_rows(eng, chars), g['sprites'][t])
    assert int(res.has_reward[0]) == 0 and float(res.discount[0]) == 1.0
    if t < T:
      res = eng.play(torch.from_numpy(rows[t]).cuda())
  assert int(eng.error_codes().abs().max()) == 0


@pytest.mark.parametrize('name', gc.names('fixture_scrolly_'))
def test_fixture_scrolly_golden(name):
  torch = _torch()
  g = gc.load(name)
  kw, cfg = gc.fixture_kwargs(g)
  eng = _engine(kw, batch=2)
  T = len(g['actions'])
  order = ''.join(eng.game.groups)
  rows = _rows(eng, {ch: g['actions'] for ch in order}, T)   # everybody: same motion
  res = eng.its_showtime()
  for t in range(T + 1):
    torch.cuda.synchronize()
    np.testing.assert_array_equal(res.board[0].cpu().numpy(), g['boards'][t],
                                  err_msg='%s t=%d' % (name, t))
    np.testing.assert_array_equal(res.board[1].cpu().numpy(), g['boards'][t])
    np.testing.assert_array_equal(_sprite_rows(eng, 'Pq'), g['sprites'][t])
    cur = np.stack([eng.curtain('#')[0].cpu().numpy(), eng.curtain('@')[0].cpu().numpy()])
    np.testing.assert_array_equal(cur, g['curtains'][t].astype(bool))
    if t < T:
      res = eng.play(torch.from_numpy(rows[t]).cuda())
  assert int(eng.error_codes().abs().max()) == 0


@pytest.mark.parametrize('name', gc.names('fixture_groups_'))
def test_fixture_two_scrolling_groups_golden(name):
  """Named scrolling groups on the device (protocols/scrolling.py:198-241,
  `scrolling_group` of sprites.py MazeWalker / drapes.py Scrolly): group 'one' =
  '#' + P, group 'two' = '@' + q, each driven by its own motion; boards, sprite
  registers and both curtains must equal the reference's at every step."""
  torch = _torch()
  g = gc.load(name)
  kw, cfg = gc.fixture_kwargs(g)
  eng = _engine(kw, batch=2)
  assert eng.game.scroll_groups == ['one', 'two'] and eng.groups is not None
  T = len(g['actions'])
  rows = _rows(eng, {ch: g['actions'][:, k] for ch, k in cfg['motion_of'].items()}, T)
  res = eng.its_showtime()
  for t in range(T + 1):
    torch.cuda.synchronize()
    for e in range(2):
      np.testing.assert_array_equal(res.board[e].cpu().numpy(), g['boards'][t],
                                    err_msg='%s t=%d' % (name, t))
    np.testing.assert_array_equal(_sprite_rows(eng, 'Pq'), g['sprites'][t])
    cur = np.stack([eng.curtain('#')[0].cpu().numpy(), eng.curtain('@')[0].cpu().numpy()])
    np.testing.assert_array_equal(cur, g['curtains'][t].astype(bool))
    if t < T:
      res = eng.play(torch.from_numpy(rows[t]).cuda())
  assert int(eng.error_codes().abs().max()) == 0


@pytest.mark.parametrize('name', gc.names('fixture_directives_'))
def test_fixture_directives_golden(name):
  torch = _torch()
  g = gc.load(name)
  kw, cfg = gc.fixture_kwargs(g)
  eng = _engine(kw, batch=2)
  assert ''.join(eng.game.groups) == cfg['action_chars']
  T = len(g['actions'])
  res = eng.its_showtime()
  for t in range(T + 1):
    torch.cuda.synchronize()
    np.testing.assert_array_equal(res.board[1].cpu().numpy(), g['boards'][t],
                                  err_msg='%s t=%d' % (name, t))
    assert (int(res.has_reward[0]), int(res.reward[0])) == (
        int(g['has_reward'][t]), int(g['reward'][t])), t
    assert float(res.discount[0]) == float(g['discount'][t])
    assert int(res.done[0]) == int(g['game_over'][t])
    np.testing.assert_array_equal(eng.z_order[0].cpu().numpy(), g['z_orders'][t])
    if t < T:
      row = np.array(gc.new_directive_row(g['actions'][t], len(cfg['action_chars'])),
                     dtype=np.int32)
      res = eng.play(torch.from_numpy(np.stack([row, row])).cuda())
  assert int(eng.error_codes().abs().max()) == 0


@pytest.mark.parametrize('name', gc.names('fixture_unoccluded_'))
def test_unoccluded_layers_golden(name):
  """occlusion_in_layers=False through the facade (rendering.py:187-301)."""
  from pycolab_b200.games import fixtures
  g = gc.load(name)
  kw, cfg = gc.fixture_kwargs(g)
  game = fixtures.make_game(kw['art'], kw['what_lies_beneath'], kw['walkers'],
                            kw['scrollys'], kw['drapes'], kw['update_schedule'],
                            kw['z_order'], occlusion_in_layers=False)
  names = ('n', 'ne', 'e', 'se', 's', 'sw', 'w', 'nw', 'stay')
  obs, _, _ = game.its_showtime()
  chars = cfg['layer_chars']
  for t in range(len(g['actions']) + 1):
    np.testing.assert_array_equal(obs.board, g['boards'][t], err_msg='t=%d' % t)
    assert ''.join(sorted(obs.layers)) == chars
    for i, ch in enumerate(chars):
      np.testing.assert_array_equal(obs.layers[ch], g['layers'][t][i].astype(bool),
                                    err_msg='layer %r t=%d' % (ch, t))
    if t < len(g['actions']):
      obs, _, _ = game.play(names[int(g['actions'][t])])


def test_unoccluded_renderer_class():
  from pycolab_b200 import rendering
  r = rendering.BaseUnoccludedObservationRenderer(3, 4, 'ab. ')
  r.clear()
  bd = np.full((3, 4), ord('.'), np.uint8)
  bd[0, 0] = ord(' ')
  r.paint_all_of(bd)
  mask = np.zeros((3, 4), bool)
  mask[1, :] = True
  r.paint_drape('a', mask)
  r.paint_sprite('b', (1, 2))
  obs = r.render()
  want = bd.copy()
  want[1, :] = ord('a')
  want[1, 2] = ord('b')
  np.testing.assert_array_equal(obs.board, want)
  np.testing.assert_array_equal(obs.layers['a'], mask)          # not occluded by b
  np.testing.assert_array_equal(obs.layers['.'], bd == ord('.'))  # nor the backdrop by a
  assert obs.layers['b'].sum() == 1 and obs.layers['b'][1, 2]


def test_fixture_per_env_actions_vs_oracle():
  """Every env gets its own random motions; auto-reset off; vs the oracle."""
  from pycolab_b200 import batched
  from pycolab_b200.games import fixtures
  torch = _torch()
  g = gc.load('fixture_scrolly_4')
  kw, cfg = gc.fixture_kwargs(g)
  B, T = 8, 120
  game = fixtures.make_game(kw['art'], kw['what_lies_beneath'], kw['walkers'],
                            kw['scrollys'], kw['drapes'], kw['update_schedule'],
                            kw['z_order'])
  eng = batched.BatchedEngine([game], batch=B, auto_reset=False)
  worlds = [ogames.make_fixture_world(**kw) for _ in range(B)]
  outs = [w.its_showtime() for w in worlds]
  res = eng.its_showtime()
  rs = np.random.RandomState(3)
  order = ''.join(eng.game.groups)
  alive = [True] * B
  for t in range(T + 1):
    torch.cuda.synchronize()
    boards = res.board.cpu().numpy()
    errs = eng.error_codes().cpu().numpy()
    for e in range(B):
      if alive[e]:
        np.testing.assert_array_equal(boards[e], outs[e][0], err_msg='t=%d e=%d' % (t, e))
        assert errs[e] == 0
    if t == T:
      break
    m = rs.randint(0, 9, size=B)
    rows = _rows(eng, {}, 1)[0]
    for e in range(B):
      rows[e, :len(order)] = m[e]            # one scrolling group: same motion for all
    res = eng.play(torch.from_numpy(rows).cuda())
    errs = eng.error_codes().cpu().numpy()
    for e in range(B):
      if not alive[e]:
        continue
      try:
        outs[e] = worlds[e].play(int(m[e]))
      except RuntimeError:                   # reference raises; device latches the bit
        alive[e] = False
        assert errs[e] & 1
  assert sum(alive) >= 1
