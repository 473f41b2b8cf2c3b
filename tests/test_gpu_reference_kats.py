"""The reference's own known-answer tests, replayed on the DEVICE.

Same fixtures as tests/test_reference_kats.py (16 `assertMachinima` calls, 127
frames of hand-drawn expected boards from the reference's maze_walker_test.py,
scrolling_test.py and cropping_test.py).  Each one is lowered to the general
device program from the captured engine state — entity registers, the
scrolling-protocol blackboard (order, egocentric set, permits), cropper corners
— and stepped through the C ABI; every frame's board and every cropper's view
must equal the HAND-DRAWN art.
"""

import numpy as np
import pytest

import reference_kats as rk
from oracle import engine_model as em

pytestmark = pytest.mark.gpu

KATS = rk.load()


def device_game(snap):
  """(BatchedEngine, LoweredGame) holding exactly the captured engine state."""
  import torch
  from pycolab_b200 import _lib, batched, engine as engine_lib, lowering, things
  from pycolab_b200.games import fixtures
  backdrop = rk.u8(snap['backdrop'])
  eng = engine_lib.Engine(snap['rows'], snap['cols'],
                          occlusion_in_layers=snap.get('occlusion_in_layers', True))
  eng.set_prefilled_backdrop(sorted(set(chr(c) for c in backdrop.ravel())), backdrop,
                             things.Backdrop)
  shape = (snap['rows'], snap['cols'])
  for i, group in enumerate(snap['groups']):
    eng.update_group('%02d' % i)
    for ch in group:
      if ch in snap['walkers']:
        w = snap['walkers'][ch]
        sprite = eng.add_sprite(ch, tuple(w['position']), fixtures.FixtureMazeWalker,
                                impassable=w['impassable'], confined_to_board=w['confined'],
                                egocentric_scroller=w['egocentric'])
        sprite._virtual_row, sprite._virtual_col = w['virtual_position']
        sprite._visible, sprite._prior_visible = w['visible'], w['prior_visible']
      elif ch in snap.get('sprites', {}):
        # A plain things.Sprite whose update does nothing (test_things.TestSprite):
        # a walker that is always told to stay.
        w = snap['sprites'][ch]
        sprite = eng.add_sprite(ch, tuple(w['position']), fixtures.FixtureMazeWalker,
                                impassable='')
        sprite._visible = w['visible']
      elif ch in snap.get('scrollys', {}):
        s = snap['scrollys'][ch]
        eng.add_drape(ch, fixtures.FixtureScrolly, board_shape=shape,
                      whole_pattern=rk.bits(s['pattern']),
                      board_northwest_corner=tuple(s['corner']),
                      scroll_margins=None if s['margins'] is None else tuple(s['margins']))
      else:
        eng.add_prefilled_drape(ch, rk.bits(snap['drapes'][ch]), fixtures.FixtureDrape)
  eng.set_z_order(snap['z_order'])
  game = lowering.lower(eng)
  assert game.program == _lib.PROG_FIXTURE
  # Registers the entity objects do not carry: pre-scroll corners, the frame, and
  # the default scrolling group's blackboard (protocols/scrolling.py:198-241).
  for d, ch in enumerate(game.drape_chars):
    if ch in snap.get('scrollys', {}):
      s = snap['scrollys'][ch]
      game.drapes[d, _lib.D_PRE_R], game.drapes[d, _lib.D_PRE_C] = s['prescroll']
      game.drapes[d, _lib.D_LAST_FRAME] = (_lib.NEVER if s['last_move_frame'] is None
                                           else s['last_move_frame'])
  game.plot[_lib.P_FRAME] = snap['frame']
  regs = snap.get('scrolling', {}).get('', None)
  assert set(snap.get('scrolling', {})) <= {''}, 'only the default scrolling group is lowered'
  if regs is not None:
    if regs['order'] is not None:
      game.plot[_lib.P_ORDER_R], game.plot[_lib.P_ORDER_C] = regs['order']
      game.plot[_lib.P_ORDER_FRAME] = regs['order_frame']
    ego = 0
    for ch in regs['egocentrists']:
      ego |= 1 << game.sprite_chars.index(ch)
    game.plot[_lib.P_EGO_MASK] = ego
    for ch, motions in regs['permitted'].items():
      i = game.sprite_chars.index(ch)
      mask = 0
      for m in motions:
        mask |= 1 << em.MOTIONS.index(tuple(m))
      game.sprites[i, _lib.S_AUX0] = mask
      game.sprites[i, _lib.S_AUX1] = regs['permitted_frame'][ch]
  b = batched.BatchedEngine([game], batch=1, auto_reset=False)
  b.its_showtime()                       # runs a frame 0 of its own: discard it ...
  b.sprites.copy_(b._sprites_init)       # ... and put the captured state back
  b.drapes.copy_(b._drapes_init)
  b.plot.copy_(b._plot_init)
  b.z_order.copy_(b._z_init)
  torch.cuda.synchronize()
  return b, game


def device_cropper(spec, b, game):
  from pycolab_b200 import _lib, batched
  if spec['kind'] == 'identity':
    return lambda: b.board[0].cpu().numpy()
  if spec['kind'] == 'fixed':
    crop = _lib.CropSpec(spec['rows'], spec['cols'], -1,
                         -1 if spec['pad_char'] is None else ord(spec['pad_char']),
                         0, 0, spec['top_left'][0], spec['top_left'][1], 0)
    return lambda: b.crop(crop)[0].cpu().numpy()
  codes = [game.sprite_chars.index(ch) + 1 if ch in game.sprite_chars
           else -(game.drape_chars.index(ch) + 1) for ch in spec['to_track']]
  crop = batched.scrolling_crop_spec(
      spec['rows'], spec['cols'], 0, track=codes,
      pad_char=spec['pad_char'], scroll_margins=tuple(spec['scroll_margins']),
      initial_offset=tuple(spec['initial_offset']), saccade=spec['saccade'])
  state = b.new_crop_state()
  if spec['corner'] is not None:         # row, col, initialised, episode (of the live plot: 0)
    state[0, 0], state[0, 1], state[0, 2], state[0, 3] = (spec['corner'][0],
                                                           spec['corner'][1], 1, 0)
  return lambda: b.crop(crop, state=state)[0].cpu().numpy().copy()


@pytest.mark.parametrize('kat', KATS, ids=rk.ids(KATS))
def test_device_reproduces_reference_kat(kat):
  import torch
  from pycolab_b200.games import fixtures
  b, game = device_game(kat['snapshot'])
  chars = ''.join(game.groups)
  croppers = None if kat['croppers'] is None else [device_cropper(c, b, game)
                                                   for c in kat['croppers']]
  for i, frame in enumerate(kat['frames']):
    motions = rk.motion_of(frame['action'], chars)
    if not isinstance(motions, dict):
      motions = {ch: motions for ch in chars}
    res = b.play([fixtures.action_rows(game, motions)])
    torch.cuda.synchronize()
    board = res.board[0].cpu().numpy()
    where = '%s frame %d' % (kat['test'], i)
    if croppers is None:
      np.testing.assert_array_equal(board, rk.u8(frame['art']), err_msg=where)
    else:
      for j, (crop, art) in enumerate(zip(croppers, frame['art'])):
        np.testing.assert_array_equal(crop(), rk.u8(art), err_msg='%s crop %d' % (where, j))
    np.testing.assert_array_equal(board, rk.u8(frame['board']), err_msg=where)
    assert int(res.has_reward[0]) == 0 and float(res.discount[0]) == frame['discount']
    sprites = b.sprites[0].cpu().numpy()
    if kat['test'].endswith('testNotConfinedToBoard'):     # maze_walker_test.py:383-390
      p = sprites[game.sprite_chars.index('P')]
      assert [[int(p[0]), int(p[1])], [int(p[2]), int(p[3])]] == frame['args'], where
  assert int(b.error_codes().abs().max()) == 0
