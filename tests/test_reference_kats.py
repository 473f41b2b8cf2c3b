"""The reference's own known-answer tests, replayed on the oracle (CPU).

Every `assertMachinima` call of the reference's maze_walker_test.py,
scrolling_test.py and cropping_test.py (16 calls, 127 frames) is a list of
(action, hand-drawn expected board[, expected motion result]).  The oracle
starts from the captured engine state and must reproduce the HAND-DRAWN art of
every frame (and every cropper's view), plus the rewards, discounts and motion
results the reference produced while passing its own test.
"""

import numpy as np
import pytest

import reference_kats as rk
from oracle import engine_model as em

KATS = rk.load()


def oracle_cropper(spec, world):
  if spec['kind'] == 'identity':
    return lambda board: board
  if spec['kind'] == 'fixed':
    return lambda board: em.crop_window(board, tuple(spec['top_left']), spec['rows'],
                                        spec['cols'], spec['pad_char'])
  crop = em.ScrollingCrop(spec['rows'], spec['cols'], spec['to_track'],
                          pad_char=spec['pad_char'], scroll_margins=tuple(spec['scroll_margins']),
                          initial_offset=tuple(spec['initial_offset']), saccade=spec['saccade'])
  crop.set_engine(world)
  crop.corner = None if spec['corner'] is None else tuple(spec['corner'])
  return crop.crop


@pytest.mark.parametrize('kat', KATS, ids=rk.ids(KATS))
def test_oracle_reproduces_reference_kat(kat):
  world = rk.oracle_world(kat['snapshot'])
  chars = list(world.things)
  croppers = None if kat['croppers'] is None else [oracle_cropper(c, world)
                                                   for c in kat['croppers']]
  for i, frame in enumerate(kat['frames']):
    board, reward, discount = world.play(rk.motion_of(frame['action'], chars))
    where = '%s frame %d' % (kat['test'], i)
    if croppers is None:
      np.testing.assert_array_equal(board, rk.u8(frame['art']), err_msg=where)
    else:
      for j, (crop, art) in enumerate(zip(croppers, frame['art'])):
        np.testing.assert_array_equal(crop(board), rk.u8(art), err_msg='%s crop %d' % (where, j))
    np.testing.assert_array_equal(board, rk.u8(frame['board']), err_msg=where)
    assert (reward, discount, world.game_over) == (frame['reward'], frame['discount'],
                                                   frame['game_over']), where
    for ch, result in frame['walks'].items():
      assert world.things[ch].last_result == rk.result_code(result), (where, ch)
    name = kat['test'].split('.')[-1]
    if name in ('testBasicWalking', 'testConfinedToBoard'):
      # the test's own expectation for P's motion result (maze_walker_test.py:62-71, 499-504)
      assert world.things['P'].last_result == rk.result_code(frame['args'][0]), where
    elif name == 'testNotConfinedToBoard':
      # ... and for P's true and virtual positions (maze_walker_test.py:383-390)
      p = world.things['P']
      assert [list(p.position), list(p.virtual_position)] == frame['args'], where


def test_every_machinima_of_the_reference_suites_is_covered():
  tests = sorted(set(k['test'] for k in KATS))
  assert len(KATS) == 16 and sum(len(k['frames']) for k in KATS) == 127
  assert [t.split('.')[-1] for t in tests] == [
      'testDefaultCropper', 'testEgocentricScrolling', 'testFixedCropper',
      'testScrollingInitialOffset', 'testScrollingMargins', 'testScrollingSaccade',
      'testWeirdFixedCrops', 'testBasicWalking', 'testConfinedToBoard',
      'testNotConfinedToBoard', 'testScrolly']
