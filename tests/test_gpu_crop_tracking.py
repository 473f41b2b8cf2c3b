"""ScrollingCropper with a `to_track` LIST (cropping.py:544-598): the window follows
the first visible entity; a drape's position is the median of its curtain cells.
Device (`pcl_crop_tracking`) against the oracle's ScrollingCrop."""

import numpy as np
import pytest

from oracle import engine_model as em
from oracle import games as ogames

pytestmark = pytest.mark.gpu

ART = ['..............',
       '..%%..........',
       '..%...........',
       '......P.......',
       '..........%%%.',
       '..........%...',
       '..............',
       '..............']


@pytest.mark.parametrize('to_track,pad,margins,saccade', [
    (['P', '%'], '.', (1, 2), True), (['%', 'P'], None, (1, 1), False), (['%'], '.', (1, 2), True)])
def test_tracking_list_with_a_drape(to_track, pad, margins, saccade):
  import torch
  from pycolab_b200 import cropping
  from pycolab_b200.games import fixtures
  walkers = {'P': dict(impassable='', confined=False, egocentric=False)}
  world = ogames.make_fixture_world(ART, '.', walkers, drapes='%')
  game = fixtures.make_game(ART, '.', walkers, drapes='%')
  want_crop = em.ScrollingCrop(3, 5, to_track, pad_char=pad, scroll_margins=margins,
                               saccade=saccade)
  got_crop = cropping.ScrollingCropper(3, 5, to_track, pad_char=pad, scroll_margins=margins,
                                       saccade=saccade)
  want_crop.set_engine(world)
  got_crop.set_engine(game)
  w_out, g_out = world.its_showtime(), game.its_showtime()
  rs = np.random.RandomState(2)
  # A drift to the north-west takes P off the board (invisible: the drape takes over)
  # and back again.
  moves = ['nw'] * 6 + ['se'] * 9 + list(
      rs.choice(['n', 'ne', 'e', 'se', 's', 'sw', 'w', 'nw'], size=60))
  hidden = 0
  for t, move in enumerate([None] + moves):
    if move is not None:
      w_out = world.play({'P': em.MOTION_OF_NAME[move]})
      g_out = game.play({'P': move})
    np.testing.assert_array_equal(g_out[0].board, w_out[0], err_msg='t=%d' % t)
    np.testing.assert_array_equal(got_crop.crop(g_out[0]).board, want_crop.crop(w_out[0]),
                                  err_msg='t=%d crop' % t)
    hidden += not world.things['P'].visible
  assert hidden > 3
