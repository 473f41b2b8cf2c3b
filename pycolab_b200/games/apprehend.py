"""Apprehend set-up (reference `pycolab/examples/apprehend.py:34-131`): a ball falls
along a random straight line, the player slides along the bottom row to catch it.

Set-up only; per-step logic is csrc/apprehend.cu.  As upstream, the ball's slope is
drawn from Python's global `random` when the sprite is BUILT (:103), so
`random.seed(s); make_game()` builds the same game here and there.  A batched
engine with auto-reset draws every episode's slope on the device from a per-env
`random.Random(seed)` state instead.
"""

import random

from pycolab_b200 import ascii_art
from pycolab_b200.prefab_parts import sprites as prefab_sprites

GAME_ART = ['   b   ',
            '       ',
            '       ',
            '       ',
            '       ',
            '       ',
            '       ',
            '       ',
            '       ',
            '   P   ']

# In Catch, both the ball and the player look identical (apprehend.py:50).
REPAINT_MAPPING = {'b': 'X', 'P': 'X'}


def make_game(art=None):
  """apprehend.py:56-60."""
  return ascii_art.ascii_art_to_game(
      art or GAME_ART, what_lies_beneath=' ',
      sprites={'P': PlayerSprite, 'b': BallSprite},
      update_schedule=['b', 'P'])


class PlayerSprite(prefab_sprites.MazeWalker):
  """Left / right along its row; catching the ball pays 1 and ends the game (:63-87)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(
        corner, position, character, impassable='', confined_to_board=True)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/apprehend.cu')


class BallSprite(prefab_sprites.MazeWalker):
  """Falls one row per frame, drifting by `_dx` columns per row (:90-131)."""

  def __init__(self, corner, position, character):
    super(BallSprite, self).__init__(corner, position, character, impassable='')
    self._dx = random.uniform(-2.499, 2.499) / (corner[0] - 1.0)
    self._x_accumulator = 0.0

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/apprehend.cu')
