"""Shockwave set-up (reference `pycolab/examples/shockwave.py:40-197`): climb to the
safe top row while rings of fire expand from random impact points; bunkers ('+')
shelter the player, walls ('=') stop both the player and the fire.

Set-up only; per-step logic is csrc/shockwave.cu.  The impact points come from
NumPy's global generator (`np.random.randint`, :133), whose MT19937 state the facade
hands to the device and takes back every step, as for extraterrestrial_marauders.
"""

import numpy as np

from pycolab_b200 import ascii_art
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import sprites as prefab_sprites

LEVELS = [
    ['^^^^^^^^^^^^^^^',
     '               ',
     '  +           +',
     '  ==   ++  == +',
     '              +',
     '=======       +',
     ' +            +',
     '   +      ++   ',
     '+        ==    ',
     '+        +     ',
     '   =           ',
     ' +++ P    ++   '],
]


def make_game(level=0):
  """shockwave.py:181-197; `level` is an index into LEVELS or the art itself."""
  level_art = LEVELS[level] if isinstance(level, int) else level
  return ascii_art.ascii_art_to_game(
      level_art,
      what_lies_beneath='+',
      sprites={'P': PlayerSprite},
      drapes={'@': ShockwaveDrape, ' ': MinimalDrape, '^': MinimalDrape},
      update_schedule=[' ', '^', 'P', '@'],
      z_order=[' ', '^', '@', 'P'],
  )


class PlayerSprite(prefab_sprites.MazeWalker):
  """Up / left / right / stay; walls are impassable (:91-109)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(
        corner, position, character, impassable='=', confined_to_board=True)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/shockwave.cu')


class ShockwaveDrape(plab_things.Drape):
  """A ring `width` cells thick around a random impact point, one cell wider per frame;
  decides who wins and who burns (:112-165)."""

  def __init__(self, curtain, character, width=2):
    super(ShockwaveDrape, self).__init__(curtain, character)
    self._width = width
    self._distance_from_impact = np.zeros(self.curtain.shape)
    self._steps_since_impact = 0

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/shockwave.cu')


class MinimalDrape(plab_things.Drape):
  """Holds a curtain, no logic (:168-172)."""

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/shockwave.cu')
