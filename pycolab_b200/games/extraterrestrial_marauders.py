"""Extraterrestrial Marauders set-up
(reference `pycolab/examples/extraterrestrial_marauders.py:91-256`).

Set-up only; per-step logic (incl. the NumPy-compatible MT19937 draw for the
shooting marauder) is the fused kernel csrc/marauders.cu.
"""

from pycolab_b200 import ascii_art
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import sprites as prefab_sprites

UPWARD_BOLT_CHARS = 'abcd'
DOWNWARD_BOLT_CHARS = 'yz'


def make_game(art):
  sprites = {'P': PlayerSprite}
  sprites.update({c: UpwardLaserBoltSprite for c in UPWARD_BOLT_CHARS})
  sprites.update({c: DownwardLaserBoltSprite for c in DOWNWARD_BOLT_CHARS})
  return ascii_art.ascii_art_to_game(
      art, what_lies_beneath=' ', sprites=sprites,
      drapes={'X': MarauderDrape, 'B': BunkerDrape},
      update_schedule=['P', 'B', 'X'] + list(UPWARD_BOLT_CHARS + DOWNWARD_BOLT_CHARS))


class BunkerDrape(plab_things.Drape):
  """Bunkers eroded by bolts, -1 per hit (extraterrestrial_marauders.py:104-120)."""

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/marauders.cu')


class MarauderDrape(plab_things.Drape):
  """The marching marauders, +10 per hit (extraterrestrial_marauders.py:123-163)."""

  def __init__(self, curtain, character):
    super(MarauderDrape, self).__init__(curtain, character)
    self._dx = -1

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/marauders.cu')


class PlayerSprite(prefab_sprites.MazeWalker):
  """Left/right player confined to the board (:166-186)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(
        corner, position, character, impassable='', confined_to_board=True)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/marauders.cu')


class UpwardLaserBoltSprite(prefab_sprites.MazeWalker):
  """Player bolts; start hidden off-board (:189-220)."""

  def __init__(self, corner, position, character):
    super(UpwardLaserBoltSprite, self).__init__(corner, position, character, impassable='')
    self._teleport((-1, -1))

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/marauders.cu')


class DownwardLaserBoltSprite(prefab_sprites.MazeWalker):
  """Marauder bolts; start hidden off-board (:223-256)."""

  def __init__(self, corner, position, character):
    super(DownwardLaserBoltSprite, self).__init__(corner, position, character,
                                                  impassable='')
    self._teleport((-1, -1))

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/marauders.cu')
