"""Warehouse Manager set-up (reference `pycolab/examples/warehouse_manager.py:139-295`).

Set-up only; per-step logic is the fused kernel csrc/warehouse.cu.
"""

from pycolab_b200 import ascii_art
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import sprites as prefab_sprites

_BOX_ORDER = '1234567890'


def make_game(art, what_lies_beneath=' '):
  flat = ''.join(art)
  boxes = [c for c in _BOX_ORDER if c in flat]
  sprites = {c: BoxSprite for c in boxes}
  sprites['P'] = PlayerSprite
  return ascii_art.ascii_art_to_game(
      art, what_lies_beneath, sprites, {'X': JudgeDrape},
      update_schedule=[boxes, ['X'], ['P']])


class BoxSprite(prefab_sprites.MazeWalker):
  """A box: moves only when pushed by the player (warehouse_manager.py:181-226)."""

  def __init__(self, corner, position, character):
    super(BoxSprite, self).__init__(
        corner, position, character, set('#.0123456789PX') - set(character))

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/warehouse.cu')


class JudgeDrape(plab_things.Drape):
  """Marks boxes on goals, scores, decides termination (warehouse_manager.py:229-266)."""

  def __init__(self, curtain, character):
    super(JudgeDrape, self).__init__(curtain, character)
    self._last_num_boxes_on_goals = 0

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/warehouse.cu')


class PlayerSprite(prefab_sprites.MazeWalker):
  """The warehouse manager (warehouse_manager.py:269-295)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(
        corner, position, character, impassable='#.0123456789X')

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/warehouse.cu')
