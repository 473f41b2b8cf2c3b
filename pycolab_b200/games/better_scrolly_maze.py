"""Better Scrolly Maze set-up (reference `pycolab/examples/better_scrolly_maze.py:209-324`).

The cropper-based maze: the board is the whole world, the walls live in the
backdrop, one update group, and the egocentric views come from
`cropping.ScrollingCropper` / `FixedCropper` applied after the step (device:
`pcl_crop`).  Set-up only; per-step logic is csrc/better_scrolly.cu.
"""

from pycolab_b200 import ascii_art
from pycolab_b200 import cropping
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import sprites as prefab_sprites


def make_game(maze_art):
  return ascii_art.ascii_art_to_game(
      maze_art, what_lies_beneath=' ',
      sprites={'P': PlayerSprite, 'a': PatrollerSprite, 'b': PatrollerSprite,
               'c': PatrollerSprite},
      drapes={'@': CashDrape},
      update_schedule=['a', 'b', 'c', 'P', '@'], z_order='abc@P')


def make_croppers(starter_offset=(0, 0), teaser_corner=(0, 0)):
  """The three views of better_scrolly_maze.py:224-251: the player's, patroller
  c's, and a fixed "teaser" window."""
  return [
      cropping.ScrollingCropper(rows=10, cols=30, to_track=['P'],
                                initial_offset=starter_offset),
      cropping.ScrollingCropper(rows=7, cols=10, to_track=['c'], pad_char=' ',
                                scroll_margins=(None, 3)),
      cropping.FixedCropper(top_left_corner=teaser_corner, rows=12, cols=20, pad_char=' '),
  ]


class PlayerSprite(prefab_sprites.MazeWalker):
  """The maze explorer (better_scrolly_maze.py:254-276)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(corner, position, character, impassable='#')

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/better_scrolly.cu')


class PatrollerSprite(prefab_sprites.MazeWalker):
  """Horizontal patroller, fatal on contact (better_scrolly_maze.py:279-305)."""

  def __init__(self, corner, position, character):
    super(PatrollerSprite, self).__init__(corner, position, character, impassable='#')
    self._moving_east = bool(ord(character) % 2)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/better_scrolly.cu')


class CashDrape(plab_things.Drape):
  """Coins: +100 each, episode ends with the last (better_scrolly_maze.py:308-324)."""

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/better_scrolly.cu')
