"""Scrolly Maze game set-up (reference `pycolab/examples/scrolly_maze.py:212-364`).

Set-up only: the constructors place the entities; the per-step logic of these
four classes is the fused kernel csrc/scrolly_maze.cu.  `make_game` takes the
level art as arguments (the reference indexes a built-in list), so stock and
generated levels (`pycolab_b200.levels.scrolly_maze_level`) load the same way.
"""

from pycolab_b200 import ascii_art
from pycolab_b200.prefab_parts import drapes as prefab_drapes
from pycolab_b200.prefab_parts import sprites as prefab_sprites


def make_game(maze_art, board_art, what_lies_beneath='#', corner_mark='+'):
  info = prefab_drapes.Scrolly.PatternInfo(
      maze_art, board_art, board_northwest_corner_mark=corner_mark,
      what_lies_beneath=what_lies_beneath)
  sprites = {'P': ascii_art.Partial(PlayerSprite, info.virtual_position('P'))}
  for ch in 'abc':
    sprites[ch] = ascii_art.Partial(PatrollerSprite, info.virtual_position(ch))
  return ascii_art.ascii_art_to_game(
      board_art, what_lies_beneath=' ', sprites=sprites,
      drapes={'#': ascii_art.Partial(MazeDrape, **info.kwargs('#')),
              '@': ascii_art.Partial(CashDrape, **info.kwargs('@'))},
      update_schedule=[['#'], ['a', 'b', 'c', 'P'], ['@']],
      z_order='abc@#P')


class PlayerSprite(prefab_sprites.MazeWalker):
  """Egocentric maze explorer; walls are impassable (scrolly_maze.py:245-271)."""

  def __init__(self, corner, position, character, virtual_position):
    super(PlayerSprite, self).__init__(
        corner, position, character, egocentric_scroller=True, impassable='#')
    self._teleport(virtual_position)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/scrolly_maze.cu')


class PatrollerSprite(prefab_sprites.MazeWalker):
  """Horizontal patroller, fatal on contact (scrolly_maze.py:274-305)."""

  def __init__(self, corner, position, character, virtual_position):
    super(PatrollerSprite, self).__init__(corner, position, character, '#')
    self._teleport(virtual_position)
    self._moving_east = bool(ord(character) % 2)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/scrolly_maze.cu')


class MazeDrape(prefab_drapes.Scrolly):
  """The scrolling walls (scrolly_maze.py:308-329)."""

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/scrolly_maze.cu')


class CashDrape(prefab_drapes.Scrolly):
  """The scrolling coins: +100 each, episode ends with the last one
  (scrolly_maze.py:332-364)."""

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/scrolly_maze.cu')
