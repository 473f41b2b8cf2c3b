"""`MazeWalker` (reference `pycolab/prefab_parts/sprites.py:34-550`).

Set-up side of the MazeWalker prefab: constructor arguments, the virtual/true
position pair, and `_teleport` (which subclasses call from their constructors
to start off-board, e.g. examples/extraterrestrial_marauders.py:196).  The
per-step half — `_move` = obey scroll order -> `_check_motion` -> `_raw_move`
-> `_update_scroll_permissions` (sprites.py:356-546) — is the device function
`pcl::walker_move` in csrc/pcl_device.cuh; the nine motion helpers therefore
raise `DeviceOnlyError` if game code tries to run them from Python.
"""

from pycolab_b200 import things
from pycolab_b200.errors import DeviceOnlyError


class MazeWalker(things.Sprite):
  """A sprite that moves one cell at a time and respects impassable chars."""

  EDGE = 'edge!'

  _NORTH, _NORTHEAST, _EAST, _SOUTHEAST = (-1, 0), (-1, 1), (0, 1), (1, 1)
  _SOUTH, _SOUTHWEST, _WEST, _NORTHWEST = (1, 0), (1, -1), (0, -1), (-1, -1)
  _STAY = (0, 0)

  def __init__(self, corner, position, character, impassable,
               confined_to_board=False, egocentric_scroller=False,
               scrolling_group=''):
    super(MazeWalker, self).__init__(corner, position, character)
    for item in impassable:
      try:
        ord(item)
      except TypeError:
        raise TypeError(
            'the MazeWalker constructor requires all elements in its impassable '
            'argument to be single-character ASCII strings, but {!r} was found '
            'inside impassable.'.format(item))
    if character in impassable:
      raise ValueError('A MazeWalker must not designate its own character {!r} as '
                       'impassable.'.format(character))
    self._impassable = set(impassable)
    self._confined_to_board = confined_to_board
    self._egocentric_scroller = egocentric_scroller
    self._scrolling_group = scrolling_group
    self._virtual_row, self._virtual_col = position
    self._prior_visible = None

  @property
  def virtual_position(self):
    return self.Position(self._virtual_row, self._virtual_col)

  @property
  def on_the_board(self):
    return self._on_board(self._virtual_row, self._virtual_col)

  @property
  def impassable(self):
    return self._impassable

  # Board exit/entry hooks (sprites.py:223-275).  The device implements exactly
  # this default behaviour; overriding them is not lowered.
  def _on_board_exit(self):
    self._prior_visible = self._visible
    self._visible = False

  def _on_board_enter(self):
    self._visible = self._prior_visible

  def _teleport(self, virtual_position):
    """Place the walker at `virtual_position` (sprites.py:315-352)."""
    row, col = virtual_position
    was_on = self._on_board(self._virtual_row, self._virtual_col)
    now_on = self._on_board(row, col)
    if was_on and not now_on:
      self._on_board_exit()
    self._virtual_row, self._virtual_col = row, col
    self._position = self.Position(row, col) if now_on else self.Position(0, 0)
    if now_on and not was_on:
      self._on_board_enter()

  def _on_board(self, row, col):
    return 0 <= row < self.corner.row and 0 <= col < self.corner.col

  def _device_only(self, *unused_args, **unused_kwargs):
    raise DeviceOnlyError(
        'MazeWalker motion helpers run inside the CUDA step kernel '
        '(pcl::walker_move); pycolab_b200 never executes update() in Python.')

  _northwest = _north = _northeast = _east = _southeast = _device_only
  _south = _southwest = _west = _stay = _device_only
