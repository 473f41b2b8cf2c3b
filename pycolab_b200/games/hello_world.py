"""Hello World set-up (reference `pycolab/examples/hello_world.py:36-118`): four
plain Sprites that slide diagonally and wrap around the board, and a Drape that
rolls its curtain along either axis; every move pays 1, action 4 quits.

Set-up only; per-step logic is csrc/hello.cu (the rolled curtain is the static
art shifted by two counters, never copied).
"""

from pycolab_b200 import ascii_art
from pycolab_b200 import things

HELLO_ART = ['                                    ',
             '  #   #  ### #    #     ###         ',
             '  #   # #    #    #    #   #        ',
             '  ##### #### #    #    #   #        ',
             '  #   # #    #    #    #   #        ',
             '  #   #  ###  ###  ###  ###         ',
             '                                    ',
             '     @   @  @@@   @@@  @    @@@@  1 ',
             '     @   @ @   @ @   @ @    @   @ 2 ',
             '     @ @ @ @   @ @@@@  @    @   @ 3 ',
             '     @ @ @ @   @ @   @ @    @   @   ',
             '      @@@   @@@  @   @  @@@ @@@@  4 ',
             '                                    ']


def make_game(art=None):
  """hello_world.py:58-68."""
  return ascii_art.ascii_art_to_game(
      art or HELLO_ART, what_lies_beneath=' ',
      sprites={'1': ascii_art.Partial(SlidingSprite, 0),
               '2': ascii_art.Partial(SlidingSprite, 1),
               '3': ascii_art.Partial(SlidingSprite, 2),
               '4': ascii_art.Partial(SlidingSprite, 3)},
      drapes={'@': RollingDrape},
      z_order='12@34')


class RollingDrape(things.Drape):
  """np.roll of the curtain by one cell per action 0-3 (:71-87)."""

  def update(self, actions, board, layers, backdrop, all_things, the_plot):
    raise NotImplementedError('runs on the device: csrc/hello.cu')


class SlidingSprite(things.Sprite):
  """Diagonal motion with wrap-around; `direction_set` picks the mapping (:90-118)."""
  _DX = ([-1, 1, -1, 1], [-1, 1, -1, 1], [1, -1, 1, -1], [1, -1, 1, -1])
  _DY = ([-1, 1, 1, -1], [1, -1, -1, 1], [1, -1, -1, 1], [-1, 1, 1, -1])

  def __init__(self, corner, position, character, direction_set):
    super(SlidingSprite, self).__init__(corner, position, character)
    self._dx = self._DX[direction_set]
    self._dy = self._DY[direction_set]

  def update(self, actions, board, layers, backdrop, all_things, the_plot):
    raise NotImplementedError('runs on the device: csrc/hello.cu')
