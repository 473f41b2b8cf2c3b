"""Four-rooms set-up (reference `pycolab/examples/classics/four_rooms.py:28-85`).

One MazeWalker that cannot pass '#'; reaching cell (4, 3) pays 1.0 and ends the
episode.  Set-up only; per-step logic is csrc/classics.cu.
"""

from pycolab_b200 import ascii_art
from pycolab_b200.prefab_parts import sprites as prefab_sprites

GAME_ART = ['#############',
            '#     #     #',
            '#     #     #',
            '#     #     #',
            '#           #',
            '#     #     #',
            '#### ###### #',
            '#     #     #',
            '#     #     #',
            '#           #',
            '#     #     #',
            '# P   #     #',
            '#############']


def make_game(art=None):
  return ascii_art.ascii_art_to_game(art or GAME_ART, what_lies_beneath=' ',
                                     sprites={'P': PlayerSprite})


class PlayerSprite(prefab_sprites.MazeWalker):
  """Actions 0-3 = N, S, W, E (four_rooms.py:52-80)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(corner, position, character, impassable='#')

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/classics.cu')
