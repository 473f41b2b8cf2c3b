"""Cliff-walk set-up (reference `pycolab/examples/classics/cliff_walk.py:28-86`).

One MazeWalker confined to the board; every move costs 1.0, stepping onto the
cliff (bottom row between the end columns) costs 100.0, and the bottom row past
column 0 ends the episode.  Set-up only; per-step logic is csrc/classics.cu.
"""

from pycolab_b200 import ascii_art
from pycolab_b200.prefab_parts import sprites as prefab_sprites

GAME_ART = ['............',
            '............',
            '............',
            'P...........']


def make_game(art=None):
  return ascii_art.ascii_art_to_game(art or GAME_ART, what_lies_beneath='.',
                                     sprites={'P': PlayerSprite})


class PlayerSprite(prefab_sprites.MazeWalker):
  """Actions 0-3 = N, S, W, E (cliff_walk.py:46-86)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(corner, position, character, impassable='',
                                       confined_to_board=True)

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/classics.cu')
