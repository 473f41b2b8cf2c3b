"""Chain-walk set-up (reference `pycolab/examples/classics/chain_walk.py:28-73`).

One MazeWalker on a single row; the left end pays 1.0, the right end 100.0, and
either ends the episode.  Set-up only; per-step logic is csrc/classics.cu.
"""

from pycolab_b200 import ascii_art
from pycolab_b200.prefab_parts import sprites as prefab_sprites

GAME_ART = ['..P...................']


def make_game(art=None):
  return ascii_art.ascii_art_to_game(art or GAME_ART, what_lies_beneath='.',
                                     sprites={'P': PlayerSprite})


class PlayerSprite(prefab_sprites.MazeWalker):
  """Actions 0, 1 = W, E (chain_walk.py:44-73)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(corner, position, character, impassable='')

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/classics.cu')
