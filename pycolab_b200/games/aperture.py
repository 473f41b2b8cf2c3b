"""Aperture set-up (reference `pycolab/examples/aperture.py:118-196`).

A player with a blaster: shots that meet a special wall '@' open an aperture
there (two at most, the older one closes), and stepping into an aperture comes
out of the other.  Reaching the cranachan 'C' pays 1 and ends the episode.
Set-up only; per-step logic is csrc/aperture.cu.  Levels are passed in as art
(the reference ships three; `pycolab_b200.levels.aperture_level()` draws another).
"""

from pycolab_b200 import ascii_art
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import sprites as prefab_sprites


def make_game(art):
  return ascii_art.ascii_art_to_game(
      art=art, what_lies_beneath=' ', sprites={'A': PlayerSprite}, drapes={'X': ApertureDrape},
      update_schedule=[['A'], ['X']], z_order=['X', 'A'])


class PlayerSprite(prefab_sprites.MazeWalker):
  """Actions 0-3 walk N, S, W, E; 9 quits (aperture.py:118-149)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(corner, position, character, impassable='#.@')

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/aperture.cu')


class ApertureDrape(plab_things.Drape):
  """Actions 5-8 fire up, left, down, right (aperture.py:152-185)."""

  def __init__(self, curtain, character):
    super(ApertureDrape, self).__init__(curtain, character)
    self._apertures = [None, None]

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/aperture.cu')

  @property
  def apertures(self):
    return tuple(a for a in self._apertures if a is not None)
