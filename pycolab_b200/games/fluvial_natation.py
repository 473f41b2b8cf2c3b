"""Fluvial Natation set-up (reference `pycolab/examples/fluvial_natation.py:28-110`).

A swimmer in a river that flows west: a MazeWalker plus a `Backdrop` WITH update
logic (rows 1..3 rotate one cell on even frames).  Set-up only; per-step logic is
the PCL_CLASSIC_FLUVIAL rule of csrc/classics.cu, where the flow is a rotation
count in the plot record and the backdrop array stays static.
"""

from pycolab_b200 import ascii_art
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import sprites as prefab_sprites

GAME_ART = ['===================================================',
            '     .      :   ,     `     ~          ,    .    ` ',
            '   ,    ~   P     :     .  `    ,    ,    ~    `   ',
            '     `   .     ~~   ,     .   :     .   `     `   ~',
            '===================================================']


def make_game(art=None):
  return ascii_art.ascii_art_to_game(art or GAME_ART, what_lies_beneath=' ',
                                     sprites={'P': PlayerSprite}, backdrop=RiverBackdrop)


class PlayerSprite(prefab_sprites.MazeWalker):
  """Actions 0, 1 = swim W, E; swept west on even frames; leaving the board to
  the east wins (+1), to the west loses (-1) (fluvial_natation.py:61-93)."""

  def __init__(self, corner, position, character):
    super(PlayerSprite, self).__init__(corner, position, character, impassable='')

  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/classics.cu')


class RiverBackdrop(plab_things.Backdrop):
  """Rows 1..3 rotate one cell west on even frames (fluvial_natation.py:96-110)."""

  def update(self, actions, board, layers, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/classics.cu')
