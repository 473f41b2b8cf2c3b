"""General-purpose entities: plain MazeWalkers, Scrollys and static drapes.

Counterparts of the reference's test fixtures (`pycolab/tests/test_things.py`:
`TestMazeWalker` :203-250, `TestScrolly` :253-295, `TestDrape` :178-200): each
entity simply performs the motion named by its action slot every frame
(`_stay` when none is given).  Their per-step logic is the general device
program csrc/fixture.cu (`PCL_PROG_FIXTURE`), which also accepts Plot
directives (reward, termination, z-order change) as extra action words — the
device stand-in for `test_things.post_update` code injection.

`make_game` mirrors `oracle.games.make_fixture_world`, so one description
builds the reference fixture game, the oracle world and the device game.
"""

import numpy as np

from pycolab_b200 import _lib
from pycolab_b200 import ascii_art
from pycolab_b200 import things as plab_things
from pycolab_b200.prefab_parts import drapes as prefab_drapes
from pycolab_b200.prefab_parts import sprites as prefab_sprites

NO_REWARD = -(2 ** 31)


class FixtureMazeWalker(prefab_sprites.MazeWalker):
  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/fixture.cu')


class FixtureScrolly(prefab_drapes.Scrolly):
  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/fixture.cu')


class FixtureDrape(plab_things.Drape):
  def update(self, actions, board, layers, backdrop, things, the_plot):
    raise NotImplementedError('runs on the device: csrc/fixture.cu')


def make_game(art, what_lies_beneath, walkers, scrollys=None, drapes='',
              update_schedule=None, z_order=None, occlusion_in_layers=True):
  """walkers: {char: dict(impassable, confined, egocentric, group)}; scrollys:
  {char: dict(pattern, corner, margins, group)}; drapes: chars of static drapes."""
  scrollys = scrollys or {}
  shape = (len(art), len(art[0]))
  sprites = {
      ch: ascii_art.Partial(FixtureMazeWalker, impassable=kw.get('impassable', ''),
                            confined_to_board=kw.get('confined', False),
                            egocentric_scroller=kw.get('egocentric', False),
                            scrolling_group=kw.get('group', ''))
      for ch, kw in walkers.items()}
  dr = {
      ch: ascii_art.Partial(FixtureScrolly, board_shape=shape,
                            whole_pattern=np.array(kw['pattern'], dtype=bool),
                            board_northwest_corner=tuple(kw['corner']),
                            scroll_margins=kw.get('margins', (2, 3)),
                            scrolling_group=kw.get('group', ''))
      for ch, kw in scrollys.items()}
  for ch in drapes:
    dr[ch] = FixtureDrape
  chars = list(walkers) + list(scrollys) + list(drapes)
  if update_schedule is None:
    update_schedule = [chars]
  return ascii_art.ascii_art_to_game(art, what_lies_beneath, sprites, dr,
                                     update_schedule=update_schedule, z_order=z_order,
                                     occlusion_in_layers=occlusion_in_layers)


def action_rows(game_or_lowered, motions, reward=None, terminate=False, z=None,
                directives=None):
  """One device action row: motions {char: code} (missing = stay) in update
  order, then the Plot directives of the step as (opcode, argument) pairs.

  directives: ordered list of calls, e.g. [('add_reward', 3),
  ('terminate_episode', 0.5), ('change_default_discount', 0.9),
  ('change_z_order', 'b', None)]; or the shorthands reward / terminate / z
  (applied in that order)."""
  import struct
  order = ''.join(game_or_lowered.groups)
  row = [int(motions.get(ch, 8)) for ch in order]
  if directives is None:
    directives = []
    if reward is not None:
      directives.append(('add_reward', reward))
    if terminate:
      directives.append(('terminate_episode', 0.0))
    if z is not None:
      directives.append(('change_z_order', z[0], z[1]))
  if len(directives) > _lib.FIXTURE_DIRECTIVES:
    raise ValueError('at most {} Plot directives per step are lowered'.format(
        _lib.FIXTURE_DIRECTIVES))
  f32 = lambda x: struct.unpack('<i', struct.pack('<f', float(x)))[0]
  for call in directives:
    name, args = call[0], call[1:]
    if name == 'add_reward':
      row += [_lib.DIR_ADD_REWARD, int(args[0])]
    elif name == 'terminate_episode':
      discount = args[0] if args else 0.0
      if not 0.0 <= discount <= 1.0:
        raise ValueError('Discount must be in range [0,1].')
      row += [_lib.DIR_TERMINATE, f32(discount)]
    elif name == 'change_default_discount':
      if not 0.0 <= args[0] <= 1.0:
        raise ValueError('Default discount must be in range [0,1].')
      row += [_lib.DIR_DEFAULT_DISCOUNT, f32(args[0])]
    elif name == 'change_z_order':
      row += [_lib.DIR_Z_ORDER, ord(args[0]) | ((0 if args[1] is None else ord(args[1])) << 8)]
    else:
      raise ValueError('unknown Plot directive {!r}'.format(name))
  row += [_lib.DIR_NONE, 0] * (_lib.FIXTURE_DIRECTIVES - len(directives))
  return row
