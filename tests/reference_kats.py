"""Load the reference's own known-answer tests (tests/golden/reference_kats.json,
captured by tests/golden/make_kats.py from maze_walker_test.py, scrolling_test.py
and cropping_test.py) and rebuild their starting states on the oracle."""

import json
import os

import numpy as np

from oracle import engine_model as em
from oracle import games

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_kats.json')


def load():
  with open(PATH) as f:
    return json.load(f)['kats']


def ids(kats):
  seen, out = {}, []
  for k in kats:
    n = seen[k['test']] = seen.get(k['test'], 0) + 1
    out.append('%s-%d' % (k['test'].split('.')[-1], n))
  return out


def u8(art):
  return np.vstack([np.frombuffer(line.encode('ascii'), dtype=np.uint8) for line in art])


def bits(rows):
  return np.array([[c == '1' for c in row] for row in rows], dtype=bool)


def motion_of(action, chars):
  """Fixture action conventions (test_things.py:219-250, 268-295) -> the oracle's
  motion codes: one direction for everybody, or {char: direction}."""
  code = lambda d: em.MOTION_OF_NAME.get(d, em.M_STAY) if isinstance(d, str) else em.M_STAY
  if isinstance(action, dict):
    return {ch: code(action.get(ch)) for ch in chars}
  return code(action)


def result_code(result):
  """A recorded `walk_result_X` (JSON) in the oracle's encoding."""
  code = lambda x: em.EDGE if x == 'edge!' else ord(x)
  if result is None:
    return None
  if isinstance(result, list):
    return tuple(code(x) for x in result)
  return code(result)


class InertSprite(object):
  """A plain `things.Sprite` whose update does nothing (test_things.py TestSprite)."""
  is_sprite = True

  def __init__(self, position, visible):
    self.row, self.col = position
    self.vrow, self.vcol = position
    self.visible = visible


def oracle_world(snap):
  """An oracle World in exactly the state the reference Engine was in."""
  shape = (snap['rows'], snap['cols'])
  things = {}
  for ch, w in snap['walkers'].items():
    walker = em.Walker(ch, shape, tuple(w['position']), impassable=w['impassable'],
                       confined=w['confined'], egocentric=w['egocentric'], group=w['group'])
    walker.vrow, walker.vcol = w['virtual_position']
    walker.visible, walker.prior_visible = w['visible'], w['prior_visible']
    things[ch] = walker
  for ch, s in snap['scrollys'].items():
    drape = em.Scrolly(ch, shape, bits(s['pattern']), tuple(s['corner']),
                       margins=None if s['margins'] is None else tuple(s['margins']),
                       group=s['group'])
    drape.prescroll = tuple(s['prescroll'])
    drape.last_move_frame = s['last_move_frame']
    things[ch] = drape
  for ch, rows in snap['drapes'].items():
    things[ch] = em.PlainDrape(ch, bits(rows))
  for ch, s in snap.get('sprites', {}).items():
    things[ch] = InertSprite(tuple(s['position']), s['visible'])
  world = em.World(shape[0], shape[1], u8(snap['backdrop']), things, z_order=snap['z_order'],
                   groups=snap['groups'], program=games.fixture_program)
  world.plot.frame = snap['frame']
  for name, g in snap['scrolling'].items():
    regs = world.plot.group(name)
    regs.order = None if g['order'] is None else tuple(g['order'])
    regs.order_frame = g['order_frame']
    regs.ego = list(g['egocentrists'])
    regs.permit_frame = dict(g['permitted_frame'])
    regs.permits = {ch: set(tuple(m) for m in ms) for ch, ms in g['permitted'].items()}
  world._render()
  return world
