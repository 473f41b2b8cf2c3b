"""Scrolling protocol names (reference `pycolab/protocols/scrolling.py:255-578`).

Upstream this module is a set of functions that keep scrolling orders/permits
in Plot dict keys.  Here that bookkeeping is per-env device state (plot record
words PCL_P_ORDER_R/C, PCL_P_ORDER_FRAME, PCL_P_EGO_MASK and the permit
mask/frame words of each egocentric sprite; see csrc/pcl_device.cuh
`scroll_is_possible`, `walker_move`, `scrolly_move`).  The module keeps the
motion constants and the `Error` type so game code imports resolve; the
functions exist for set-up-time use on a host `Plot` with upstream semantics.
"""

from pycolab_b200 import things

NORTH, NORTHEAST, EAST, SOUTHEAST = (-1, 0), (-1, 1), (0, 1), (1, 1)
SOUTH, SOUTHWEST, WEST, NORTHWEST = (1, 0), (1, -1), (0, -1), (-1, -1)


class Error(RuntimeError):
  """Mishandling of the scrolling protocol (scrolling.py:279)."""


def _key(group, what):
  return 'scrolling_{}_{}'.format(group, what)


def _check(entity, the_plot, group):
  if not isinstance(entity, (things.Backdrop, things.Drape, things.Sprite)):
    raise TypeError('an object that was not a pycolab game entity ({}) attempted to '
                    'use the scrolling protocol.'.format(entity))
  known = the_plot.setdefault('scrolling_everyone', {}).setdefault(entity, group)
  if known != group:
    raise Error('{} has attempted to participate in the scrolling protocol as part of '
                'scrolling group {}, but is already known to belong to scrolling group '
                '{}.'.format(_who(entity), repr(group), repr(known)))


def _who(entity):
  """How error messages name an entity (scrolling.py:572-578)."""
  character = getattr(entity, 'character', None)
  if character is not None:
    return 'a Sprite or Drape handling character {}'.format(repr(character))
  return 'the Backdrop'


def participate_as_egocentric(entity, the_plot, scrolling_group=''):
  _check(entity, the_plot, scrolling_group)
  the_plot.setdefault(_key(scrolling_group, 'egocentrists'), set()).add(entity)


def egocentric_participants(entity, the_plot, scrolling_group=''):
  _check(entity, the_plot, scrolling_group)
  return the_plot.get(_key(scrolling_group, 'egocentrists'), set())


def get_order(entity, the_plot, scrolling_group=''):
  _check(entity, the_plot, scrolling_group)
  if the_plot.setdefault(_key(scrolling_group, 'order_frame'), None) != the_plot.frame:
    return None
  return the_plot.setdefault(_key(scrolling_group, 'order'), None)


def permit(entity, the_plot, motions, scrolling_group=''):
  _check(entity, the_plot, scrolling_group)
  if entity not in the_plot.setdefault(_key(scrolling_group, 'egocentrists'), set()):
    raise Error('{} is not registered as an egocentric entity in scrolling group '
                '{}'.format(_who(entity), repr(scrolling_group)))
  valid_at = the_plot.frame + 1
  frames = the_plot.setdefault(_key(scrolling_group, 'permitted_frame'), {})
  mine = the_plot.setdefault(_key(scrolling_group, 'permitted'), {}).setdefault(
      entity, set())
  if frames.setdefault(entity, valid_at) != valid_at:
    frames[entity] = valid_at
    mine.clear()
  mine.update(motions)


def is_possible(entity, the_plot, motion, scrolling_group=''):
  _check(entity, the_plot, scrolling_group)
  frames = the_plot.get(_key(scrolling_group, 'permitted_frame'), {})
  permits = the_plot.get(_key(scrolling_group, 'permitted'), {})
  for other in the_plot.get(_key(scrolling_group, 'egocentrists'), set()):
    if frames.get(other) != the_plot.frame or motion not in permits.get(other, ()):
      return False
  return True


def order(entity, the_plot, motion, scrolling_group='', check_possible=True):
  _check(entity, the_plot, scrolling_group)
  if the_plot.setdefault(_key(scrolling_group, 'order_frame'), None) == the_plot.frame:
    raise Error('{} attempted to issue a second scrolling order for scrolling group {}.'
                ''.format(_who(entity), repr(scrolling_group)))
  if check_possible and not is_possible(entity, the_plot, motion, scrolling_group):
    raise Error('{} attempted to order an impossible scrolling motion "{}" for scrolling '
                'group {}.'.format(_who(entity), motion, repr(scrolling_group)))
  the_plot[_key(scrolling_group, 'order_frame')] = the_plot.frame
  the_plot[_key(scrolling_group, 'order')] = motion
