"""ASCII-art game builder (reference `pycolab/ascii_art.py:31-364`).

`ascii_art_to_game` keeps the reference signature and validation and drives the
same `Engine` set-up calls in the same order (ascii_art.py:241-289), so an
example's `make_game()` builds the same entity objects; the resulting `Engine`
is then lowered to the device at `its_showtime()`.  Host-only, runs once per
level.
"""

import itertools

import numpy as np

from pycolab_b200 import things


class Partial(object):
  """A thing class plus constructor arguments (ascii_art.py:331-364)."""

  def __init__(self, pycolab_thing, *args, **kwargs):
    if not issubclass(pycolab_thing, (things.Backdrop, things.Sprite, things.Drape)):
      raise TypeError('the pycolab_thing argument to ascii_art.Partial must be a '
                      'Backdrop, Sprite, or Drape subclass.')
    self.pycolab_thing = pycolab_thing
    self.args = args
    self.kwargs = kwargs


def ascii_art_to_uint8_nparray(art):
  """List of equal-length ASCII strings -> 2-D uint8 array (ascii_art.py:295-328)."""
  problem = ('the argument to ascii_art_to_uint8_nparray must be a list (or tuple) of '
             'strings containing the same number of strictly-ASCII characters.')
  try:
    rows = [np.frombuffer(line.encode('ascii'), dtype=np.uint8) for line in art]
    out = np.vstack(rows).copy()
  except AttributeError as e:
    if isinstance(art, (list, tuple)) and all(isinstance(r, (list, tuple)) for r in art):
      problem += ' Did you pass a list of list of single characters?'
    raise TypeError('{} (original error: {})'.format(problem, e))
  except ValueError as e:
    raise ValueError('{} (original error from numpy: {})'.format(problem, e))
  if np.any(out > 127):
    raise ValueError(problem)
  return out


def _as_partial(thing):
  return thing if isinstance(thing, Partial) else Partial(thing)


def ascii_art_to_game(art, what_lies_beneath, sprites=None, drapes=None,
                      backdrop=things.Backdrop, update_schedule=None, z_order=None,
                      occlusion_in_layers=True):
  """Build an `Engine` from ASCII art (ascii_art.py:31-292)."""
  from pycolab_b200 import engine
  sprites = {c: _as_partial(s) for c, s in (sprites or {}).items()}
  drapes = {c: _as_partial(d) for c, d in (drapes or {}).items()}
  backdrop = _as_partial(backdrop)
  entity_chars = set(sprites) | set(drapes)

  if update_schedule is None:
    update_schedule = list(entity_chars)
  if isinstance(update_schedule, str):
    update_schedule = list(update_schedule)
  if all(isinstance(item, str) for item in update_schedule):
    update_schedule = [update_schedule]
  try:
    flat_schedule = list(itertools.chain.from_iterable(update_schedule))
  except TypeError:
    raise TypeError('if any element in update_schedule is an iterable (like a list), '
                    'all elements in update_schedule must be')
  if set(flat_schedule) != entity_chars:
    raise ValueError('if specified, update_schedule must list each sprite and drape '
                     'exactly once.')
  if z_order is None:
    z_order = flat_schedule
  if set(z_order) != entity_chars:
    raise ValueError('if specified, z_order must list each sprite and drape exactly '
                     'once.')
  if isinstance(what_lies_beneath, str) and len(what_lies_beneath) != 1:
    raise ValueError('what_lies_beneath may either be a single-character ASCII string '
                     'or a list of ASCII-character strings')
  try:
    for group in (''.join(what_lies_beneath), entity_chars, z_order, flat_schedule):
      for ch in group:
        ord(ch)
  except TypeError:
    raise ValueError('keys of sprites, keys of drapes, what_lies_beneath (or its '
                     'entries), values in z_order, and (possibly nested) values in '
                     'update_schedule must all be single-character ASCII strings.')
  if entity_chars.intersection(''.join(what_lies_beneath)):
    raise ValueError('any character specified in what_lies_beneath must not be one of '
                     'the characters used as keys in the sprites or drapes arguments.')

  art = ascii_art_to_uint8_nparray(art)
  if isinstance(what_lies_beneath, str):
    beneath = np.full_like(art, ord(what_lies_beneath))
  else:
    beneath = ascii_art_to_uint8_nparray(what_lies_beneath)
    if beneath.shape != art.shape:
      raise ValueError('if not a single ASCII character, what_lies_beneath must be '
                       'ASCII art whose shape is the same as that of the ASCII art in '
                       'art.')

  group_of = {}
  for i, group in enumerate(update_schedule):
    for ch in group:
      group_of[ch] = '{:05d}'.format(i)

  game = engine.Engine(*art.shape, occlusion_in_layers=occlusion_in_layers)
  for ch in flat_schedule:
    game.update_group(group_of[ch])
    mask = art == ord(ch)
    if ch in drapes:
      part = drapes[ch]
      game.add_prefilled_drape(ch, mask, part.pycolab_thing, *part.args, **part.kwargs)
    if ch in sprites:
      rows, cols = np.where(mask)
      if len(rows) > 1:
        raise ValueError('sprite character {} can appear in at most one place in '
                         'art.'.format(ch))
      where = (int(rows[0]), int(cols[0])) if len(rows) else (0, 0)
      part = sprites[ch]
      game.add_sprite(ch, where, part.pycolab_thing, *part.args, **part.kwargs)
    art[mask] = beneath[mask]
  game.set_z_order(z_order)
  game.set_prefilled_backdrop(
      *backdrop.args, characters=''.join(chr(c) for c in np.unique(art)),
      prefill=art.view(np.uint8), backdrop_class=backdrop.pycolab_thing,
      **backdrop.kwargs)
  return game
