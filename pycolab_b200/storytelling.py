"""Stories: several games played back to back as one episode.

Host-side mirror of the reference's `pycolab/storytelling.py:35-654`.  A `Story`
is a CALLER of the hot path, not part of it: every chapter is an ordinary
`engine.Engine` (here: a game lowered to a device program) and the story only
decides which one `its_showtime()` / `play()` go to.  Behaviour follows the
reference:

* `chapters` is a dict (the finishing game names its successor through
  `Plot.next_chapter`) or a list/tuple (chapters follow one another; a game may
  still redirect through `Plot.next_chapter`) — storytelling.py:105-170.
* When a chapter ends inside `its_showtime()`/`play()`, the next one is started
  in the same call; the ended chapter's last observation and discount are dropped
  and its last reward is added to the first reward of whatever starts next
  (storytelling.py:391-474).  The old Plot's entries are copied into the new one.
* All chapters must produce observations of one shape (croppers may help) and
  use each character consistently as Sprite, Drape or backdrop
  (storytelling.py:556-622).

On the device nothing changes: the chapter in play is one batch-1 engine.  Games
whose entities run on the device cannot set `Plot.next_chapter` from `update()`;
dict-style stories therefore need builders that set it on the Engine's plot
before returning it (or use list-style sequencing).
"""

import collections.abc

import numpy as np

from pycolab_b200 import cropping
from pycolab_b200 import engine as engine_lib
from pycolab_b200 import things


class Story(object):
  """A programmable sequence of mutually compatible games (storytelling.py:35)."""

  def __init__(self, chapters, first_chapter=None, croppers=None):
    self._auto_advance = not isinstance(chapters, collections.abc.Mapping)
    if self._auto_advance and first_chapter is None:
      first_chapter = 0
    self._chapters, self._croppers = _normalise(chapters, first_chapter, croppers)
    (self._chars_sprites, self._chars_drapes, self._chars_backdrops,
     (self._rows, self._cols)) = _survey_games(self._chapters, self._croppers)

    self._showtime = False
    self._game_over = False
    self._dummy_sprites = {}
    self._dummy_drapes = {}
    self._current_game = None
    self._current_cropper = None
    self._install(first_chapter, prior=None, old_plot=None)

  # ------------------------------------------------------------ Engine-like API
  def its_showtime(self):
    if self._showtime:
      raise RuntimeError('its_showtime should not be called more than once.')
    self._showtime = True
    return self._deliver(self._current_game.its_showtime())

  def play(self, actions):
    if not self._showtime:
      raise RuntimeError('play() cannot be called until the Story is placed in '
                         '"play mode" via the its_showtime() method.')
    if self._game_over:
      raise RuntimeError('play() was called after the last game managed by the '
                         'Story has terminated.')
    return self._deliver(self._current_game.play(actions))

  @property
  def the_plot(self):
    return self._current_game.the_plot

  @property
  def rows(self):
    return self._rows

  @property
  def cols(self):
    return self._cols

  @property
  def game_over(self):
    return self._game_over

  @property
  def z_order(self):
    """The current game's z-order with every other chapter's Sprite and Drape
    characters (sorted) underneath (storytelling.py:309-324)."""
    current = self._current_game.z_order
    others = (self._chars_sprites | self._chars_drapes) - set(current)
    return sorted(others) + current

  @property
  def backdrop(self):
    """Current curtain, palette of all chapters (storytelling.py:327-342)."""
    return things.Backdrop(curtain=self._current_game.backdrop.curtain,
                           palette=engine_lib.Palette(self._chars_backdrops))

  @property
  def things(self):
    """The current game's entities plus invisible stand-ins for the characters
    only other chapters use (storytelling.py:345-376)."""
    out = self._current_game.things
    shape = (self._current_game.rows, self._current_game.cols)
    for ch in self._chars_sprites:
      if ch not in out:
        if shape not in self._dummy_sprites:
          self._dummy_sprites[shape] = _DummySprite(things.Sprite.Position(*shape), ch)
        out[ch] = self._dummy_sprites[shape]
    for ch in self._chars_drapes:
      if ch not in out:
        if shape not in self._dummy_drapes:
          self._dummy_drapes[shape] = _DummyDrape(np.zeros(shape, dtype=bool), ch)
        out[ch] = self._dummy_drapes[shape]
    return out

  @property
  def current_game(self):
    return self._current_game

  # ------------------------------------------------------------------ internals
  def _install(self, key, prior, old_plot):
    """Build chapter `key`, hand it the previous Plot's contents and the chapter
    bookkeeping (storytelling.py:148-160, 446-463)."""
    game = self._chapters[key]()
    plot = game.the_plot
    if old_plot is not None:
      plot.update(old_plot)
    plot.prior_chapter = prior
    plot.this_chapter = key
    if self._auto_advance:
      plot.next_chapter = key + 1 if (key + 1) in self._chapters else None
    self._current_game = game
    self._current_cropper = self._croppers[key]
    self._current_cropper.set_engine(game)

  def _deliver(self, step):
    """Crop one chapter step; if the chapter ended, keep starting successors
    until one survives its first frame or none is left (storytelling.py:391-474)."""
    observation, reward, discount = step
    observation = self._current_cropper.crop(observation)
    while self._current_game.game_over:
      old_plot = self._current_game.the_plot
      successor = old_plot.next_chapter
      if successor is None:
        self._game_over = True
        break
      if successor not in self._chapters:
        raise KeyError(
            'The game that just finished in the Story currently underway (identified by '
            'the key/index "{}") said that the next game in the story should be {}, but '
            'no game was supplied to the Story constructor under that key or '
            'index.'.format(old_plot.this_chapter, repr(successor)))
      self._install(successor, prior=old_plot.this_chapter, old_plot=old_plot)
      observation, more, discount = self._current_game.its_showtime()
      observation = self._current_cropper.crop(observation)
      if more is not None:
        reward = more if reward is None else reward + more
    return observation, reward, discount


def is_fictional(thing):
  """Is `thing` one of the stand-ins `Story.things` invents (storytelling.py:477)?"""
  return isinstance(thing, (_DummySprite, _DummyDrape))


def _normalise(chapters, first_chapter, croppers):
  """Argument checks of storytelling.py:493-553: chapters -> dict of builders,
  croppers -> dict of ObservationCroppers with the same keys."""
  if not chapters:
    raise ValueError('The chapters argument to the Story constructor must not be empty.')
  if isinstance(chapters, collections.abc.Sequence):
    chapters = dict(enumerate(chapters))
    if isinstance(croppers, collections.abc.Sequence):
      croppers = dict(enumerate(croppers))
  if not isinstance(chapters, collections.abc.Mapping):
    raise ValueError('The chapters argument to the Story constructor must be either a '
                     'dict or a list.')
  if None in chapters:
    raise ValueError('None may not be a key in a Story chapters dict.')
  if first_chapter not in chapters:
    raise ValueError('The key "{}", specified as a Story\'s first_chapter, does not appear '
                     'in the chapters supplied to the Story constructor.'.format(first_chapter))
  if croppers is None:
    croppers = cropping.ObservationCropper()
  if isinstance(croppers, cropping.ObservationCropper):
    croppers = {key: croppers for key in chapters}
  if (not isinstance(croppers, collections.abc.Mapping) or
      set(croppers) != set(chapters)):
    raise ValueError('Since the croppers argument to the Story constructor was not None or '
                     'a single ObservationCropper, it must be a collection with the same '
                     'keys or indices as the chapters argument.')
  croppers = {key: cropping.ObservationCropper() if c is None else c
              for key, c in croppers.items()}
  return dict(chapters), croppers


def _survey_games(chapters, croppers):
  """Build and start every chapter once to learn the observation shape and how
  characters are used; reject incompatible games (storytelling.py:556-622)."""
  shapes, sprites, drapes, backdrops = set(), set(), set(), set()
  for key in sorted(chapters):
    game, cropper = chapters[key](), croppers[key]
    cropper.set_engine(game)
    observation, _, _ = game.its_showtime()
    shapes.add(tuple(cropper.crop(observation).board.shape))
    backdrops.update(game.backdrop.palette)
    for ch, thing in game.things.items():
      (sprites if isinstance(thing, things.Sprite) else drapes).add(ch)
  if len(shapes) != 1:
    raise ValueError(
        'All pycolab games supplied to the Story constructor should have observations that '
        'are the same shape, either naturally or with the help of observation croppers. The '
        'games provided to the constructor have diverse shapes: {}.'.format(list(shapes)))
  sd, sb, db = sprites & drapes, sprites & backdrops, drapes & backdrops
  if sd or sb or db:
    raise ValueError(
        'No two pycolab games supplied to the Story constructor should use the same '
        'character in two different ways: if a character is a Sprite in one game, it '
        'shouldn\'t be a Drape in another. Across the games supplied to this Story, these '
        'characters are both a Sprite and a Drape: [{}]; these are both a Sprite and in a '
        'Backdrop: [{}]; and these are both a Drape and in a Backdrop: [{}].'.format(
            *[''.join(sorted(s)) for s in (sd, sb, db)]))
  return sprites, drapes, backdrops, shapes.pop()


class _DummySprite(things.Sprite):
  """Invisible, inert Sprite under a character the current chapter does not use."""

  def __init__(self, corner, character):
    super(_DummySprite, self).__init__(corner=corner, position=self.Position(0, 0),
                                       character=character)
    self._visible = False

  def update(self, *args, **kwargs):
    raise RuntimeError('_DummySprite.update should never be called.')


class _DummyDrape(things.Drape):
  """Empty, inert Drape under a character the current chapter does not use."""

  def update(self, *args, **kwargs):
    raise RuntimeError('_DummyDrape.update should never be called.')
