"""`storytelling.Story` (host-side caller of the path) on CPU.

The Story logic is driven here by oracle Worlds dressed as Engines, against the
golden trajectory the reference's own Story produced for the same chapters
(tests/golden/story_classics_list.npz); plus the constructor's argument checks
(storytelling.py:493-622).
"""

import importlib

import numpy as np
import pytest

import golden_cases as gc
import story_cases
import trajectory as tj
from oracle import games as ogames
from pycolab_b200 import engine as engine_lib
from pycolab_b200 import plot as plot_lib
from pycolab_b200 import rendering
from pycolab_b200 import storytelling
from pycolab_b200 import things


class _Walker(things.Sprite):
  def update(self, *args, **kwargs):
    raise AssertionError('never called')


class OracleEngine(object):
  """An oracle World with the attributes Story reads from an Engine."""

  def __init__(self, world, palette):
    self._world = world
    self.the_plot = plot_lib.Plot()
    self.rows, self.cols = world.rows, world.cols
    self._palette = palette

  def _obs(self, out):
    board = np.asarray(out[0], dtype=np.uint8)
    chars = set(self._palette) | set(self._world.things)
    return rendering.Observation(board=board, layers=rendering.LazyLayers(board, chars)), out[1], out[2]

  def its_showtime(self):
    return self._obs(self._world.its_showtime())

  def play(self, actions):
    return self._obs(self._world.play(actions))

  @property
  def game_over(self):
    return self._world.game_over

  @property
  def z_order(self):
    return list(self._world.things)

  @property
  def backdrop(self):
    return things.Backdrop(curtain=self._world.backdrop, palette=engine_lib.Palette(self._palette))

  @property
  def things(self):
    return {ch: _Walker(things.Sprite.Position(self.rows, self.cols),
                        things.Sprite.Position(w.row, w.col), ch)
            for ch, w in self._world.things.items()}


def _oracle_chapter(kind, art):
  stock = importlib.import_module('pycolab_b200.games.classics.' + kind).GAME_ART
  palette = ' #' if kind == 'four_rooms' else '.'
  return lambda: OracleEngine(ogames.make_classic(kind, art or list(stock)), palette)


def test_list_story_matches_reference_golden():
  g = gc.load('story_classics_list')
  chapters = []
  make = lambda: storytelling.Story([_oracle_chapter(k, a) for k, a in story_cases.LIST_CHAPTERS])
  got = tj.run_trajectory(make, g['actions'].tolist(),
                          on_frame=lambda env, out: chapters.append(str(env.the_plot.this_chapter)))
  tj.assert_same_trajectory(g, got, 'story_classics_list')
  assert chapters == g['chapters'].tolist()


def test_story_views_and_plot_hand_over():
  story = storytelling.Story([_oracle_chapter(k, a) for k, a in story_cases.LIST_CHAPTERS])
  story.its_showtime()
  assert (story.rows, story.cols) == (4, 12)
  assert story.z_order == ['P'] and set(story.backdrop.palette) == {'.'}
  assert not storytelling.is_fictional(story.things['P'])
  story.the_plot['note'] = 42
  first = story.current_game
  # one step east falls off the cliff: chapter 0 ends, chapter 1 starts in the same call
  obs, reward, discount = story.play(3)
  assert first.game_over and story.current_game is not first and not story.game_over
  assert story.the_plot['note'] == 42                       # Plot entries travel
  assert (story.the_plot.prior_chapter, story.the_plot.this_chapter,
          story.the_plot.next_chapter) == (0, 1, 2)
  assert reward == -100.0 and discount == 1.0               # new game's discount, old reward
  with pytest.raises(RuntimeError):
    story.its_showtime()


def test_story_ends_after_last_chapter_and_refuses_more_play():
  story = storytelling.Story([_oracle_chapter('chain_walk', None)])
  story.its_showtime()
  for _ in range(2):
    obs, reward, discount = story.play(0)
  assert story.game_over and reward == 1.0 and discount == 0.0
  with pytest.raises(RuntimeError):
    story.play(0)


def test_dict_story_follows_next_chapter_and_rejects_unknown_keys():
  def cliff_then(target):
    def build():
      game = _oracle_chapter('cliff_walk', None)()
      game.the_plot.next_chapter = target
      return game
    return build
  story = storytelling.Story({'a': cliff_then('b'), 'b': cliff_then(None)}, first_chapter='a')
  story.its_showtime()
  story.play(3)
  assert story.the_plot.this_chapter == 'b' and story.the_plot.prior_chapter == 'a'
  bad = storytelling.Story({'a': cliff_then('nowhere')}, first_chapter='a')
  bad.its_showtime()
  with pytest.raises(KeyError):
    bad.play(3)


def test_constructor_argument_checks():
  rooms, cliff = _oracle_chapter('four_rooms', None), _oracle_chapter('cliff_walk', None)
  with pytest.raises(ValueError):
    storytelling.Story([])
  with pytest.raises(ValueError):
    storytelling.Story({None: cliff}, first_chapter=None)
  with pytest.raises(ValueError):
    storytelling.Story({'a': cliff}, first_chapter='b')
  with pytest.raises(ValueError):
    storytelling.Story([cliff, cliff], croppers=[None])          # keys differ
  with pytest.raises(ValueError):
    storytelling.Story([rooms, cliff])                           # 13x13 vs 4x12 observations

  class _DrapeP(OracleEngine):                                   # 'P' as a Drape elsewhere
    @property
    def things(self):
      class D(things.Drape):
        def update(self, *args, **kwargs):
          pass
      return {'P': D(np.zeros((self.rows, self.cols), dtype=bool), 'P')}
  clash = lambda: _DrapeP(ogames.make_classic('cliff_walk', ['............'] * 3 + ['P...........']), '.')
  with pytest.raises(ValueError):
    storytelling.Story([cliff, clash])


def test_reference_story_tests_pass_against_this_story_class(monkeypatch):
  """The reference's own `tests/story_test.py` (6 tests: sequences, dicts,
  cropping, inter-game reward accumulation, stand-ins, compatibility checking),
  UNMODIFIED, with `storytelling.Story` replaced by this package's class.  The
  chapters are reference Engines (the tests define entities with Python update
  logic), so the class is pointed at the reference's `things` / `cropping` /
  `engine` types for the duration."""
  import sys
  import unittest
  import refdriver
  if not refdriver.available():
    pytest.skip('/root/reference not present')
  refdriver.ref_storytelling()          # the Python 3.12 collections shim + import path
  from pycolab import cropping as ref_cropping
  from pycolab import engine as ref_engine
  from pycolab import things as ref_things
  from pycolab.tests import story_test
  monkeypatch.setattr(storytelling, 'cropping', ref_cropping)
  monkeypatch.setattr(storytelling, 'things', ref_things)
  monkeypatch.setattr(storytelling, 'engine_lib', ref_engine)
  monkeypatch.setattr(story_test, 'storytelling', storytelling)
  suite = unittest.defaultTestLoader.loadTestsFromModule(story_test)
  result = unittest.TextTestRunner(verbosity=0).run(suite)
  assert result.testsRun == 6 and result.wasSuccessful(), result.failures + result.errors
