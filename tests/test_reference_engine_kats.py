"""The known answers of the reference's `tests/engine_test.py`, replayed on the
oracle (CPU).  Captured by tests/golden/make_engine_kats.py:

* per frame, the hand-drawn board every entity was PROMISED to see when its
  update ran (the staged renders between update groups, engine.py:725-735) and
  the hand-drawn board asserted on the observation `play()` returned;
* the Plot directives the test's injected callables issued (add_reward — with
  string rewards —, terminate_episode with default and custom discounts,
  change_z_order), and the reward / discount / game_over / z-order the reference
  returned for them;
* every layer of every observation, with and without occlusion
  (rendering.py:98-179, 187-301);
* every observation post-processor call of the test (repainter, to-array,
  to-feature-array, with dtypes and permutations).
"""

import json
import os

import numpy as np
import pytest

import reference_kats as rk
from oracle import engine_model as em
from oracle import games

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                    'reference_engine_kats.json')
with open(PATH) as f:
  DATA = json.load(f)


def world_of(snap, directives_of_frame):
  world = rk.oracle_world(dict(snap, scrollys={}, scrolling={}))
  last = snap['groups'][-1][-1]

  def program(w, ch, actions):
    if not isinstance(w.things[ch], rk.InertSprite):
      games.fixture_program(w, ch, actions)
    if ch == last:                       # the frame's Plot directives, in call order
      for name, args, kwargs in directives_of_frame():
        getattr(w.plot, name)(*args, **kwargs)
  world.program = program
  world._render()
  return world


def _ids():
  seen, out = {}, []
  for e in DATA['engines']:
    n = seen[e['test']] = seen.get(e['test'], 0) + 1
    out.append('%s-%d' % (e['test'].split('.')[-1], n))
  return out


@pytest.mark.parametrize('kat', DATA['engines'], ids=_ids())
def test_oracle_reproduces_engine_test(kat):
  current = {}
  world = world_of(kat['snapshot'], lambda: current['directives'])
  chars = list(world.things)
  group_of = {ch: g for g, group in enumerate(kat['snapshot']['groups']) for ch in group}
  for i, frame in enumerate(kat['frames']):
    where = '%s frame %d' % (kat['test'], i)
    current['directives'] = frame['directives']
    before = world.board
    board, reward, discount = world.play(
        None if frame['action'] is None else rk.motion_of(frame['action'], chars))
    # what each entity was promised to see: the render that preceded its group
    for ch, art in frame['expect_seen'].items():
      g = group_of[ch]
      seen = before if g == 0 else world.staged[g - 1]
      np.testing.assert_array_equal(seen, rk.u8(art), err_msg='%s: board seen by %s' % (where, ch))
    if frame['expect_final'] is not None:
      np.testing.assert_array_equal(board, rk.u8(frame['expect_final']), err_msg=where)
    np.testing.assert_array_equal(board, rk.u8(frame['board']), err_msg=where)
    assert reward == frame['reward'] and type(reward) is type(frame['reward']), where
    assert discount == frame['discount'] and world.game_over == frame['game_over'], where
    assert world.z_order == frame['z_order'], where
    if kat['occlusion_in_layers']:
      layers = em.layers_of(board, list(frame['layers']))
    else:
      layers = em.unoccluded_layers_of(world.backdrop, world.things, list(frame['layers']))
    for ch, rows in frame['layers'].items():
      np.testing.assert_array_equal(layers[ch], rk.bits(rows), err_msg='%s layer %r' % (where, ch))


@pytest.mark.parametrize('call', DATA['observers'],
                         ids=['%s-%d' % (c['kind'], i) for i, c in enumerate(DATA['observers'])])
def test_oracle_reproduces_observer_calls(call):
  board = rk.u8(call['board'])
  want = np.array(call['out'], dtype=call['out_dtype']).reshape(call['out_shape'])
  args, kwargs = call['args'], call['kwargs']
  if call['kind'] == 'ObservationCharacterRepainter':
    got = em.observation_repaint(board, args[0])
  elif call['kind'] == 'ObservationToArray':
    mapping = {k: (tuple(v) if isinstance(v, list) else v) for k, v in args[0].items()}
    permute = kwargs.get('permute')
    got = em.observation_to_array(board, mapping, np.dtype(kwargs['dtype']) if kwargs.get('dtype')
                                  else None, None if permute is None else tuple(permute))
  else:
    permute = kwargs.get('permute')
    got = em.observation_to_feature_array(board, args[0],
                                          None if permute is None else tuple(permute))
  assert got.shape == want.shape and got.dtype == want.dtype, (got.dtype, want.dtype)
  np.testing.assert_array_equal(got, want)


def test_capture_covers_the_whole_module():
  tests = sorted(set(e['test'].split('.')[-1] for e in DATA['engines']))
  assert tests == ['testChangingZOrdering', 'testOcclusionInLayers', 'testPlotStateVariables',
                   'testRenderingWithOcclusion', 'testRenderingWithoutOcclusion',
                   'testRewardAndEpisodeEndWithCustomDiscount',
                   'testRewardAndEpisodeEndWithDefaultDiscount', 'testUpdateScheduleAndZOrder']
  assert sum(len(f['expect_seen']) for e in DATA['engines'] for f in e['frames']) == 17
  assert len(DATA['observers']) == 14
