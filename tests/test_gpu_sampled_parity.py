"""Oracle parity AT FULL BATCH for BASELINE.json configs[1..4].

The lockstep tests in test_gpu_parity.py stop at B = 8..24; the full-size tests
in test_gpu_full_size.py use invariants only.  Here a full-size BatchedEngine
(4096 / 8192 / 16384 / 8192 envs, per-env levels, per-env random actions,
auto-reset) is stepped >= 200 steps while 64 randomly chosen env indices are
replayed by the oracle (reference engine.py:583-639 restated) and compared bit
for bit on every step: board u8, reward incl. None-ness, discount, game_over.
"""

import numpy as np
import pytest

from oracle import engine_model as em
from oracle import games as ogames
from oracle import sampled_check

pytestmark = pytest.mark.gpu

SAMPLE = 64
STEPS = 200


def _sample(B, seed):
  rs = np.random.RandomState(seed)
  ids = set(int(i) for i in rs.choice(B, size=SAMPLE - 4, replace=False))
  ids.update([0, 1, B - 2, B - 1])             # the edges of the batch too
  return sorted(ids)


def test_scrolly_64x64_batch_4096_sampled_vs_oracle():
  """configs[1]: the bench.py workload (32 generated levels, actions 0-4)."""
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import scrolly_maze as g
  arts = [levels.scrolly_maze_level(1000 + i) for i in range(32)]
  games = [lowering.lower(g.make_game(*a)) for a in arts]
  B = 4096
  eng = batched.BatchedEngine(games, batch=B)
  eng.its_showtime()
  rs = np.random.RandomState(1234)
  actions = rs.randint(0, 5, size=(STEPS, B)).astype(np.int32)
  actions[rs.random_sample(actions.shape) < 0.002] = 5      # a few quits: auto-reset path
  make = lambda e: ogames.make_scrolly_maze(arts[e % 32][0], arts[e % 32][1], '+',
                                            arts[e % 32][2])
  n = sampled_check.lockstep(eng, make, _sample(B, 1), actions)
  assert n == SAMPLE * (STEPS + 1)
  assert int(eng.error_codes().abs().max()) == 0


def test_warehouse_80x80_batch_8192_sampled_vs_oracle():
  """configs[2]."""
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import warehouse_manager as g
  arts = [levels.warehouse_level(100 + i) for i in range(16)]
  games = [lowering.lower(g.make_game(a)) for a in arts]
  B = 8192
  eng = batched.BatchedEngine(games, batch=B)
  eng.its_showtime()
  rs = np.random.RandomState(77)
  actions = rs.randint(0, 4, size=(STEPS, B)).astype(np.int32)
  actions[rs.random_sample(actions.shape) < 0.002] = 5
  make = lambda e: ogames.make_warehouse(arts[e % 16])
  n = sampled_check.lockstep(eng, make, _sample(B, 2), actions)
  assert n == SAMPLE * (STEPS + 1)
  assert int(eng.error_codes().abs().max()) == 0


def test_marauders_batch_16384_sampled_vs_oracle():
  """configs[3] (all four shards' env indices on one GPU): per-env MT19937
  streams seeded rng_seed + env index, as the sharded run does."""
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import extraterrestrial_marauders as g
  art = levels.marauders_level()
  game = lowering.lower(g.make_game(art))
  B = 16384
  eng = batched.BatchedEngine([game], batch=B, rng_seed=500)
  eng.its_showtime()
  rs = np.random.RandomState(78)
  T = 300                                       # long enough for episodes to end
  actions = rs.randint(0, 4, size=(T, B)).astype(np.int32)
  # One RandomState per env that survives auto-resets, as the device stream does.
  rngs = {}
  def make(e):
    if e not in rngs:
      rngs[e] = np.random.RandomState(500 + e)
    return ogames.make_marauders(art, rngs[e])
  n = sampled_check.lockstep(eng, make, _sample(B, 3), actions)
  assert n == SAMPLE * (T + 1)
  assert int(eng.error_codes().abs().max()) == 0


def test_scrolly_crop_batch_8192_sampled_vs_oracle():
  """configs[4], one GPU's share (65536 / 8): board AND the 9x9 egocentric crop."""
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import scrolly_maze as g
  arts = [levels.scrolly_maze_level(300 + i) for i in range(8)]
  games = [lowering.lower(g.make_game(*a)) for a in arts]
  B = 8192
  eng = batched.BatchedEngine(games, batch=B)
  eng.its_showtime()
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  state = eng.new_crop_state()
  rs = np.random.RandomState(79)
  actions = rs.randint(0, 5, size=(STEPS, B)).astype(np.int32)
  actions[rs.random_sample(actions.shape) < 0.002] = 5
  make = lambda e: ogames.make_scrolly_maze(arts[e % 8][0], arts[e % 8][1], '+', arts[e % 8][2])
  cropper = lambda: em.ScrollingCrop(9, 9, ['P'], pad_char=' ', scroll_margins=(None, None))
  n = sampled_check.lockstep(eng, make, _sample(B, 4), actions, crop=(spec, state, cropper))
  assert n == SAMPLE * (STEPS + 1)
  assert int(eng.error_codes().abs().max()) == 0


def test_run_rotating_and_final_state_check():
  """`pcl_run_many` over three engines in rotation == the oracle's replay of each
  sampled env's own action stream (what bench.py's parity_checked leg does)."""
  import torch
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import scrolly_maze as g
  arts = [levels.scrolly_maze_level(40 + i, world_shape=(65, 65), board_shape=(32, 32))
          for i in range(4)]
  games = [lowering.lower(g.make_game(*a)) for a in arts]
  B, R, K = 96, 3, 90
  engines = [batched.BatchedEngine(games, batch=B, env_offset=r * B) for r in range(R)]
  for e in engines:
    e.its_showtime()
  rs = np.random.RandomState(5)
  acts = rs.randint(0, 5, size=(K, B)).astype(np.int32)
  dev = torch.from_numpy(acts).cuda()
  n0 = sum(e.launch_count() for e in engines)
  batched.run_rotating(engines, [dev[t] for t in range(K)])
  torch.cuda.synchronize()
  assert sum(e.launch_count() for e in engines) - n0 == K
  make = lambda e: ogames.make_scrolly_maze(arts[e % 4][0], arts[e % 4][1], '+', arts[e % 4][2])
  ids = [0, 5, 17, 95]
  for r, eng in enumerate(engines):
    streams = {e: [int(acts[t, e]) for t in range(r, K, R)] for e in ids}
    steps = sampled_check.final_state_check(eng, make, ids, streams, 'Pabc')
    assert steps == len(ids) * (K // R)


def test_play_host_async_pipeline_matches_sync_path():
  """pcl_step_host_async over two engines with slots in flight == pcl_step_host;
  the cropped variant returns the crops only."""
  import torch
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import scrolly_maze as g
  art = levels.scrolly_maze_level(3, world_shape=(65, 65), board_shape=(32, 32))
  game = lowering.lower(g.make_game(*art))
  B = 64
  sync = [batched.BatchedEngine([game], batch=B) for _ in range(2)]
  pipe = [batched.BatchedEngine([game], batch=B) for _ in range(2)]
  cropped = batched.BatchedEngine([game], batch=B)
  for e in sync + pipe + [cropped]:
    e.its_showtime()
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  state = cropped.new_crop_state()
  ref_state = sync[0].new_crop_state()
  rs = np.random.RandomState(11)
  acts = rs.randint(0, 5, size=(40, B)).astype(np.int32)
  want = []
  for t in range(40):
    board, reward, has, disc, done = sync[t % 2].play_host(acts[t])
    want.append((board.copy(), reward.copy(), has.copy(), disc.copy(), done.copy()))
  # pipeline: submit step t, then collect step t - 1 (two slots in flight)
  got = [None] * 40
  for t in range(40):
    pipe[t % 2].play_host_async(acts[t], slot=t % 4)
    if t >= 1:
      got[t - 1] = tuple(x.copy() for x in pipe[(t - 1) % 2].host_wait((t - 1) % 4))
  got[39] = tuple(x.copy() for x in pipe[39 % 2].host_wait(39 % 4))
  for t in range(40):
    for a, b in zip(want[t], got[t]):
      np.testing.assert_array_equal(a, b, err_msg='step %d' % t)
  # cropped view: equals pcl_crop over the synchronous engine's boards
  for t in range(0, 40, 2):
    cropped.play_host_async(acts[t], slot=0, crop_spec=spec, crop_state=state)
    view, reward, _, _, _ = cropped.host_wait(0)
    assert view.shape == (B, 9, 9)
    np.testing.assert_array_equal(reward, want[t][1])
  torch.cuda.synchronize()
  sync0 = batched.BatchedEngine([game], batch=B)
  sync0.its_showtime()
  for t in range(0, 40, 2):
    sync0.play(torch.from_numpy(acts[t]).cuda())
    ref = sync0.crop(spec, state=ref_state)
  np.testing.assert_array_equal(view, ref.cpu().numpy())
