"""BASELINE.json's full configuration sizes on the GPU, checked through
size-independent properties (no oracle run at these sizes): lockstep
determinism, sharding invariance, conservation laws of each game, legal
characters only, crop consistency."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
  import torch
  return torch


def _checksum(t):
  torch = _torch()
  w = torch.arange(1, t.numel() + 1, device=t.device, dtype=torch.int64) % 1000003
  return int((t.reshape(-1).long() * w).sum())


def test_warehouse_80x80_batch_8192_invariants():
  """configs[2]."""
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import warehouse_manager as g
  torch = _torch()
  games = [lowering.lower(g.make_game(levels.warehouse_level(200 + i))) for i in range(8)]
  B = 8192
  a = batched.BatchedEngine(games, batch=B)
  b = batched.BatchedEngine(games, batch=B)
  a.its_showtime()
  b.its_showtime()
  rs = np.random.RandomState(0)
  total = torch.zeros(B, dtype=torch.int64, device='cuda')
  for t in range(40):
    act = torch.from_numpy(rs.randint(0, 4, size=B).astype(np.int32)).cuda()
    ra = a.play(act)
    rb = b.play(act)
    total += ra.reward.long()
    assert bool((ra.board == rb.board).all())            # determinism
  board = a.board
  # ten boxes + the player are always on the board: digits + 'X' == 10, one 'P'
  digits = ((board >= ord('0')) & (board <= ord('9'))).sum(dim=(1, 2))
  marks = (board == ord('X')).sum(dim=(1, 2))
  assert bool(((digits + marks) == 10).all())
  assert bool(((board == ord('P')).sum(dim=(1, 2)) == 1).all())
  # the summed reward is the number of boxes on goals now (it starts at 0)
  done_now = a.done.bool()
  assert bool((total[~done_now] == marks[~done_now]).all())
  assert bool(a.has_reward.bool().all())
  assert int(a.error_codes().abs().max()) == 0
  assert _checksum(a.board) == _checksum(b.board)


def test_marauders_batch_16384_invariants():
  """configs[3] on one GPU (the 4-GPU run shards these env indices)."""
  from pycolab_b200 import batched, dist, levels, lowering
  from pycolab_b200.games import extraterrestrial_marauders as g
  torch = _torch()
  game = lowering.lower(g.make_game(levels.marauders_level()))
  B = 16384
  whole = batched.BatchedEngine([game], batch=B, rng_seed=11)
  first, count = dist.shard_range(B, 2, 4)             # the third of four shards
  shard = dist.make_shard_engine([game], B, 2, 4, device=0, rng_seed=11)
  assert shard.batch == count == 4096
  whole.its_showtime()
  shard.its_showtime()
  rs = np.random.RandomState(1)
  for t in range(60):
    act = torch.from_numpy(rs.randint(0, 4, size=B).astype(np.int32)).cuda()
    rw = whole.play(act)
    r = shard.play(act[first:first + count].contiguous())
    assert bool((r.board == rw.board[first:first + count]).all()), t   # sharding invariance
    assert bool((r.reward == rw.reward[first:first + count]).all()), t
  board = whole.board
  legal = torch.tensor([ord(c) for c in ' PBXabcdyz'], device='cuda', dtype=torch.uint8)
  assert bool(torch.isin(board, legal).all())
  assert bool(((board == ord('P')).sum(dim=(1, 2)) <= 1).all())
  assert bool(((board == ord('X')).sum(dim=(1, 2)) <= 40).all())
  # rewards: -1 per bunker hit, +10 per marauder: never below -6 in one step
  assert int(whole.reward.min()) >= -6 and int(whole.reward.max()) <= 40
  assert int(whole.error_codes().abs().max()) == 0


def test_scrolly_crop_batch_8192_invariants():
  """configs[4] per-GPU share: 8192 envs, 64x64 board, 9x9 egocentric crop."""
  from pycolab_b200 import batched, levels, lowering
  from pycolab_b200.games import scrolly_maze as g
  torch = _torch()
  games = [lowering.lower(g.make_game(*levels.scrolly_maze_level(300 + i))) for i in range(8)]
  B = 8192
  eng = batched.BatchedEngine(games, batch=B)
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  state = eng.new_crop_state()
  eng.its_showtime()
  rs = np.random.RandomState(2)
  idx = torch.arange(B, device='cuda')
  for t in range(30):
    res = eng.play(torch.from_numpy(rs.randint(0, 5, size=B).astype(np.int32)).cuda())
    crop = eng.crop(spec, state=state)
    # egocentric: the centre cell is the player whenever the player is visible
    rec = eng.sprites[:, 0]
    visible = (rec[:, 4] & 1).bool()
    assert bool((crop[visible, 4, 4] == ord('P')).all()), t
    # the crop equals the window of the board around the player (pad ' ')
    r, c = rec[:, 0].long(), rec[:, 1].long()
    for dr, dc in ((0, 0), (-4, -4), (4, 4), (-3, 2)):
      rr, cc = r + dr, c + dc
      inside = visible & (rr >= 0) & (rr < 64) & (cc >= 0) & (cc < 64)
      want = res.board[idx[inside], rr[inside], cc[inside]]
      assert bool((crop[inside, 4 + dr, 4 + dc] == want).all()), (t, dr, dc)
      outside = visible & ~inside
      assert bool((crop[outside, 4 + dr, 4 + dc] == ord(' ')).all())
  assert bool((eng.reward % 100 == 0).all())
  assert int(eng.error_codes().abs().max()) == 0
