"""Sharding logic: partition arithmetic + world_size-2 gloo all-gather on CPU,
and (gpu) shard-vs-single-engine equivalence on one device."""

import os
import socket

import numpy as np
import pytest

from pycolab_b200 import dist as pdist


def test_shard_range_partitions_exactly():
  for total in (1, 7, 8, 4096, 65536, 65537):
    for world in (1, 2, 3, 4, 8):
      spans = [pdist.shard_range(total, r, world) for r in range(world)]
      assert spans[0][0] == 0
      for (f0, c0), (f1, _) in zip(spans, spans[1:]):
        assert f0 + c0 == f1
      assert spans[-1][0] + spans[-1][1] == total
      counts = [c for _, c in spans]
      assert max(counts) - min(counts) <= 1
  with pytest.raises(ValueError):
    pdist.shard_range(8, 2, 2)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, total, ok):
  import torch
  import torch.distributed as dist
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  first, count = pdist.shard_range(total, rank, world)
  ids = torch.arange(first, first + count, dtype=torch.int32)
  board = (ids[:, None, None] % 251).to(torch.uint8).expand(count, 3, 5).contiguous()
  reward = ids * 7
  boards, rewards = pdist.allgather_outputs([board, reward], total)
  want = torch.arange(total, dtype=torch.int32)
  good = bool((rewards == want * 7).all()) and boards.shape == (total, 3, 5)
  good = good and bool((boards[:, 0, 0] == (want % 251).to(torch.uint8)).all())
  ok[rank] = 1 if good else 0
  dist.destroy_process_group()


@pytest.mark.parametrize('total', [10, 11])
def test_allgather_outputs_gloo_world2(total):
  import torch.multiprocessing as mp
  world = 2
  ok = mp.get_context('spawn').Array('i', [0] * world)
  mp.spawn(_worker, args=(world, _free_port(), total, ok), nprocs=world, join=True)
  assert list(ok) == [1] * world


def pack_records(view, reward, discount, done, has_reward):
  """NumPy restatement of the hand-off record (include/pcl.h:
  PCL_HANDOFF_RECORD_BYTES) — test infrastructure."""
  n = view.shape[0]
  flat = view.reshape(n, -1)
  padded = (flat.shape[1] + 3) & ~3
  rec = np.zeros((n, padded + 12), dtype=np.uint8)
  rec[:, :flat.shape[1]] = flat
  rec[:, padded:padded + 4] = reward.astype('<i4').view(np.uint8).reshape(n, 4)
  rec[:, padded + 4:padded + 8] = discount.astype('<f4').view(np.uint8).reshape(n, 4)
  rec[:, padded + 8], rec[:, padded + 9] = done, has_reward
  return rec


def _shard_outputs(first, count):
  ids = np.arange(first, first + count)
  view = ((ids[:, None, None] * 3 + np.arange(15).reshape(3, 5)) % 251).astype(np.uint8)
  return (view, (ids * 7 - 3).astype(np.int32), (ids % 2).astype(np.float32),
          (ids % 3 == 0).astype(np.uint8), (ids % 5 != 0).astype(np.uint8))


def test_unpack_handoff_views():
  import torch
  outs = _shard_outputs(4, 9)
  rec = torch.from_numpy(pack_records(*outs))
  assert rec.shape[1] == pdist.handoff_record_bytes(15) == 28
  for got, want in zip(pdist.unpack_handoff(rec, (3, 5)), outs):
    np.testing.assert_array_equal(got.numpy(), want)


def _handoff_worker(rank, world, port, total, ok):
  import torch
  import torch.distributed as dist
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  first, count = pdist.shard_range(total, rank, world)

  class Shard(object):            # what Handoff needs of a BatchedEngine
    batch, device = count, torch.device('cpu')
  handoff = pdist.Handoff(Shard(), (3, 5), total)
  handoff.packed[:count] = torch.from_numpy(pack_records(*_shard_outputs(first, count)))
  got = handoff.exchange()        # the ONE all-gather + unpack
  good = all(np.array_equal(g.numpy(), w) for g, w in zip(got, _shard_outputs(0, total)))
  ok[rank] = 1 if good else 0
  dist.destroy_process_group()


@pytest.mark.parametrize('total', [10, 11])
def test_handoff_single_collective_gloo_world2(total):
  import torch.multiprocessing as mp
  world = 2
  ok = mp.get_context('spawn').Array('i', [0] * world)
  mp.spawn(_handoff_worker, args=(world, _free_port(), total, ok), nprocs=world, join=True)
  assert list(ok) == [1] * world


def _peer_worker(rank, world, port, ok):
  import torch
  import torch.distributed as dist
  from pycolab_b200 import batched, levels
  from pycolab_b200.games import scrolly_maze
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
  total, good = 2 * 33 + 1, True               # uneven shards: 34 + 33
  arts = [levels.scrolly_maze_level(80 + i, world_shape=(65, 65), board_shape=(30, 45))
          for i in range(3)]
  eng = pdist.make_shard_engine([scrolly_maze.make_game(*a) for a in arts], total, rank, world,
                                rank)
  eng.its_showtime()
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  nccl, peer = pdist.Handoff(eng, (9, 9), total), pdist.PeerHandoff(eng, (9, 9), total)
  # the one-kernel path (crop + pack + stores + flag barrier), unicast and multicast
  fused = [pdist.FusedHandoff(eng, spec, total, multicast=False, signal_kernel=False),
           pdist.FusedHandoff(eng, spec, total, multicast=True, signal_kernel=False),
           pdist.FusedHandoff(eng, spec, total, multicast=True)]   # + the publish kernel
  # split phase: signal this step, wait for the previous one, read one step behind
  lagged = pdist.FusedHandoff(eng, spec, total, lag=1)
  rs = np.random.RandomState(rank)
  previous = None
  for t in range(25):
    eng.play(torch.from_numpy(rs.randint(0, 5, size=eng.batch).astype(np.int32)).cuda())
    crop = eng.crop(spec).clone()
    want = [t.clone() for t in nccl.gather(crop)]
    got = peer.gather(crop)
    torch.cuda.synchronize()
    good = good and all(bool((g == w).all()) for g, w in zip(got, want))
    good = good and got[0].shape == (total, 9, 9)
    for f in fused:                            # each owns its cropper state: same windows
      got = f.gather()
      torch.cuda.synchronize()
      good = good and all(bool((g == w).all()) for g, w in zip(got, want))
    got = lagged.gather()                      # the PREVIOUS step's records, complete
    torch.cuda.synchronize()
    good = good and ((got is None) == (t == 0))
    if got is not None:
      good = good and all(bool((g == w).all()) for g, w in zip(got, previous))
    previous = want
  got = lagged.flush()                         # ... and the last step after a host barrier
  good = good and all(bool((g == w).all()) for g, w in zip(got, previous))
  ok[rank] = 1 if good else 0
  dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_handoff_equals_nccl_handoff_on_two_gpus():
  """The all-gather fused into the pack kernel (P2P stores into symmetric memory)
  delivers what pack + NCCL all-gather delivers.  Needs two GPUs."""
  import torch
  import torch.multiprocessing as mp
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  ok = mp.get_context('spawn').Array('i', [0, 0])
  mp.spawn(_peer_worker, args=(2, _free_port(), ok), nprocs=2, join=True)
  assert list(ok) == [1, 1]


@pytest.mark.gpu
def test_pack_handoff_kernel_matches_the_record_layout():
  import torch
  from pycolab_b200 import batched, levels
  from pycolab_b200.games import scrolly_maze
  arts = [levels.scrolly_maze_level(70 + i, world_shape=(65, 65), board_shape=(30, 45))
          for i in range(2)]
  eng = batched.BatchedEngine([scrolly_maze.make_game(*a) for a in arts], batch=37)
  eng.its_showtime()
  rs = np.random.RandomState(1)
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  packed = torch.zeros((40, pdist.handoff_record_bytes(81)), dtype=torch.uint8, device='cuda')
  for _ in range(40):
    res = eng.play(torch.from_numpy(rs.randint(0, 5, size=37).astype(np.int32)).cuda())
    crop = eng.crop(spec)
    eng.pack_handoff(crop, packed)
    torch.cuda.synchronize()
    want = pack_records(crop.cpu().numpy(), res.reward.cpu().numpy(), res.discount.cpu().numpy(),
                        res.done.cpu().numpy(), res.has_reward.cpu().numpy())
    np.testing.assert_array_equal(packed[:37].cpu().numpy(), want)
    assert not packed[37:].any()


@pytest.mark.gpu
def test_fused_crop_handoff_kernel_single_rank():
  """`pcl_crop_handoff` with one rank = crop_kernel + pack_handoff_kernel: same
  windows (own cropper state), same records, alternating halves, CUDA-graph
  replayable because the step counter lives on the device."""
  import torch
  from pycolab_b200 import batched, levels
  from pycolab_b200.games import scrolly_maze
  arts = [levels.scrolly_maze_level(70 + i, world_shape=(65, 65), board_shape=(30, 45))
          for i in range(2)]
  eng = batched.BatchedEngine([scrolly_maze.make_game(*a) for a in arts], batch=37)
  eng.its_showtime()
  rs = np.random.RandomState(1)
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  state = eng.new_crop_state()
  fused = pdist.FusedHandoff(eng, spec, 37, signal_kernel=False)
  assert fused.rec == 96 and fused.transport.startswith('local')
  for t in range(41):
    res = eng.play(torch.from_numpy(rs.randint(0, 5, size=37).astype(np.int32)).cuda())
    crop = eng.crop(spec, state=state)
    view, reward, discount, done, has = fused.gather()
    torch.cuda.synchronize()
    assert bool((view == crop).all()), t
    assert bool((reward == res.reward).all()) and bool((discount == res.discount).all())
    assert bool((done == res.done).all()) and bool((has == res.has_reward).all())
    assert int(fused.local[0]) == t + 1 and int(fused.local[1]) == 0
  # an odd view size whose record needs zero padding up to a multiple of 16
  spec2 = batched.scrolling_crop_spec(5, 7, 0, pad_char='.', scroll_margins=(1, 2))
  fused2 = pdist.FusedHandoff(eng, spec2, 37)
  state2 = eng.new_crop_state()
  assert fused2.rec == 48
  for t in range(6):
    eng.play(torch.from_numpy(rs.randint(0, 5, size=37).astype(np.int32)).cuda())
    crop = eng.crop(spec2, state=state2)
    view = fused2.gather()[0]
    torch.cuda.synchronize()
    assert bool((view == crop).all()), t
    assert not fused2.halves[t & 1][:, 35 + 1 + 12:].any()    # padding stays zero


@pytest.mark.gpu
def test_fused_crop_handoff_split_phase_single_rank():
  """lag=1 (three buffer parts, wait for the previous step): gather() hands back step
  t - 1, flush() the last step."""
  import torch
  from pycolab_b200 import batched, levels
  from pycolab_b200.games import scrolly_maze
  art = levels.scrolly_maze_level(71, world_shape=(65, 65), board_shape=(30, 45))
  eng = batched.BatchedEngine([scrolly_maze.make_game(*art)], batch=21)
  eng.its_showtime()
  rs = np.random.RandomState(2)
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  state = eng.new_crop_state()
  fused = pdist.FusedHandoff(eng, spec, 21, lag=1)
  assert fused.n_bufs == 3
  prev = None
  for t in range(10):
    res = eng.play(torch.from_numpy(rs.randint(0, 5, size=21).astype(np.int32)).cuda())
    now = (eng.crop(spec, state=state).clone(), res.reward.clone(), res.discount.clone(),
           res.done.clone(), res.has_reward.clone())
    got = fused.gather()
    torch.cuda.synchronize()
    assert (got is None) == (t == 0)
    if got is not None:
      assert all(bool((g == w).all()) for g, w in zip(got, prev)), t
    prev = now
  assert all(bool((g == w).all()) for g, w in zip(fused.flush(), prev))


@pytest.mark.gpu
def test_two_shards_reproduce_one_engine():
  """Global env i behaves identically whether it lives in a 1-rank or a 2-rank job
  (levels and per-env RNG streams are keyed by GLOBAL env index)."""
  import torch
  from pycolab_b200 import batched, levels
  from pycolab_b200.games import extraterrestrial_marauders as marauders
  art = levels.marauders_level()
  total, T = 24, 150
  games = [marauders.make_game(art)]
  whole = batched.BatchedEngine(games, batch=total, rng_seed=5)
  shards = [pdist.make_shard_engine(games, total, r, 2, device=0, rng_seed=5)
            for r in range(2)]
  whole.its_showtime()
  for s in shards:
    s.its_showtime()
  rs = np.random.RandomState(2)
  for t in range(T):
    a = torch.from_numpy(rs.randint(0, 4, size=total).astype(np.int32)).cuda()
    rw = whole.play(a)
    first = 0
    for s in shards:
      r = s.play(a[first:first + s.batch].contiguous())
      assert bool((r.board == rw.board[first:first + s.batch]).all()), t
      assert bool((r.reward == rw.reward[first:first + s.batch]).all()), t
      assert bool((r.done == rw.done[first:first + s.batch]).all()), t
      first += s.batch
