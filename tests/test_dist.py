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
