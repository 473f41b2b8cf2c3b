#!/usr/bin/env python
"""Why is the pipelined host path slower than the blocking one?  Times variants of
the C2 end-to-end step (4096 envs, 64x64 boards to pinned host memory)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
  import torch
  from pycolab_b200 import batched, lowering
  from pycolab_b200.games import scrolly_maze
  dev = torch.device('cuda', 0)
  arts = bench.make_levels(8)
  games = [lowering.lower(scrolly_maze.make_game(*a)) for a in arts]
  B, R = 4096, 6
  engines = [batched.BatchedEngine(games, batch=B, env_offset=r * B) for r in range(R)]
  for e in engines:
    e.its_showtime()
  acts = np.random.RandomState(0).randint(0, 5, size=(64, B)).astype(np.int32)
  n = 60

  def timeit(name, fn):
    for t in range(12):
      fn(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(n):
      fn(t)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%-46s %.3f ms/step  %.1f M env-steps/s' % (name, 1000 * dt / n, B * n / dt / 1e6))

  timeit('pcl_step_host (blocking)', lambda t: engines[t % R].play_host(acts[t % 64]))

  def depth1(t):
    e = engines[t % R]
    e.play_host_async(acts[t % 64], slot=0)
    e.host_wait(0)
  timeit('async, wait at once (depth 1)', depth1)

  state = {'prev': None}
  def depth2(t):
    e = engines[t % R]
    e.play_host_async(acts[t % 64], slot=(t // R) % 2)
    if state['prev'] is not None:
      state['prev'][0].host_wait(state['prev'][1])
    state['prev'] = (e, (t // R) % 2)
  timeit('async, collect previous (depth 2)', depth2)
  state['prev'][0].host_wait(state['prev'][1])
  state['prev'] = None

  # torch-level version of the same pipeline for comparison
  pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
  hb = [pin((B, 64, 64), torch.uint8) for _ in range(2)]
  copy_stream = torch.cuda.Stream(dev)
  evs = [torch.cuda.Event() for _ in range(2)]
  d_act = torch.zeros(B, dtype=torch.int32, device=dev)
  h_act = pin((B,), torch.int32)
  def torch_pipe(t):
    e = engines[t % R]
    h_act.numpy()[:] = acts[t % 64]
    d_act.copy_(h_act, non_blocking=True)
    e.play(d_act)
    k = t & 1
    ready = torch.cuda.Event()
    ready.record()
    with torch.cuda.stream(copy_stream):
      copy_stream.wait_event(ready)
      hb[k].copy_(e._board, non_blocking=True)
      evs[k].record()
    evs[k ^ 1].synchronize()
  timeit('torch: play + D2H on a side stream (depth 2)', torch_pipe)
  def torch_sync(t):
    e = engines[t % R]
    h_act.numpy()[:] = acts[t % 64]
    d_act.copy_(h_act, non_blocking=True)
    e.play(d_act)
    hb[0].copy_(e._board, non_blocking=True)
    torch.cuda.synchronize()
  timeit('torch: play + D2H same stream + sync', torch_sync)
  def d2h_only(t):
    hb[0].copy_(engines[t % R]._board, non_blocking=True)
    torch.cuda.synchronize()
  timeit('D2H of one board batch only', d2h_only)


if __name__ == '__main__':
  main()
