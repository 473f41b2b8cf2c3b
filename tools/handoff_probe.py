#!/usr/bin/env python
"""Single-GPU break-down of the per-step hand-off (C5 shape, 4096 envs, 6 rotating
batches, one CUDA graph per variant): step alone, step + crop_kernel, step + the fused
crop + pack + exchange kernel with ONE rank (its stores and flag are local, so what is
left is the kernel's own cost without NVLink latency or cross-rank skew)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
  import torch
  from pycolab_b200 import batched, dist as pdist, lowering
  from pycolab_b200.games import scrolly_maze
  dev = torch.device('cuda', 0)
  arts = bench.make_levels(8)
  games = [lowering.lower(scrolly_maze.make_game(*a)) for a in arts]
  B, R, K = 4096, 6, 240
  engines = [batched.BatchedEngine(games, batch=B, env_offset=r * B) for r in range(R)]
  for e in engines:
    e.its_showtime()
  acts = torch.from_numpy(np.random.RandomState(0).randint(0, 5, size=(K, B)).astype(np.int32)).to(dev)
  spec = batched.scrolling_crop_spec(9, 9, 0, pad_char=' ', scroll_margins=(None, None))
  states = [e.new_crop_state() for e in engines]
  crops = [torch.empty((B, 9, 9), dtype=torch.uint8, device=dev) for _ in engines]
  handoffs = [pdist.FusedHandoff(e, spec, B) for e in engines]

  def step(t):
    engines[t % R].play(acts[t % K])
  def step_crop(t):
    e = engines[t % R]
    e.play(acts[t % K])
    e.crop(spec, state=states[t % R], out=crops[t % R])
  def step_handoff(t):
    engines[t % R].play(acts[t % K])
    handoffs[t % R].gather()
  sync = lambda: torch.cuda.synchronize(dev)
  for name, fn in (('step', step), ('step + crop_kernel', step_crop),
                   ('step + crop_handoff_kernel (1 rank)', step_handoff)):
    for t in range(2 * R):
      fn(t)
    timed = bench.Timed(torch, dev, fn, 0, K)
    timed.run()
    ms = min(timed.time_ms(sync) for _ in range(3)) / K
    print('%-40s %.2f us/step  (%s)' % (name, ms * 1e3, timed.path[:40]))


if __name__ == '__main__':
  main()
