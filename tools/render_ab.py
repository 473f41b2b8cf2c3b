#!/usr/bin/env python
"""Stand-alone renderer microbenchmarks on the C2 shape: 4096-env launches rotating
over 6 batches (working set > L2), one 24576-env launch, and (`sweep`) a launch-size
sweep.  Round 2 used it to A/B seven kernel variants through a PCL_RENDER_VARIANT
switch in csrc/render.cu (results: profiles/r02_render_ab.txt); only the winner (one
CTA per env) is left in the library, so the variant argument is now just a label.

    python tools/render_ab.py 1          # the renderer as built
    python tools/render_ab.py sweep 1    # us per launch by batch size
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
  import torch
  import bench
  from pycolab_b200 import batched, lowering
  from pycolab_b200.games import scrolly_maze
  arts = bench.make_levels(8)
  games = [lowering.lower(scrolly_maze.make_game(*a)) for a in arts]
  B, R = 4096, 6
  engines = [batched.BatchedEngine(games, batch=B, env_offset=r * B) for r in range(R)]
  acts = torch.randint(0, 5, (24, B), dtype=torch.int32, device='cuda')
  for e in engines:
    e.its_showtime()
    for t in range(24):
      e.play(acts[t])
  peak = bench.hbm_peak()[0] if hasattr(bench, 'hbm_peak') else 6580.3
  best = None
  for rep in range(3):
    ms, big = bench.render_microbench(engines, n=120)
    if best is None or ms < best[0]:
      best = (ms, big)
  ms, big = best
  a = B * (64 * 64 * 4 + 48)
  print(json.dumps({'variant': os.environ.get('PCL_RENDER_VARIANT'),
                    'ctas_per_sm': os.environ.get('PCL_RENDER_CTAS_PER_SM'),
                    'ms_4096': round(ms, 5), 'frac_4096': round(a / (ms / 1e3) / 1e9 / peak, 3),
                    'ms_24576': round(big.get('kernel_ms_mean', 0), 5),
                    'frac_24576': round(6 * a / (big.get('kernel_ms_mean', 1e9) / 1e3) / 1e9 / peak, 3)}))


def sweep():
  """Launch-size sweep on one 24576-env data set: t(B) = a + b * B separates the
  per-launch cost from the streaming rate."""
  import ctypes as C
  import torch
  import bench
  from pycolab_b200 import _lib, batched, lowering
  from pycolab_b200.games import scrolly_maze
  lib = _lib.load()
  arts = bench.make_levels(8)
  games = [lowering.lower(scrolly_maze.make_game(*a)) for a in arts]
  BB = 24576
  eng = batched.BatchedEngine(games, batch=BB)
  eng.its_showtime()
  acts = torch.randint(0, 5, (8, BB), dtype=torch.int32, device='cuda')
  for t in range(8):
    eng.play(acts[t])
  H, pitch, dev = eng.rows, eng.pitch, eng.device
  curtains = torch.zeros((BB, 2, H, pitch), dtype=torch.uint8, device=dev)
  curtains[:, 0, :, :eng.cols] = eng.curtain('#')
  curtains[:, 1, :, :eng.cols] = eng.curtain('@')
  backdrop = eng.backdrop[eng.level.long()].contiguous()
  z = torch.tensor([ord(c) for c in eng.game.z_order], dtype=torch.uint8, device=dev)[None].repeat(BB, 1).contiguous()
  out = torch.zeros((BB, H, pitch), dtype=torch.uint8, device=dev)
  res = []
  for B in (512, 1024, 2048, 4096, 8192, 12288, 24576):
    spec = _lib.Spec()
    spec.abi_version, spec.program = _lib.ABI_VERSION, _lib.PROG_NONE
    spec.rows, spec.cols, spec.pitch = eng.rows, eng.cols, eng.pitch
    spec.n_sprites, spec.n_drapes = len(eng.sprite_chars), len(eng.drape_chars)
    for i, ch in enumerate(eng.sprite_chars):
      spec.sprite_char[i] = ord(ch)
    for i, ch in enumerate(eng.drape_chars):
      spec.drape_char[i] = ord(ch)
    handle = C.c_void_p()
    _lib.check(lib.pcl_create(C.byref(spec), B, dev.index, C.byref(handle)), 'pcl_create')
    parts = BB // B
    def launch(i):
      o = (i % parts) * B
      _lib.check(lib.pcl_render(handle, backdrop[o:].data_ptr(), H * pitch, curtains[o:].data_ptr(),
                                eng.sprites[o:].data_ptr(), z[o:].data_ptr(), out[o:].data_ptr(),
                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'pcl_render')
    n = max(parts * 4, 24)
    timed = bench.Timed(torch, dev, launch, 0, n)
    timed.run()
    ms = min(timed.time_ms(lambda: torch.cuda.synchronize(dev)) for _ in range(3)) / n
    res.append((B, ms))
    lib.pcl_destroy(handle)
  assert bool((out[:, :, :eng.cols] == eng.board).all())
  print(json.dumps({'variant': os.environ.get('PCL_RENDER_VARIANT'),
                    'us_by_B': {b: round(ms * 1e3, 2) for b, ms in res}}))


if __name__ == '__main__':
  if os.environ.get('PCL_RENDER_AB_CHILD') == 'sweep':
    sweep()
  elif os.environ.get('PCL_RENDER_AB_CHILD'):
    child()
  else:
    mode = '1'
    args = sys.argv[1:]
    if args and args[0] == 'sweep':
      mode, args = 'sweep', args[1:]
    for v in (args or ['1', '2', '3', '4']):
      env = dict(os.environ, PCL_RENDER_VARIANT=v.split(':')[0], PCL_RENDER_AB_CHILD=mode)
      if ':' in v:
        env['PCL_RENDER_CTAS_PER_SM'] = v.split(':')[1]
      r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env,
                         capture_output=True, text=True)
      print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:])
      sys.stdout.flush()
