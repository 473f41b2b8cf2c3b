"""The pipelined render kernel path (B >= 1024, boards <= 64x64) vs the oracle."""

import ctypes as C

import numpy as np
import pytest

from oracle import engine_model as em

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape,S,D,B', [((64, 64), 4, 2, 1500), ((10, 30), 4, 2, 2048),
                                         ((16, 39), 7, 2, 1025), ((64, 64), 8, 1, 1184)])
def test_pipelined_render_vs_oracle(shape, S, D, B):
  import torch
  from pycolab_b200 import _lib
  lib = _lib.load()
  H, W = shape
  pitch = (W + 15) // 16 * 16
  rs = np.random.RandomState(B + H)
  schars, dchars = 'ABCDEFGH'[:S], 'st'[:D]
  spec = _lib.Spec()
  spec.abi_version, spec.program = _lib.ABI_VERSION, _lib.PROG_NONE
  spec.rows, spec.cols, spec.pitch, spec.n_sprites, spec.n_drapes = H, W, pitch, S, D
  for i, c in enumerate(schars):
    spec.sprite_char[i] = ord(c)
  for i, c in enumerate(dchars):
    spec.drape_char[i] = ord(c)
  h = C.c_void_p()
  _lib.check(lib.pcl_create(C.byref(spec), B, 0, C.byref(h)), 'pcl_create')
  backdrop = np.zeros((B, H, pitch), np.uint8)
  backdrop[:, :, :W] = rs.choice([32, 46, 35], size=(B, H, W))
  curtains = np.zeros((B, D, H, pitch), np.uint8)
  curtains[:, :, :, :W] = rs.random_sample((B, D, H, W)) < 0.3
  sprites = np.zeros((B, S, 8), np.int32)
  sprites[:, :, 0] = rs.randint(0, H, size=(B, S))
  sprites[:, :, 1] = rs.randint(0, W, size=(B, S))
  sprites[:, :, 4] = rs.randint(0, 2, size=(B, S))
  z = np.stack([rs.permutation([ord(c) for c in schars + dchars]) for _ in range(B)])
  z = z.astype(np.uint8)
  dev = lambda a: torch.from_numpy(a).cuda()
  t_bd, t_cur, t_sp, t_z = dev(backdrop), dev(curtains), dev(sprites), dev(z)
  out = torch.zeros((B, H, pitch), dtype=torch.uint8, device='cuda')
  stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
  _lib.check(lib.pcl_render(h, t_bd.data_ptr(), H * pitch, t_cur.data_ptr(),
                            t_sp.data_ptr(), t_z.data_ptr(), out.data_ptr(), stream),
             'pcl_render')
  torch.cuda.synchronize()
  got = out.cpu().numpy()
  lib.pcl_destroy(h)

  class Ent(object):
    pass
  for b in list(range(0, B, 97)) + [B - 1]:
    things = {}
    for i, c in enumerate(schars):
      e = Ent()
      e.is_sprite, e.row, e.col = True, int(sprites[b, i, 0]), int(sprites[b, i, 1])
      e.visible = bool(sprites[b, i, 4])
      things[c] = e
    for i, c in enumerate(dchars):
      e = Ent()
      e.is_sprite, e.curtain = False, curtains[b, i, :, :W].astype(bool)
      things[c] = e
    want = em.render(H, W, backdrop[b, :, :W], [chr(c) for c in z[b]], things)
    np.testing.assert_array_equal(got[b, :, :W], want, err_msg='env %d' % b)
  assert not got[:, :, W:].any()
