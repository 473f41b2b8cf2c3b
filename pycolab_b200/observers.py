"""Observation post-processors on the device (reference `pycolab/rendering.py:304-661`).

`ObservationCharacterRepainter`, `ObservationToArray` and
`ObservationToFeatureArray` are all "a function of the character, per cell", so
one kernel (`pcl_observe`, csrc/observe.cu) serves the three through a
[128, depth] table.  This module builds the tables from the reference's
constructor arguments and launches the kernel over a batch of boards; the
reference-named single-observation classes live in `rendering.py`.
"""

import ctypes as C

import numpy as np

from pycolab_b200 import _lib

_DTYPES = {np.dtype(np.uint8): 0, np.dtype(np.int32): 1, np.dtype(np.float32): 2,
           np.dtype(np.int64): 3, np.dtype(np.float64): 4}


def value_table(value_mapping, dtype=None):
  """ObservationToArray's mapping (rendering.py:423-470) as (table [128, depth],
  valid u8 [128], is_3d)."""
  first = next(iter(value_mapping.values()))
  dt = np.dtype(dtype) if dtype is not None else np.array(first).dtype
  try:
    depth, is_3d = len(first), True
  except TypeError:
    depth, is_3d = 1, False
  if dt not in _DTYPES:
    raise TypeError('ObservationToArray on the device supports uint8, int32, int64, '
                    'float32 and float64 outputs, not {}'.format(dt))
  table = np.zeros((128, depth), dtype=dt)
  valid = np.zeros((128,), dtype=np.uint8)
  for ch, value in value_mapping.items():
    code = ord(ch)
    if code > 127:
      raise ValueError('non-ASCII character {!r} in a value mapping'.format(ch))
    table[code] = value
    valid[code] = 1
  return table, valid, is_3d


def feature_table(layers):
  """ObservationToFeatureArray (rendering.py:545-661): one-hot float32 planes."""
  table = np.zeros((128, len(layers)), dtype=np.float32)
  for d, ch in enumerate(layers):
    table[ord(ch), d] = 1.0
  return table


def repaint_table(character_mapping):
  """ObservationCharacterRepainter (rendering.py:304-357): identity LUT with the
  mapped characters replaced."""
  table = np.arange(128, dtype=np.uint8).reshape(128, 1).copy()
  for src, dst in character_mapping.items():
    table[ord(src), 0] = ord(dst)
  return table


def check_permute(permute, is_3d, who):
  if permute is None:
    return None
  permute = tuple(permute)
  want = [0, 1, 2] if is_3d else [0, 1]
  if sorted(permute) != want:
    raise ValueError(
        'The permute argument to the {} constructor must be a list or tuple containing '
        'some permutation of the integers {}.'.format(who, ', '.join(map(str, want))))
  return permute


def observe(lib, handle, board, rows, cols, table, valid, is_3d, permute, stream,
            unknown=None):
  """Run pcl_observe over `board` (u8 [B, rows, pitch] CUDA tensor).  Returns a
  CUDA tensor shaped [B] + permuted([depth,] rows, cols)."""
  import torch
  B = board.shape[0]
  depth = table.shape[1]
  base = [depth, rows, cols] if is_3d else [rows, cols]
  perm = list(permute) if permute is not None else list(range(len(base)))
  shape = [base[i] for i in perm]
  torch_dtype = {0: torch.uint8, 1: torch.int32, 2: torch.float32, 3: torch.int64,
                 4: torch.float64}[_DTYPES[table.dtype]]
  out = torch.empty([B] + shape, dtype=torch_dtype, device=board.device)
  strides = list(out.stride())[1:]
  at = {base_dim: strides[perm.index(base_dim)] for base_dim in range(len(base))}
  spec = _lib.ObserveSpec(depth, _DTYPES[table.dtype], out.stride()[0],
                          at[0] if is_3d else 0,
                          at[1] if is_3d else at[0], at[2] if is_3d else at[1])
  t_table = torch.from_numpy(np.ascontiguousarray(table).view(
      np.uint8 if table.dtype == np.uint8 else np.int32)).to(board.device)
  t_valid = None if valid is None else torch.from_numpy(valid).to(board.device)
  _lib.check(lib.pcl_observe(handle, C.byref(spec), t_table.data_ptr(),
                             None if t_valid is None else t_valid.data_ptr(),
                             board.data_ptr(), out.data_ptr(),
                             None if unknown is None else unknown.data_ptr(), stream),
             'pcl_observe')
  return out
