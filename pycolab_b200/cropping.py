"""Observation croppers (reference `pycolab/cropping.py:27-598`).

Same classes and constructor arguments.  `crop()` runs on the device
(`pcl_crop`, csrc/render.cu `crop_kernel`): the window corner of a
`ScrollingCropper` is per-env device state next to the plot record, so in a
`BatchedEngine` every env pans/saccades its own window; this facade drives the
batch-1 engine behind a `pycolab_b200.engine.Engine`.
"""

import copy

from pycolab_b200 import rendering
from pycolab_b200.errors import NotLoweredError


class ObservationCropper(object):
  """Identity cropper / base class (cropping.py:27-227)."""

  def __init__(self):
    self._engine = None
    self._pad_char = None

  def set_engine(self, engine):
    if engine is not self._engine:
      self._state = None          # a new Engine forgets the window (cropping.py:375-391)
    self._engine = engine

  def crop(self, observation):
    return observation

  @property
  def rows(self):
    return self._engine.rows

  @property
  def cols(self):
    return self._engine.cols

  def _check_pad(self):
    if self._pad_char is not None:
      legal = set(self._engine.things) | set(self._engine.backdrop.palette)
      if self._pad_char not in legal:
        raise ValueError("An `ObservationCropper` tried to fill empty space with a "
                         "character that isn't used by the current game engine.")

  def _device_crop(self, spec):
    from pycolab_b200 import batched
    self._check_pad()
    b = self._engine.batched
    if b is None:
      raise RuntimeError('crop() called before the Engine entered play mode')
    if getattr(self, '_state', None) is None:
      self._state = b.new_crop_state()
    board = b.crop(spec, state=self._state)[0].cpu().numpy().copy()
    chars = set(self._engine.things) | set(self._engine.backdrop.palette)
    return rendering.Observation(board=board, layers=rendering.LazyLayers(board, chars))


class FixedCropper(ObservationCropper):
  """A fixed window, optionally padded (cropping.py:229-310)."""

  def __init__(self, top_left_corner, rows, cols, pad_char=None):
    super(FixedCropper, self).__init__()
    self._top_row, self._left_col = top_left_corner
    self._rows, self._cols = rows, cols
    self._pad_char = pad_char

  def crop(self, observation):
    from pycolab_b200 import _lib
    if self._pad_char is None and (
        self._top_row < 0 or self._left_col < 0 or
        self._top_row + self._rows > self._engine.rows or
        self._left_col + self._cols > self._engine.cols):
      raise RuntimeError('An ObservationCropper attempted to crop a region that extends '
                         'beyond the observation without specifying a character to fill '
                         'the void that exists out there.')
    spec = _lib.CropSpec(self._rows, self._cols, -1,
                         -1 if self._pad_char is None else ord(self._pad_char),
                         0, 0, self._top_row, self._left_col, 0)
    return self._device_crop(spec)

  @property
  def rows(self):
    return self._rows

  @property
  def cols(self):
    return self._cols


class ScrollingCropper(ObservationCropper):
  """A window that follows an entity (cropping.py:313-598)."""

  def __init__(self, rows, cols, to_track, pad_char=None, scroll_margins=(2, 3),
               initial_offset=None, saccade=True):
    super(ScrollingCropper, self).__init__()
    from pycolab_b200 import batched
    self._rows, self._cols = rows, cols
    self._to_track = copy.copy(to_track)
    self._pad_char = pad_char
    # Validates and resolves margins exactly as cropping.py:362-380.
    self._spec_args = dict(pad_char=pad_char, scroll_margins=scroll_margins,
                           initial_offset=initial_offset, saccade=saccade)
    batched.scrolling_crop_spec(rows, cols, 0, **self._spec_args)

  def set_engine(self, engine):
    prior = self._engine
    super(ScrollingCropper, self).set_engine(engine)
    if engine is not prior:
      if ((engine.rows < self._rows or engine.cols < self._cols) and
          self._pad_char is None):
        raise ValueError(
            "A ScrollingCropper with a size of {} and no pad character can't be used "
            'with a pycolab engine that produces smaller observations in any dimension '
            '(in this case, {})'.format((self._rows, self._cols),
                                        (engine.rows, engine.cols)))

  def crop(self, observation):
    from pycolab_b200 import batched
    from pycolab_b200 import _lib
    b = self._engine.batched
    if b is None:
      raise RuntimeError('crop() called before the Engine entered play mode')
    if not 0 < len(self._to_track) <= _lib.MAX_TRACK:
      raise NotLoweredError('the device cropper tracks 1..{} entities'.format(_lib.MAX_TRACK))
    codes = []                     # priority list: k > 0 sprite k - 1, k < 0 drape -k - 1
    for char in self._to_track:
      if char not in self._engine.things:
        raise RuntimeError('ScrollingCropper was told to track a nonexistent game entity '
                           '{!r}.'.format(char))
      codes.append(b.sprite_chars.index(char) + 1 if char in b.sprite_chars
                   else -(b.drape_chars.index(char) + 1))
    spec = batched.scrolling_crop_spec(self._rows, self._cols, 0, track=codes,
                                       **self._spec_args)
    return self._device_crop(spec)

  @property
  def rows(self):
    return self._rows

  @property
  def cols(self):
    return self._cols
