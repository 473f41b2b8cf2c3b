"""Game-entity base classes: `Backdrop`, `Drape`, `Sprite`.

Same public surface as the reference's `pycolab/things.py` (Backdrop :57-143,
Drape :146-217, Sprite :220-391) so that game code written against pycolab
subclasses these unchanged.  In this engine the objects are *set-up time*
descriptions: constructors run (they decide initial positions, curtains,
visibility), then `lowering.lower()` turns the finished objects into device
records and `update()` is never called from Python — the per-step logic of
every recognised entity class runs inside the fused CUDA step kernels.  After
each step the facade refreshes `position` / `visible` / `curtain` from the
device so read-only peeking (`engine.things[c].position`) keeps working.
"""

import abc
import collections


class Backdrop(object):
  """Background scenery: a uint8 `curtain` plus the `palette` of legal chars."""

  def __init__(self, curtain, palette):
    self._c_u_r_t_a_i_n = curtain
    self._p_a_l_e_t_t_e = palette

  def update(self, actions, board, layers, things, the_plot):
    """The stock Backdrop is static (things.py:102-129)."""

  @property
  def curtain(self):
    return self._c_u_r_t_a_i_n

  @property
  def palette(self):
    return self._p_a_l_e_t_t_e


class Drape(object, metaclass=abc.ABCMeta):
  """A bool mask `curtain` painted with one `character`."""

  def __init__(self, curtain, character):
    self._c_u_r_t_a_i_n = curtain
    self._c_h_a_r_a_c_t_e_r = character

  @abc.abstractmethod
  def update(self, actions, board, layers, backdrop, things, the_plot):
    """Game logic; executed on the device for lowered classes."""

  @property
  def character(self):
    return self._c_h_a_r_a_c_t_e_r

  @property
  def curtain(self):
    return self._c_u_r_t_a_i_n


class Sprite(object, metaclass=abc.ABCMeta):
  """A single cell painted with `character` at `position` when `visible`."""

  Position = collections.namedtuple('Position', ['row', 'col'])

  def __init__(self, corner, position, character):
    self._c_o_r_n_e_r = corner
    self._c_h_a_r_a_c_t_e_r = character
    self._position = position
    self._visible = True

  @abc.abstractmethod
  def update(self, actions, board, layers, backdrop, things, the_plot):
    """Game logic; executed on the device for lowered classes."""

  @property
  def character(self):
    return self._c_h_a_r_a_c_t_e_r

  @property
  def corner(self):
    return self._c_o_r_n_e_r

  @property
  def position(self):
    return self._position

  @property
  def visible(self):
    return self._visible
