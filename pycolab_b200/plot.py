"""`Plot`: the per-game blackboard (reference `pycolab/plot.py:30-385`).

Same methods and properties as the reference.  During play the authoritative
per-env plot state (frame, reward/termination directives, scrolling-protocol
registers) lives in the device plot record (include/pcl.h PCL_P_*); the facade
`Engine` mirrors `frame` into this object after every step.  The directive
methods work at set-up time exactly as upstream (same ValueErrors).
"""


class Plot(dict):

  class _EngineDirectives(object):
    __slots__ = ('z_updates', 'summed_reward', 'game_over', 'discount')

    def __init__(self):
      self.z_updates = []
      self.summed_reward = None
      self.game_over = False
      self.discount = 1.0

  def __init__(self):
    dict.__init__(self)
    self._frame = -1
    self._update_group = None
    self._prior_chapter = self._this_chapter = self._next_chapter = None
    self._clear_engine_directives()

  # -- directives (plot.py:136-260) ---------------------------------------
  def change_z_order(self, move_this, in_front_of_that):
    for ch in (move_this,) + (() if in_front_of_that is None else (in_front_of_that,)):
      try:
        ord(ch)
      except TypeError:
        raise ValueError('{!r} was used as an argument in a call to change_z_order, '
                         'but only single ASCII characters are valid '
                         'arguments'.format(ch))
    self._engine_directives.z_updates.append((move_this, in_front_of_that))

  def terminate_episode(self, discount=0.0):
    if not 0.0 <= discount <= 1.0:
      raise ValueError('Discount must be in range [0,1].')
    self._engine_directives.game_over = True
    self._engine_directives.discount = discount

  def add_reward(self, reward):
    d = self._engine_directives
    d.summed_reward = reward if d.summed_reward is None else d.summed_reward + reward

  def log(self, message):
    from pycolab_b200.protocols import logging as plab_logging
    plab_logging.log(self, message)

  def change_default_discount(self, discount):
    if not 0.0 <= discount <= 1.0:
      raise ValueError('Default discount must be in range [0,1].')
    self._engine_directives.discount = discount

  # -- read-only state (plot.py:262-341) ----------------------------------
  @property
  def frame(self):
    return self._frame

  @frame.setter
  def frame(self, val):
    assert val == self._frame + 1
    self._frame = val

  @property
  def update_group(self):
    return self._update_group

  @update_group.setter
  def update_group(self, group):
    self._update_group = group

  @property
  def default_discount(self):
    return self._engine_directives.discount

  @property
  def prior_chapter(self):
    return self._prior_chapter

  @prior_chapter.setter
  def prior_chapter(self, val):
    self._prior_chapter = val

  @property
  def this_chapter(self):
    return self._this_chapter

  @this_chapter.setter
  def this_chapter(self, val):
    self._this_chapter = val

  @property
  def next_chapter(self):
    return self._next_chapter

  @next_chapter.setter
  def next_chapter(self, val):
    self._next_chapter = val

  def _clear_engine_directives(self):
    self._engine_directives = self._EngineDirectives()

  def _get_engine_directives(self):
    return self._engine_directives
