"""Errors specific to the B200 engine."""


class NotLoweredError(NotImplementedError):
  """The game uses an entity class (or structure) that has no device program.

  pycolab_b200 has no CPU path: game logic runs only inside the fused CUDA
  step kernels, which exist for the entity classes listed in
  `pycolab_b200.lowering.LOWERED_CLASSES`.
  """


class DeviceOnlyError(RuntimeError):
  """A per-step helper was called from Python; per-step logic is device code."""
