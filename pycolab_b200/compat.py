"""Let code written for `pycolab` import this package unchanged.

    import pycolab_b200.compat; pycolab_b200.compat.install()
    from pycolab import ascii_art            # -> pycolab_b200.ascii_art
    from pycolab.prefab_parts import sprites # -> pycolab_b200.prefab_parts.sprites

`install()` registers aliases in `sys.modules`.  A game module that lives
elsewhere (e.g. a checkout of the reference's `pycolab/examples/*.py`) can then
be loaded with `load_example(path)`: its `from pycolab import ...` lines bind
to this package, its entity classes subclass this package's prefabs, and
`lowering` recognises them by (module name, class name).
"""

import importlib
import importlib.util
import os
import sys
import types

_ALIASED = ('things', 'plot', 'engine', 'ascii_art', 'rendering', 'cropping', 'storytelling',
            'prefab_parts', 'prefab_parts.sprites', 'prefab_parts.drapes',
            'protocols', 'protocols.scrolling', 'protocols.logging')


def install():
  """Alias `pycolab[.x]` -> `pycolab_b200[.x]` in sys.modules (idempotent)."""
  import pycolab_b200
  if sys.modules.get('pycolab') is pycolab_b200:
    return
  if 'pycolab' in sys.modules:
    raise RuntimeError('a different `pycolab` is already imported')
  sys.modules['pycolab'] = pycolab_b200
  for name in _ALIASED:
    mod = importlib.import_module('pycolab_b200.' + name)
    sys.modules['pycolab.' + name] = mod
  # The example modules import human_ui (curses front-end, out of scope) at
  # module level; give them an inert stand-in.
  ui = types.ModuleType('pycolab.human_ui')

  class CursesUi(object):
    def __init__(self, *args, **kwargs):
      raise NotImplementedError('the curses front-end is not part of pycolab_b200')
  ui.CursesUi = CursesUi
  sys.modules['pycolab.human_ui'] = ui
  pycolab_b200.human_ui = ui


def uninstall():
  for name in list(sys.modules):
    if name == 'pycolab' or name.startswith('pycolab.'):
      del sys.modules[name]


def load_example(path, name=None):
  """Import the game module at `path` against this package."""
  install()
  name = name or os.path.splitext(os.path.basename(path))[0]
  spec = importlib.util.spec_from_file_location('pycolab.examples.' + name, path)
  module = importlib.util.module_from_spec(spec)
  sys.modules[spec.name] = module
  spec.loader.exec_module(module)
  return module
