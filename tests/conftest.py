import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200 box)')


def _cuda_state():
  """(usable, reason): can the `gpu` tests run in this process?"""
  try:
    import torch
  except ImportError:
    return False, 'torch is not importable'
  if not torch.cuda.is_available():
    return False, 'no CUDA device in this process'
  return True, ''


def pytest_collection_modifyitems(config, items):
  """Without a CUDA device the `gpu` tests are SKIPPED (with the reason), so a
  bare `pytest tests` on a CPU box separates real regressions from a missing
  device.  Where a device IS visible nothing is skipped: a missing or unloadable
  libpcl.so fails loudly there (the product has no CPU path)."""
  if os.environ.get('PCL_TESTS_REQUIRE_GPU') == '1':
    return
  ok, reason = _cuda_state()
  if ok:
    return
  skip = pytest.mark.skip(reason='gpu test: ' + reason)
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)
