"""Plot logging protocol (reference `pycolab/protocols/logging.py:33-70`).

Host-side only: strings never travel to the device.  Lowered entity classes
that log upstream (e.g. CashDrape's "Coin collected" message,
examples/scrolly_maze.py:348) do not log here.
"""

_KEY = 'log_messages'


def log(the_plot, message):
  the_plot.setdefault(_KEY, []).append(message)


def consume(the_plot):
  messages = the_plot.setdefault(_KEY, [])
  the_plot[_KEY] = []
  return messages
