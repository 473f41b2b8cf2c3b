"""Shared trajectory protocol for golden fixtures and parity tests.

An "env" is anything with `its_showtime()`, `play(action)` and `game_over`
(a reference Engine, an oracle World, the B200 facade Engine).  The protocol is
the batched engine's auto-reset rule: an env that reported game over is rebuilt
at the NEXT step (that step's action is ignored) and returns its
`its_showtime()` outputs.
"""

import numpy as np


def board_of(obs):
  return np.asarray(obs.board if hasattr(obs, 'board') else obs, dtype=np.uint8)


def run_trajectory(make_env, actions, convert_action=None, on_frame=None):
  """Returns dict(boards[T+1,H,W] u8, reward[T+1] i64, has_reward[T+1] u8,
  discount[T+1] f64, game_over[T+1] u8)."""
  env = make_env()
  out = env.its_showtime()
  boards, reward, has_reward, discount, over = [], [], [], [], []

  def record(env, out):
    boards.append(board_of(out[0]).copy())
    reward.append(0 if out[1] is None else int(out[1]))
    has_reward.append(0 if out[1] is None else 1)
    discount.append(float(out[2]))
    over.append(1 if env.game_over else 0)
    if on_frame is not None:
      on_frame(env, out)

  record(env, out)
  for a in actions:
    if env.game_over:
      env = make_env()
      out = env.its_showtime()
    else:
      out = env.play(convert_action(a) if convert_action else a)
    record(env, out)
  return dict(boards=np.stack(boards), reward=np.array(reward, dtype=np.int64),
              has_reward=np.array(has_reward, dtype=np.uint8),
              discount=np.array(discount, dtype=np.float64),
              game_over=np.array(over, dtype=np.uint8))


def art_to_u8(art):
  return np.vstack([np.frombuffer(l.encode('ascii'), dtype=np.uint8) for l in art])


def u8_to_art(arr):
  return [bytes(row).decode('ascii') for row in np.asarray(arr, dtype=np.uint8)]


def assert_same_trajectory(want, got, label=''):
  for key in ('boards', 'reward', 'has_reward', 'discount', 'game_over'):
    w, g = np.asarray(want[key]), np.asarray(got[key])
    assert w.shape == g.shape, (label, key, w.shape, g.shape)
    if not np.array_equal(w, g):
      bad = np.argwhere(w != g)[0]
      raise AssertionError('%s: %s differs first at %s: want %r got %r' % (
          label, key, tuple(bad), w[tuple(bad)], g[tuple(bad)]))
