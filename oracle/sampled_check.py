"""Sampled-env parity checks at full batch.  TEST INFRASTRUCTURE ONLY.

The oracle (a Python restatement of reference pycolab/engine.py:583-639 and the
example games) steps ~2x10^4 env-steps/s/core, so a 4096..16384-env batch cannot
be replayed whole.  These helpers pick a SAMPLE of env indices out of a
full-size `BatchedEngine`, step one oracle world per sampled env with that env's
own level and action stream (auto-reset = a fresh world on game-over, the
batched stand-in for "one Engine per episode", engine.py:103-104), and compare
every step's (board, reward, has_reward, discount, done) bit for bit.

Used by tests/test_gpu_sampled_parity.py and by bench.py's post-timing
`parity_checked` leg (the checker, never the thing measured).
"""

import numpy as np


class Mismatch(AssertionError):
  pass


def _compare(t, env, got, want, world):
  board, reward, has, disc, done = got
  w_board, w_reward, w_disc = want
  if not np.array_equal(board, w_board):
    bad = np.argwhere(board != w_board)
    raise Mismatch('board differs at step %d env %d, first cell %s: device %r oracle %r' % (
        t, env, tuple(bad[0]), chr(board[tuple(bad[0])]), chr(w_board[tuple(bad[0])])))
  want_has = 0 if w_reward is None else 1
  want_reward = 0 if w_reward is None else int(w_reward)
  if (int(has), int(reward)) != (want_has, want_reward):
    raise Mismatch('reward differs at step %d env %d: device (%d, %d) oracle %r' % (
        t, env, int(has), int(reward), w_reward))
  if float(disc) != float(w_disc):
    raise Mismatch('discount differs at step %d env %d: %r vs %r' % (t, env, disc, w_disc))
  if bool(done) != bool(world.game_over):
    raise Mismatch('game_over differs at step %d env %d' % (t, env))


def lockstep(engine, make_world, env_ids, actions, crop=None):
  """Step `engine` (a BatchedEngine after its_showtime(), B envs) with
  actions int32 [T, B] and, for every env in `env_ids`, an oracle world built by
  make_world(env) in lockstep.  crop: optional (crop_spec, crop_state,
  make_oracle_cropper) to compare a cropper view too.  Returns the number of
  (env, step) pairs compared, resets included."""
  import torch
  env_ids = [int(e) for e in env_ids]
  idx = torch.as_tensor(env_ids, dtype=torch.long, device=engine.device)
  T = actions.shape[0]
  worlds = {e: make_world(e) for e in env_ids}
  outs = {e: worlds[e].its_showtime() for e in env_ids}
  croppers = None
  if crop is not None:
    crop_spec, crop_state, make_cropper = crop
    croppers = {e: make_cropper() for e in env_ids}
    for e in env_ids:
      croppers[e].set_engine(worlds[e])
  acts_dev = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int32)).to(engine.device)
  compared = 0

  def check(t):
    boards = engine.board.index_select(0, idx).cpu().numpy()
    reward = engine.reward.index_select(0, idx).cpu().numpy()
    has = engine.has_reward.index_select(0, idx).cpu().numpy()
    disc = engine.discount.index_select(0, idx).cpu().numpy()
    done = engine.done.index_select(0, idx).cpu().numpy()
    views = None
    if crop is not None:
      views = engine.crop(crop_spec, state=crop_state).index_select(0, idx).cpu().numpy()
    for k, e in enumerate(env_ids):
      _compare(t, e, (boards[k], reward[k], has[k], disc[k], done[k]), outs[e], worlds[e])
      if views is not None:
        want = croppers[e].crop(outs[e][0])
        if not np.array_equal(views[k], want):
          raise Mismatch('crop differs at step %d env %d' % (t, e))
    return len(env_ids)

  compared += check(0)
  for t in range(T):
    engine.play(acts_dev[t])
    for e in env_ids:
      if worlds[e].game_over:                 # the auto-reset rule
        worlds[e] = make_world(e)
        if croppers is not None:
          croppers[e].set_engine(worlds[e])
        outs[e] = worlds[e].its_showtime()
      else:
        outs[e] = worlds[e].play(int(actions[t, e]))
    compared += check(t + 1)
  return compared


def replay(make_world, env, action_stream):
  """The oracle's state of one env after `action_stream` (ints, one per step the
  device took on that env) from a fresh its_showtime(), auto-reset included.
  Returns (world, last output triple)."""
  world = make_world(env)
  out = world.its_showtime()
  for a in action_stream:
    if world.game_over:
      world = make_world(env)
      out = world.its_showtime()
    else:
      out = world.play(int(a))
  return world, out


def final_state_check(engine, make_world, env_ids, action_streams, sprite_chars):
  """After the device ran `action_streams[env]` (from a fresh reset) on each
  sampled env: the last board, the last (reward, discount, done) and every
  sprite's (row, col, visible) must equal the oracle's replay."""
  import torch
  env_ids = [int(e) for e in env_ids]
  idx = torch.as_tensor(env_ids, dtype=torch.long, device=engine.device)
  boards = engine.board.index_select(0, idx).cpu().numpy()
  reward = engine.reward.index_select(0, idx).cpu().numpy()
  has = engine.has_reward.index_select(0, idx).cpu().numpy()
  disc = engine.discount.index_select(0, idx).cpu().numpy()
  done = engine.done.index_select(0, idx).cpu().numpy()
  sprites = engine.sprites.index_select(0, idx).cpu().numpy()
  steps = 0
  for k, e in enumerate(env_ids):
    world, out = replay(make_world, e, action_streams[e])
    t = len(action_streams[e])
    _compare(t, e, (boards[k], reward[k], has[k], disc[k], done[k]), out, world)
    for i, ch in enumerate(sprite_chars):
      s = world.things[ch]
      got = (int(sprites[k, i, 0]), int(sprites[k, i, 1]), bool(sprites[k, i, 4] & 1))
      want = (int(s.position[0]), int(s.position[1]), bool(s.visible))
      if got != want:
        raise Mismatch('sprite %s differs after %d steps env %d: device %r oracle %r' % (
            ch, t, e, got, want))
    steps += t
  return steps
