"""Multi-GPU sharding: one process per GPU, envs partitioned by index.

Environments never interact (one `Engine`/`Plot` per env upstream,
engine.py:216), so the step path shards with NO data-path collective: rank r
owns the contiguous block of env indices `shard_range(global_batch, r, world)`
together with their action stream and RNG states (seeded by GLOBAL env index,
so a sharded run reproduces the single-GPU run env for env).  The only
exchange is the optional hand-off of the per-step outputs to every rank
(`allgather_outputs`, one NCCL all-gather per tensor over NVLink/NVSwitch).
"""


def shard_range(global_batch, rank, world):
  """Contiguous block partition: (first global env index, count) of `rank`."""
  if not 0 <= rank < world:
    raise ValueError('rank %d outside world of %d' % (rank, world))
  base, extra = divmod(global_batch, world)
  count = base + (1 if rank < extra else 0)
  first = rank * base + min(rank, extra)
  return first, count


def make_shard_engine(games, global_batch, rank, world, device, **kwargs):
  """`BatchedEngine` for this rank's block of a `global_batch`-env job.

  games[i] is the level of GLOBAL env i (mod len(games)): the local list is
  rotated so that local env e maps to global env first + e."""
  from pycolab_b200 import batched
  first, count = shard_range(global_batch, rank, world)
  n = len(games)
  local = [games[(first + i) % n] for i in range(min(n, count))] if n > 1 else games
  return batched.BatchedEngine(local, batch=count, device=device, env_offset=first,
                               **kwargs)


def allgather_outputs(tensors, global_batch, group=None):
  """All-gather per-env output tensors (dim 0 = local envs) to [global_batch, ...].

  Shards may differ by one env when world does not divide global_batch; every
  rank pads to the largest shard, gathers, then drops the padding."""
  import torch
  import torch.distributed as dist
  world = dist.get_world_size(group)
  counts = [shard_range(global_batch, r, world)[1] for r in range(world)]
  biggest = max(counts)
  out = []
  for t in tensors:
    pad = t
    if t.shape[0] < biggest:
      pad = torch.cat([t, t.new_zeros((biggest - t.shape[0],) + tuple(t.shape[1:]))])
    pad = pad.contiguous()
    gathered = pad.new_empty((world * biggest,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(gathered, pad, group=group)
    if all(c == biggest for c in counts):
      out.append(gathered)
    else:
      pieces = [gathered[r * biggest: r * biggest + counts[r]] for r in range(world)]
      out.append(torch.cat(pieces))
  return out
