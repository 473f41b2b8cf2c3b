import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pycolab_b200 import batched, levels, lowering
from pycolab_b200.games import scrolly_maze
games = [lowering.lower(scrolly_maze.make_game(*levels.scrolly_maze_level(1000 + i))) for i in range(8)]
B, R, K = 4096, 6, 300
engines = [batched.BatchedEngine(games, batch=B) for _ in range(R)]
for e in engines: e.its_showtime()
acts = torch.from_numpy(np.random.RandomState(0).randint(0, 5, size=(K, B)).astype(np.int32)).cuda()
for t in range(30): engines[t % R].play(acts[t])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for t in range(K): engines[t % R].play(acts[t])
b.record(); torch.cuda.synchronize()
print('PCL_DEBUG=%s  us/step=%.2f' % (os.environ.get('PCL_DEBUG', '0'), 1000 * a.elapsed_time(b) / K))
