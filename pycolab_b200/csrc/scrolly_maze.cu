// scrolly_maze.cu — fused step kernel for examples/scrolly_maze.py.
//
// One launch = Engine.play() for every env (engine.py:583-639): the three
// update groups [['#'], ['a','b','c','P'], ['@']] (scrolly_maze.py:241), the
// Plot consultation, and the final z-ordered render 'abc@#P' (:242).
//
// The two intermediate renders of the reference (one per update group,
// engine.py:735) are never materialised: the only board cells the entities
// read between groups are the <= 9 neighbours of a MazeWalker, tested against
// impassable = '#', and in z-order 'abc@#P' a cell shows '#' iff the wall
// curtain covers it and the (previously rendered) player is not standing on
// it.  Those few look-ups are answered straight from the bit-packed wall
// pattern, one lane per neighbour.
//
// Sprite order P,a,b,c (indices 0..3); drape order '#','@' (0, 1).
// Registers: patroller aux0 = moving_east; P aux0/aux1 = scroll permit mask /
// permit frame; '@' aux0/aux1 = board cell of a coin already removed from the
// pattern but still on the (not yet refreshed) curtain, or -1; plot aux0 =
// coins left in the pattern.
#include "pcl_device.cuh"
#include "pcl_kernels.cuh"

namespace pcl {

namespace {

constexpr int kS = 4;
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ int action_to_motion(int a) {   // scrolly_maze.py:262-271
  return a == 0 ? PCL_M_N : a == 1 ? PCL_M_S : a == 2 ? PCL_M_W
       : a == 3 ? PCL_M_E : a == 4 ? PCL_M_STAY : PCL_M_NONE;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
scrolly_maze_step(const StepParams p) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (env >= p.B) return;
  const int H = p.H, W = p.W;

  int32_t* g_sprites = p.st.d_sprites + (int64_t)env * kS * PCL_SPRITE_WORDS;
  int32_t* g_drapes = p.st.d_drapes + (int64_t)env * 2 * PCL_DRAPE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  const uint32_t* wall_pat = p.st.d_pattern[0] + (int64_t)env * p.st.pattern_bstride[0];
  uint32_t* coin_pat = p.st.d_pattern[1] + (int64_t)env * p.st.pattern_bstride[1];

  Plot plot = load_record_rw<Plot>(g_plot);

  // ---- which envs run, and do they restart?  (engine.py:520-581, 619-624)
  bool restart;
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) return;
  } else {
    restart = plot.game_over && p.auto_reset;
    if (plot.game_over && !p.auto_reset) return;   // reference raises; env stays frozen
  }

  Sprite sp[kS];
  Drape walls, coins;
  int action;
  if (restart) {
    const int episodes = plot.episodes, error = plot.error;
    plot = load_record<Plot>(p.st.d_plot_init + (int64_t)env * p.st.plot_init_bstride);
    plot.episodes = episodes + 1;
    plot.error = error;
    const int32_t* si = p.st.d_sprites_init + (int64_t)env * p.st.sprites_init_bstride;
#pragma unroll
    for (int i = 0; i < kS; ++i) sp[i] = load_record<Sprite>(si + i * PCL_SPRITE_WORDS);
    const int32_t* di = p.st.d_drapes_init + (int64_t)env * p.st.drapes_init_bstride;
    walls = load_record<Drape>(di);
    coins = load_record<Drape>(di + PCL_DRAPE_WORDS);
    // Fresh coins: restore the mutable pattern (one Engine per episode).
    const uint32_t* src = p.st.d_pattern_init[1] + (int64_t)env * p.st.pattern_init_bstride[1];
    const int n = p.PH * p.PWW;
    for (int i = lane; i < n; i += 32) coin_pat[i] = __ldg(src + i);
    __syncwarp();
    action = PCL_ACTION_NONE;
  } else {
#pragma unroll
    for (int i = 0; i < kS; ++i) sp[i] = load_record_rw<Sprite>(g_sprites + i * PCL_SPRITE_WORDS);
    walls = load_record_rw<Drape>(g_drapes);
    coins = load_record_rw<Drape>(g_drapes + PCL_DRAPE_WORDS);
    action = p.actions[(int64_t)env * p.actions_per_env];
  }

  const ScrollyCfg wcfg = scrolly_cfg(H, W, p.PH, p.PW, p.margin[0][0], p.margin[0][1]);
  const ScrollyCfg ccfg = scrolly_cfg(H, W, p.PH, p.PW, p.margin[1][0], p.margin[1][1]);
  Directives dir = fresh_directives();
  const int motion = action_to_motion(action);

  plot.frame += 1;                                           // engine.py:716

  // ---- update group 0: '#' MazeDrape (scrolly_maze.py:308-329)
  if (motion != PCL_M_NONE) scrolly_move(walls, wcfg, motion, plot, sp);

  // Board of render #1 as far as MazeWalkers care: is cell (r, c) a '#'?
  // P is still painted where the previous render put it.
  const bool p_vis = visible(sp[0]);
  const int p_row = sp[0].row, p_col = sp[0].col;
  const int wr = walls.corner_r, wc = walls.corner_c;
  auto is_wall = [&](int r, int c) -> bool {
    if (p_vis && r == p_row && c == p_col) return false;
    return bit_at(wall_pat + (int64_t)(wr + r) * p.PWW, wc + c);
  };

  // ---- update group 1: patrollers a, b, c then P
  const int p_vrow = sp[0].vrow, p_vcol = sp[0].vcol;   // P moves after them
#pragma unroll
  for (int i = 1; i < kS; ++i) {                             // PatrollerSprite :284-305
    if (plot.frame % 2) {
      walker_move(sp[i], i, PCL_M_STAY, plot, H, W, false, false, lane, is_wall);
    } else {
      scrolly_touch_prescroll(walls, plot);
      int pr = sp[i].vrow + walls.pre_r;
      int pc = sp[i].vcol + walls.pre_c + (sp[i].aux0 ? 1 : -1);
      // NumPy indexing: negatives wrap once, anything else is an IndexError.
      if (pr < 0) pr += p.PH;
      if (pc < 0) pc += p.PW;
      bool next_to_wall = false;
      if ((unsigned)pr < (unsigned)p.PH && (unsigned)pc < (unsigned)p.PW)
        next_to_wall = bit_at(wall_pat + (int64_t)pr * p.PWW, pc);
      else
        plot.error |= PCL_ENV_ERR_INDEX;
      if (next_to_wall) sp[i].aux0 = !sp[i].aux0;
      walker_move(sp[i], i, sp[i].aux0 ? PCL_M_E : PCL_M_W, plot, H, W, false, false,
                  lane, is_wall);
      if (sp[i].vrow == p_vrow && sp[i].vcol == p_vcol) terminate(dir);
    }
  }
  if (motion != PCL_M_NONE)                                  // PlayerSprite :258-271
    walker_move(sp[0], 0, motion, plot, H, W, false, true, lane, is_wall);

  // ---- update group 2: '@' CashDrape (scrolly_maze.py:341-364)
  scrolly_touch_prescroll(coins, plot);
  {
    const int pr = coins.pre_r + sp[0].row, pc = coins.pre_c + sp[0].col;
    uint32_t* word = coin_pat + (int64_t)pr * p.PWW + (pc >> 5);
    const uint32_t w = *word;
    if ((w >> (pc & 31)) & 1u) {
      add_reward(dir, 100);
      __syncwarp();
      if (lane == 0) *word = w & ~(1u << (pc & 31));
      __syncwarp();
      plot.aux0 -= 1;
      if (plot.aux0 == 0) terminate(dir);
      coins.aux0 = sp[0].row; coins.aux1 = sp[0].col;   // stale until next refresh
    }
  }
  if (motion != PCL_M_NONE) {
    scrolly_move(coins, ccfg, motion, plot, sp);
    coins.aux0 = -1; coins.aux1 = -1;                    // _update_curtain :689
  } else if (action == 5) {
    terminate(dir);
  }

  // ---- _apply_and_clear_plot (engine.py:761-847); no z-order changes here.
  plot.game_over = dir.game_over;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kS; ++i) store_record(g_sprites + i * PCL_SPRITE_WORDS, sp[i]);
    store_record(g_drapes, walls);
    store_record(g_drapes + PCL_DRAPE_WORDS, coins);
    store_record(g_plot, plot);
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
  }

  // ---- final render, z-order a b c @ # P (engine.py:737-759)
  const uint8_t* backdrop = p.st.d_backdrop + (int64_t)env * p.st.backdrop_bstride;
  uint8_t* board = p.out.d_board + (int64_t)env * H * p.pitch;
  const int segs_per_row = p.pitch >> 4;
  const int total = H * segs_per_row;
  const int cr = coins.corner_r, cc = coins.corner_c;
  const int wr2 = walls.corner_r, wc2 = walls.corner_c;
  const int stale_r = coins.aux0, stale_c = coins.aux1;
  for (int seg = lane; seg < total; seg += 32) {
    const int r = seg / segs_per_row;
    const int c0 = (seg - r * segs_per_row) << 4;
    const int ncols = min(16, W - c0);
    const unsigned valid = (1u << ncols) - 1u;
    uint4 px = __ldg(reinterpret_cast<const uint4*>(backdrop + (int64_t)r * p.pitch + c0));
    unsigned coin_bits = bits16(coin_pat + (int64_t)(cr + r) * p.PWW, cc + c0) & valid;
    const unsigned wall_bits = bits16(wall_pat + (int64_t)(wr2 + r) * p.PWW, wc2 + c0) & valid;
    if (r == stale_r && (unsigned)(stale_c - c0) < 16u) coin_bits |= 1u << (stale_c - c0);
#pragma unroll
    for (int i = 1; i < kS; ++i) paint_bits(px, sprite_bit(sp[i], r, c0), p.sprite_char[i]);
    paint_bits(px, coin_bits, '@');
    paint_bits(px, wall_bits, '#');
    paint_bits(px, sprite_bit(sp[0], r, c0), p.sprite_char[0]);
    *reinterpret_cast<uint4*>(board + (int64_t)r * p.pitch + c0) = px;
  }
}

}  // namespace

cudaError_t launch_scrolly_maze(const StepParams& p, cudaStream_t s) {
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  scrolly_maze_step<<<blocks, kWarpsPerBlock * 32, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pcl
