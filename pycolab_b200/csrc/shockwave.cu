// shockwave.cu — fused step kernel for examples/shockwave.py:91-197.
//
// One update group [' ', '^', 'P', '@'], z-order ' ' '^' '@' 'P' over a backdrop of '+'
// (what lies beneath everything) and '=' (walls).  Sprite 0 = the player, a MazeWalker
// confined to the board with impassable '='; drapes 0 = '@' (the shockwave), 1 = ' ' (the
// danger zone), 2 = '^' (the safe zone); ' ' and '^' never change (MinimalDrape), so
// their curtains are read from the per-level templates (pcl_state.d_bits_init[1..2]).
//
//   P: action 0 = _north, 1 = _west, 2 = _east, 3 = _stay, anything else: no call  (:98-109)
//   '@' (after P, same group, so every `layers[...]` it consults is the STALE board of
//   the previous render, engine.py:725):
//     curtain empty -> impact = np.random.randint(0, H * W) (MT19937, masked rejection),
//       distance = Euclidean distance to it, steps_since_impact = 0               (:129-140)
//     curtain = steps < distance <= steps + width, minus walls — compared on SQUARED
//       integer distances, exact because steps is an integer                       (:145-149)
//     P's position shows '^' on the stale board (safe-zone cell not covered by the OLD
//       curtain, nor by P itself one frame ago): reward +1, terminate              (:152-156)
//     P under the NEW curtain and in the danger zone: reward -1, terminate         (:158-163)
//     steps_since_impact += 1
// '=' is never covered on a rendered board (each art cell belongs to exactly one of the
// backdrop / a drape / the sprite, and the curtain excludes walls), so "stale board == '='"
// is "backdrop == '='", for the walker's impassable test as well.
// The shockwave's curtain lives bit-packed in pcl_state.d_bits[0] (one row per lane), its
// impact cell and step count in the drape record's AUX0 / AUX1; program_arg[0] = width.
// One warp per env; boards up to 32 rows x 64 columns.
#include "pcl_device.cuh"
#include "pcl_kernels.cuh"
#include "pcl_mt.cuh"

namespace pcl {

namespace {

constexpr int kWarpsPerBlock = 4;
typedef unsigned long long u64;

__device__ __forceinline__ bool in_set(const uint32_t (&set)[4], int code) {
  return (set[(code >> 5) & 3] >> (code & 31)) & 1u;
}
__device__ __forceinline__ u64 row_bits(const uint32_t* base, int r, int BW) {
  const uint32_t* row = base + (int64_t)r * BW;
  return (u64)row[0] | (BW > 1 ? (u64)row[1] << 32 : 0ull);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
shockwave_step(const StepParams p) {
  __shared__ u64 s_rows[kWarpsPerBlock][32];          // the new curtain, a row per lane, for the render
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (env >= p.B) return;
  const int64_t lvl = p.st.d_level ? p.st.d_level[env] : env;
  const int H = p.H, W = p.W, pitch = p.pitch, BW = p.BW;
  int32_t* g_sprite = p.st.d_sprites + (int64_t)env * PCL_SPRITE_WORDS;
  int32_t* g_drapes = p.st.d_drapes + (int64_t)env * 3 * PCL_DRAPE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  const uint8_t* backdrop = p.st.d_backdrop + lvl * p.st.backdrop_bstride;
  uint32_t* wave_bits = p.st.d_bits[0] + (int64_t)env * p.st.bits_bstride[0];
  const uint32_t* danger_bits = p.st.d_bits_init[1] + lvl * p.st.bits_init_bstride[1];
  const uint32_t* safe_bits = p.st.d_bits_init[2] + lvl * p.st.bits_init_bstride[2];

  const int was_over = g_plot[PCL_P_GAME_OVER];
  bool restart;
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) return;
  } else {
    restart = was_over && p.auto_reset;
    if (was_over && !p.auto_reset) return;           // reference raises; env stays frozen
  }
  const int32_t* src_s = restart ? p.st.d_sprites_init + lvl * p.st.sprites_init_bstride : g_sprite;
  const int32_t* src_d = restart ? p.st.d_drapes_init + lvl * p.st.drapes_init_bstride : g_drapes;
  const int32_t* src_p = restart ? p.st.d_plot_init + lvl * p.st.plot_init_bstride : g_plot;
  Sprite pl;
  pl.row = src_s[PCL_S_ROW]; pl.col = src_s[PCL_S_COL]; pl.vrow = src_s[PCL_S_VROW];
  pl.vcol = src_s[PCL_S_VCOL]; pl.flags = src_s[PCL_S_FLAGS]; pl.aux0 = pl.aux1 = pl.aux2 = 0;
  int impact = src_d[PCL_D_AUX0], steps = src_d[PCL_D_AUX1];
  Plot plot;
  plot.frame = src_p[PCL_P_FRAME] + 1;                                // engine.py:716
  plot.error = g_plot[PCL_P_ERROR];
  plot.order_frame = PCL_NEVER; plot.order_r = plot.order_c = 0; plot.ego_mask = 0;
  const int episodes = g_plot[PCL_P_EPISODES] + (restart ? 1 : 0);
  const int action = restart ? PCL_ACTION_NONE : p.actions[(int64_t)env * p.actions_per_env];
  Directives dir = fresh_directives();
  // Row `lane` of the curtain the LAST render showed (the art's '@' cells after a restart).
  const uint32_t* old_base = restart ? p.st.d_bits_init[0] + lvl * p.st.bits_init_bstride[0] : wave_bits;
  const u64 old_row = lane < H ? row_bits(old_base, lane, BW) : 0ull;
  const int old_row_p = pl.row, old_col_p = pl.col;
  const bool old_vis_p = visible(pl);

  // ---- PlayerSprite.update (:98-109)
  const int motion = action == 0 ? PCL_M_N : action == 1 ? PCL_M_W : action == 2 ? PCL_M_E
                   : action == 3 ? PCL_M_STAY : PCL_M_NONE;
  auto wall = [&](int r, int c) { return in_set(p.impassable[0], backdrop[r * pitch + c]); };
  if (motion != PCL_M_NONE)
    walker_move(pl, 0, motion, plot, H, W, p.confined[0] != 0, false, lane, wall);

  // ---- ShockwaveDrape.update (:126-165)
  if (!__any_sync(PCL_FULL, old_row != 0ull)) {                       // :129
    uint32_t* mt = p.st.d_rng + (int64_t)env * PCL_MT_WORDS;
    impact = (int)mt_below(mt, (uint32_t)(H * W), lane);              // np.random.randint(0, size)
    steps = 0;
  }
  const int ir = impact / W, ic = impact - ir * W;                    // np.unravel_index
  const int width = p.program_arg[0];
  const int lo2 = steps * steps, hi2 = (steps + width) * (steps + width);
  u64 new_row = 0ull;
  if (lane < H) {
    const int dr2 = (lane - ir) * (lane - ir);
    for (int c = 0; c < W; ++c) {
      const int d2 = dr2 + (c - ic) * (c - ic);
      if (d2 > lo2 && d2 <= hi2 && !wall(lane, c)) new_row |= 1ull << c;
    }
  }
  // The player's cell, as every lane needs it: rows of the old / new curtain and the
  // two static drapes at P's row.
  const int pr = pl.row, pc = pl.col;
  const u64 old_at = __shfl_sync(PCL_FULL, old_row, pr);
  const u64 new_at = __shfl_sync(PCL_FULL, new_row, pr);
  const bool safe_here = (row_bits(safe_bits, pr, BW) >> pc) & 1ull;
  const bool danger_here = (row_bits(danger_bits, pr, BW) >> pc) & 1ull;
  const bool stale_shows_safe = safe_here && !((old_at >> pc) & 1ull) &&
                                !(old_vis_p && old_row_p == pr && old_col_p == pc);
  if (stale_shows_safe) { add_reward(dir, 1); terminate(dir); }      // :152-156
  if (((new_at >> pc) & 1ull) && danger_here) { add_reward(dir, -1); terminate(dir); }
  steps += 1;

  s_rows[threadIdx.x >> 5][lane] = new_row;
  __syncwarp();
  if (lane < H) {
    uint32_t* row = wave_bits + (int64_t)lane * BW;
    row[0] = (uint32_t)new_row;
    if (BW > 1) row[1] = (uint32_t)(new_row >> 32);
  }
  if (lane == 0) {
    g_sprite[PCL_S_ROW] = pl.row; g_sprite[PCL_S_COL] = pl.col; g_sprite[PCL_S_VROW] = pl.vrow;
    g_sprite[PCL_S_VCOL] = pl.vcol; g_sprite[PCL_S_FLAGS] = pl.flags;
    g_sprite[PCL_S_AUX0] = 0; g_sprite[PCL_S_AUX1] = 0; g_sprite[PCL_S_AUX2] = 0;
    g_drapes[PCL_D_AUX0] = impact; g_drapes[PCL_D_AUX1] = steps;
    g_drapes[PCL_D_LAST_FRAME] = PCL_NEVER;
    g_plot[PCL_P_FRAME] = plot.frame; g_plot[PCL_P_GAME_OVER] = dir.game_over;
    g_plot[PCL_P_EPISODES] = episodes; g_plot[PCL_P_ERROR] = plot.error;
    g_plot[PCL_P_ORDER_FRAME] = PCL_NEVER;
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
  }

  // ---- render (engine.py:737-759): backdrop, ' ', '^', '@', P
  uint8_t* board = p.out.d_board + (int64_t)env * H * pitch;
  const int segs_per_row = pitch >> 4;
  const int total = H * segs_per_row;
  for (int seg = lane; seg < total; seg += 32) {
    const int r = seg / segs_per_row;
    const int c0 = (seg - r * segs_per_row) << 4;
    uint4 px = *reinterpret_cast<const uint4*>(backdrop + r * pitch + c0);
    const u64 wave_r = s_rows[threadIdx.x >> 5][r];
    paint_bits(px, (unsigned)((row_bits(danger_bits, r, BW) >> c0) & 0xffffull), p.drape_char[1]);
    paint_bits(px, (unsigned)((row_bits(safe_bits, r, BW) >> c0) & 0xffffull), p.drape_char[2]);
    paint_bits(px, (unsigned)((wave_r >> c0) & 0xffffull), p.drape_char[0]);
    const unsigned m = sprite_bit(pl, r, c0);
    if (m) paint_bits(px, m, p.sprite_char[0]);
    *reinterpret_cast<uint4*>(board + r * pitch + c0) = px;
  }
}

}  // namespace

cudaError_t launch_shockwave(const StepParams& p, cudaStream_t s) {
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  shockwave_step<<<blocks, kWarpsPerBlock * 32, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pcl
