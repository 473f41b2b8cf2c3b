// apprehend.cu — fused step kernel for examples/apprehend.py:56-131.
//
// Two MazeWalkers over an empty backdrop, one update group ['b', 'P'], z-order 'bP'
// (ascii_art.py:184: the flat update schedule), neither has impassable characters, so no
// entity reads the board and a step is register arithmetic + the final render:
//   ball 'b' (sprite 1, not confined): _south every frame — frame 0 included, its
//     update() ignores `actions` — then x_accumulator += dx; below -0.5: _west and += 1.0,
//     above 0.5: _east and -= 1.0; virtual row >= H: reward -1 + terminate  (:109-131)
//   player 'P' (sprite 0, confined to the board): action 0 = _west, 1 = _east; on the
//     ball's virtual position: reward +1 + terminate                           (:76-87)
// The ball's slope dx = random.uniform(-2.499, 2.499) / (H - 1.0) is a float64 drawn from
// PYTHON's `random` when the sprite is built (:103), i.e. once per episode: with a per-env
// MT19937 state bound (pcl_state.d_rng, the words of random.Random(seed).getstate()) the
// kernel draws it at every (re)start exactly as random.uniform does — a + (b - a) *
// random(), random() = genrand_res53 — with correctly-rounded f64 operations only (no FMA
// contraction); without one (the single-env facade, whose Python sprite has already drawn)
// dx comes from the reset template.  dx lives in the ball's AUX0/AUX1 (f64 bits lo/hi), the
// accumulator in the plot's AUX0/AUX1.  One warp per env.
#include "pcl_device.cuh"
#include "pcl_kernels.cuh"
#include "pcl_mt.cuh"

namespace pcl {

namespace {

constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ double f64_of(int lo, int hi) {
  return __hiloint2double(hi, lo);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
apprehend_step(const StepParams p) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (env >= p.B) return;
  const int64_t lvl = p.st.d_level ? p.st.d_level[env] : env;
  const int H = p.H, W = p.W, pitch = p.pitch;
  int32_t* g_sprites = p.st.d_sprites + (int64_t)env * 2 * PCL_SPRITE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  const uint8_t* backdrop = p.st.d_backdrop + lvl * p.st.backdrop_bstride;

  const int was_over = g_plot[PCL_P_GAME_OVER];
  bool restart;
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) return;
  } else {
    restart = was_over && p.auto_reset;
    if (was_over && !p.auto_reset) return;           // reference raises; env stays frozen
  }
  const int32_t* src_s = restart ? p.st.d_sprites_init + lvl * p.st.sprites_init_bstride : g_sprites;
  const int32_t* src_p = restart ? p.st.d_plot_init + lvl * p.st.plot_init_bstride : g_plot;
  Sprite pl, ball;
  {
    const int32_t* r = src_s;
    pl.row = r[PCL_S_ROW]; pl.col = r[PCL_S_COL]; pl.vrow = r[PCL_S_VROW]; pl.vcol = r[PCL_S_VCOL];
    pl.flags = r[PCL_S_FLAGS]; pl.aux0 = pl.aux1 = pl.aux2 = 0;
    r += PCL_SPRITE_WORDS;
    ball.row = r[PCL_S_ROW]; ball.col = r[PCL_S_COL]; ball.vrow = r[PCL_S_VROW];
    ball.vcol = r[PCL_S_VCOL]; ball.flags = r[PCL_S_FLAGS];
    ball.aux0 = r[PCL_S_AUX0]; ball.aux1 = r[PCL_S_AUX1]; ball.aux2 = 0;
  }
  Plot plot;
  plot.frame = src_p[PCL_P_FRAME] + 1;                                // engine.py:716
  plot.error = g_plot[PCL_P_ERROR];
  plot.order_frame = PCL_NEVER; plot.order_r = plot.order_c = 0; plot.ego_mask = 0;
  const int episodes = g_plot[PCL_P_EPISODES] + (restart ? 1 : 0);
  double dx = f64_of(ball.aux0, ball.aux1);
  double acc = f64_of(src_p[PCL_P_AUX0], src_p[PCL_P_AUX1]);
  if (restart && p.st.d_rng != nullptr) {                             // BallSprite.__init__ :103
    uint32_t* mt = reinterpret_cast<uint32_t*>(p.st.d_rng) + (int64_t)env * PCL_MT_WORDS;
    const double r53 = mt_random53(mt, lane);
    const double u = __dadd_rn(-2.499, __dmul_rn(__dsub_rn(2.499, -2.499), r53));   // random.uniform
    dx = __ddiv_rn(u, __dsub_rn((double)H, 1.0));
    acc = 0.0;                                                         // :107
  }
  const int action = restart ? PCL_ACTION_NONE : p.actions[(int64_t)env * p.actions_per_env];
  Directives dir = fresh_directives();
  auto never_blocked = [](int, int) { return false; };

  // ---- BallSprite.update (:109-131), first in the group
  walker_move(ball, 1, PCL_M_S, plot, H, W, false, false, lane, never_blocked);
  acc = __dadd_rn(acc, dx);
  if (acc < -0.5) {
    walker_move(ball, 1, PCL_M_W, plot, H, W, false, false, lane, never_blocked);
    acc = __dadd_rn(acc, 1.0);
  } else if (acc > 0.5) {
    walker_move(ball, 1, PCL_M_E, plot, H, W, false, false, lane, never_blocked);
    acc = __dsub_rn(acc, 1.0);
  }
  if (ball.vrow >= H) { add_reward(dir, -1); terminate(dir); }
  // ---- PlayerSprite.update (:76-87)
  if (action == 0) walker_move(pl, 0, PCL_M_W, plot, H, W, true, false, lane, never_blocked);
  else if (action == 1) walker_move(pl, 0, PCL_M_E, plot, H, W, true, false, lane, never_blocked);
  if (pl.vrow == ball.vrow && pl.vcol == ball.vcol) { add_reward(dir, 1); terminate(dir); }

  __syncwarp();
  if (lane == 0) {
    int32_t* r = g_sprites;
    r[PCL_S_ROW] = pl.row; r[PCL_S_COL] = pl.col; r[PCL_S_VROW] = pl.vrow; r[PCL_S_VCOL] = pl.vcol;
    r[PCL_S_FLAGS] = pl.flags; r[PCL_S_AUX0] = 0; r[PCL_S_AUX1] = 0; r[PCL_S_AUX2] = 0;
    r += PCL_SPRITE_WORDS;
    r[PCL_S_ROW] = ball.row; r[PCL_S_COL] = ball.col; r[PCL_S_VROW] = ball.vrow;
    r[PCL_S_VCOL] = ball.vcol; r[PCL_S_FLAGS] = ball.flags;
    r[PCL_S_AUX0] = __double2loint(dx); r[PCL_S_AUX1] = __double2hiint(dx); r[PCL_S_AUX2] = 0;
    g_plot[PCL_P_FRAME] = plot.frame; g_plot[PCL_P_GAME_OVER] = dir.game_over;
    g_plot[PCL_P_EPISODES] = episodes; g_plot[PCL_P_ERROR] = plot.error;
    g_plot[PCL_P_ORDER_FRAME] = PCL_NEVER;
    g_plot[PCL_P_AUX0] = __double2loint(acc); g_plot[PCL_P_AUX1] = __double2hiint(acc);
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
  }

  // ---- render (engine.py:737-759): backdrop, 'b', then 'P' on top
  uint8_t* board = p.out.d_board + (int64_t)env * H * pitch;
  const int segs_per_row = pitch >> 4;
  const int total = H * segs_per_row;
  for (int seg = lane; seg < total; seg += 32) {
    const int r = seg / segs_per_row;
    const int c0 = (seg - r * segs_per_row) << 4;
    uint4 px = *reinterpret_cast<const uint4*>(backdrop + r * pitch + c0);
    unsigned m = sprite_bit(ball, r, c0);
    if (m) paint_bits(px, m, p.sprite_char[1]);
    m = sprite_bit(pl, r, c0);
    if (m) paint_bits(px, m, p.sprite_char[0]);
    *reinterpret_cast<uint4*>(board + r * pitch + c0) = px;
  }
}

}  // namespace

cudaError_t launch_apprehend(const StepParams& p, cudaStream_t s) {
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  apprehend_step<<<blocks, kWarpsPerBlock * 32, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pcl
