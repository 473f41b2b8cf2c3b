// warehouse.cu — fused step kernel for examples/warehouse_manager.py.
//
// Update groups [[boxes...], ['X'], ['P']] (warehouse_manager.py:168-170),
// z-order = the same sequence flattened (:176-178): boxes, then the JudgeDrape
// 'X' over boxes that sit on goals, then the player on top.
//
// Sprite order: boxes in update order, then P (index n_boxes).  Drape 0 = 'X';
// its curtain is never stored: after every JudgeDrape.update it is exactly
// {box cells} & (backdrop == '_') (:247-254), so one bit per box ("this box
// is drawn as X") carries it: box aux0.  Drape aux0 = _last_num_boxes_on_goals.
//
// Boards the entities read (engine.py:698-735): the boxes see the PREVIOUS
// step's final board (nothing has been re-rendered yet), the player sees the
// render after the judge ran.  Both are evaluated per cell on demand from the
// registers, one lane per looked-up cell.
#include "pcl_device.cuh"
#include "pcl_kernels.cuh"

namespace pcl {

namespace {

constexpr int kMaxS = 11;           // up to ten boxes + P
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ bool in_set(const uint32_t (&set)[4], int code) {
  return (set[(code >> 5) & 3] >> (code & 31)) & 1u;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
warehouse_step(const StepParams p) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (env >= p.B) return;
  const int H = p.H, W = p.W, S = p.S, NB = p.S - 1;

  int32_t* g_sprites = p.st.d_sprites + (int64_t)env * S * PCL_SPRITE_WORDS;
  int32_t* g_drapes = p.st.d_drapes + (int64_t)env * PCL_DRAPE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  const uint8_t* backdrop = p.st.d_backdrop + (int64_t)env * p.st.backdrop_bstride;

  Plot plot = load_record_rw<Plot>(g_plot);
  bool restart;
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) return;
  } else {
    restart = plot.game_over && p.auto_reset;
    if (plot.game_over && !p.auto_reset) return;
  }

  Sprite sp[kMaxS];
  Drape judge;
  int action;
  if (restart) {
    const int episodes = plot.episodes, error = plot.error;
    plot = load_record<Plot>(p.st.d_plot_init + (int64_t)env * p.st.plot_init_bstride);
    plot.episodes = episodes + 1;
    plot.error = error;
    const int32_t* si = p.st.d_sprites_init + (int64_t)env * p.st.sprites_init_bstride;
#pragma unroll
    for (int i = 0; i < kMaxS; ++i)
      if (i < S) sp[i] = load_record<Sprite>(si + i * PCL_SPRITE_WORDS);
    judge = load_record<Drape>(p.st.d_drapes_init + (int64_t)env * p.st.drapes_init_bstride);
    action = PCL_ACTION_NONE;
  } else {
#pragma unroll
    for (int i = 0; i < kMaxS; ++i)
      if (i < S) sp[i] = load_record_rw<Sprite>(g_sprites + i * PCL_SPRITE_WORDS);
    judge = load_record_rw<Drape>(g_drapes);
    action = p.actions[(int64_t)env * p.actions_per_env];
  }
  Directives dir = fresh_directives();
  plot.frame += 1;

  // Snapshot = the board every box reads (previous final render).
  int old_row[kMaxS], old_col[kMaxS], old_x[kMaxS];
#pragma unroll
  for (int i = 0; i < kMaxS; ++i) {
    old_row[i] = i < S ? sp[i].row : -1;
    old_col[i] = i < S ? sp[i].col : -1;
    old_x[i] = i < S ? sp[i].aux0 : 0;
  }
  Sprite player = sp[0];           // P lives in slot NB; pick it with unrolled selects
#pragma unroll
  for (int i = 1; i < kMaxS; ++i) if (i == NB) player = sp[i];
  const bool pl_vis = visible(player);
  const int pl_row = player.row, pl_col = player.col;

  // Character shown by the stale board at (r, c).
  auto stale_cell = [&](int r, int c) -> int {
    if (pl_vis && r == pl_row && c == pl_col) return p.sprite_char[NB];
    int code = backdrop[(int64_t)r * p.pitch + c];
#pragma unroll
    for (int i = 0; i < kMaxS - 1; ++i) {
      if (i < NB && old_row[i] == r && old_col[i] == c)   // later boxes paint over earlier
        code = old_x[i] ? 'X' : p.sprite_char[i];
    }
    return code;
  };

  // ---- group 0: boxes (BoxSprite.update, warehouse_manager.py:208-226)
  if (action >= 0 && action <= 3) {
    // layers['P'][rows+-1, cols+-1] with NumPy index rules.
    const int dr = action == 0 ? 1 : action == 1 ? -1 : 0;
    const int dc = action == 2 ? 1 : action == 3 ? -1 : 0;
    const int motion = action == 0 ? PCL_M_N : action == 1 ? PCL_M_S
                     : action == 2 ? PCL_M_W : PCL_M_E;
#pragma unroll
    for (int i = 0; i < kMaxS - 1; ++i) {
      if (i < NB) {
        int rr = sp[i].row + dr, cc = sp[i].col + dc;
        if (rr < 0) rr += H;
        if (cc < 0) cc += W;
        bool pushed = false;
        if (rr >= H || cc >= W) plot.error |= PCL_ENV_ERR_INDEX;
        else pushed = pl_vis && rr == pl_row && cc == pl_col;
        if (pushed) {
          const uint32_t (&imp)[4] = p.impassable[i];
          walker_move(sp[i], i, motion, plot, H, W, false, false, lane,
                      [&](int r, int c) { return in_set(imp, stale_cell(r, c)); });
        }
      }
    }
  }

  // ---- group 1: JudgeDrape.update (:245-266)
  int num_boxes = 0, on_goals = 0;
#pragma unroll
  for (int i = 0; i < kMaxS - 1; ++i) {
    if (i < NB) {
      bool first = true;
#pragma unroll
      for (int j = 0; j < kMaxS - 1; ++j)
        if (j < i && sp[j].row == sp[i].row && sp[j].col == sp[i].col) first = false;
      const bool goal = backdrop[(int64_t)sp[i].row * p.pitch + sp[i].col] == '_';
      sp[i].aux0 = goal ? 1 : 0;
      num_boxes += first ? 1 : 0;
      on_goals += (first && goal) ? 1 : 0;
    }
  }
  add_reward(dir, on_goals - judge.aux0);
  judge.aux0 = on_goals;
  if (action == 5 || on_goals == num_boxes) terminate(dir);

  // ---- group 2: PlayerSprite.update (:284-295), board = boxes moved + X redrawn
  if (action >= 0 && action <= 3) {
    const int motion = action == 0 ? PCL_M_N : action == 1 ? PCL_M_S
                     : action == 2 ? PCL_M_W : PCL_M_E;
    const uint32_t (&imp)[4] = p.impassable[NB];
    auto now_cell = [&](int r, int c) -> int {
      if (pl_vis && r == pl_row && c == pl_col) return p.sprite_char[NB];
      int code = backdrop[(int64_t)r * p.pitch + c];
#pragma unroll
      for (int i = 0; i < kMaxS - 1; ++i)
        if (i < NB && visible(sp[i]) && sp[i].row == r && sp[i].col == c)
          code = sp[i].aux0 ? 'X' : p.sprite_char[i];
      return code;
    };
    walker_move(player, NB, motion, plot, H, W, false, false, lane,
                [&](int r, int c) { return in_set(imp, now_cell(r, c)); });
  }

  plot.game_over = dir.game_over;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kMaxS - 1; ++i)
      if (i < NB) store_record(g_sprites + i * PCL_SPRITE_WORDS, sp[i]);
    store_record(g_sprites + NB * PCL_SPRITE_WORDS, player);
    store_record(g_drapes, judge);
    store_record(g_plot, plot);
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
  }

  // ---- final render: backdrop, boxes (as 'X' on goals), player.
  uint8_t* board = p.out.d_board + (int64_t)env * H * p.pitch;
  const int segs_per_row = p.pitch >> 4;
  const int total = H * segs_per_row;
  for (int seg = lane; seg < total; seg += 32) {
    const int r = seg / segs_per_row;
    const int c0 = (seg - r * segs_per_row) << 4;
    uint4 px = __ldg(reinterpret_cast<const uint4*>(backdrop + (int64_t)r * p.pitch + c0));
#pragma unroll
    for (int i = 0; i < kMaxS - 1; ++i)
      if (i < NB)
        paint_bits(px, sprite_bit(sp[i], r, c0), sp[i].aux0 ? 'X' : p.sprite_char[i]);
    paint_bits(px, sprite_bit(player, r, c0), p.sprite_char[NB]);
    *reinterpret_cast<uint4*>(board + (int64_t)r * p.pitch + c0) = px;
  }
}

}  // namespace

cudaError_t launch_warehouse(const StepParams& p, cudaStream_t s) {
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  warehouse_step<<<blocks, kWarpsPerBlock * 32, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pcl
