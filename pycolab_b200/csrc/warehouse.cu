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
// records + the backdrop tile.
//
// Memory schedule (one warp per env): records -> smem with coalesced loads; the
// whole backdrop tile -> smem with cp.async (it is both the board's base layer
// and the only thing the look-ups need); all game logic then runs out of
// shared memory with one lane per box; the board is the staged tile streamed
// back out with uint4 stores plus <= 11 single-byte patches for the sprites.
#include "pcl_device.cuh"
#include "pcl_kernels.cuh"

namespace pcl {

namespace {

constexpr int kMaxS = 11;           // up to ten boxes + P
constexpr int kWarpsPerBlock = 4;
constexpr int kRecWords = 128;      // 11 * 8 sprite words + 8 drape + 16 plot, padded

__device__ __forceinline__ bool in_set(const uint32_t* set, int code) {
  return (set[(code >> 5) & 3] >> (code & 31)) & 1u;
}
__device__ __forceinline__ int motion_of_action(int a) {    // :214-226, :288-295
  return a == 0 ? PCL_M_N : a == 1 ? PCL_M_S : a == 2 ? PCL_M_W : PCL_M_E;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 8)
warehouse_step(const StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerBlock + warp;
  if (env >= p.B) return;
  const int64_t lvl = p.st.d_level ? p.st.d_level[env] : env;   // index of static level data
  const int H = p.H, W = p.W, S = p.S, NB = p.S - 1, pitch = p.pitch;
  const size_t tile = (size_t)H * pitch;
  uint8_t* my = smem_raw + warp * (kRecWords * 4 + tile);
  int32_t* rec = reinterpret_cast<int32_t*>(my);          // [0, 88) sprites, [96, 104) drape, [112, 128) plot
  uint8_t* s_bd = my + kRecWords * 4;
  int32_t* r_judge = rec + 96;
  int32_t* r_plot = rec + 112;

  int32_t* g_sprites = p.st.d_sprites + (int64_t)env * S * PCL_SPRITE_WORDS;
  int32_t* g_drapes = p.st.d_drapes + (int64_t)env * PCL_DRAPE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  const uint8_t* backdrop = p.st.d_backdrop + lvl * p.st.backdrop_bstride;

  // ---- the backdrop tile does not depend on anything: start it first, as one
  // TMA bulk copy tracked by this warp's mbarrier (rec words 104..105 are padding).
  uint64_t* bar = reinterpret_cast<uint64_t*>(rec + 104);
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  if (lane == 0) tile_load_bulk(s_bd, backdrop, (uint32_t)tile, bar);
  // ---- records -> smem: the live state first, unconditionally, in one round
  // trip; whether the env restarts is decided from the staged plot record.
  for (int i = lane; i < S * PCL_SPRITE_WORDS; i += 32) rec[i] = g_sprites[i];
  if (lane < PCL_DRAPE_WORDS) r_judge[lane] = g_drapes[lane];
  if (lane >= 16) r_plot[lane - 16] = g_plot[lane - 16];
  int action = p.mode == MODE_STEP ? p.actions[(int64_t)env * p.actions_per_env] : PCL_ACTION_NONE;
  __syncwarp();
  const int was_over = r_plot[PCL_P_GAME_OVER];
  bool restart;
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) { mbar_wait(bar, 0); return; }
  } else {
    restart = was_over && p.auto_reset;
    if (was_over && !p.auto_reset) { mbar_wait(bar, 0); return; }
  }
  if (restart) {                               // a fresh Engine: templates over the live state
    const int episodes = r_plot[PCL_P_EPISODES], error = r_plot[PCL_P_ERROR];
    __syncwarp();
    const int32_t* ss = p.st.d_sprites_init + lvl * p.st.sprites_init_bstride;
    const int32_t* sd = p.st.d_drapes_init + lvl * p.st.drapes_init_bstride;
    const int32_t* sp = p.st.d_plot_init + lvl * p.st.plot_init_bstride;
    for (int i = lane; i < S * PCL_SPRITE_WORDS; i += 32) rec[i] = ss[i];
    if (lane < PCL_DRAPE_WORDS) r_judge[lane] = sd[lane];
    if (lane >= 16) r_plot[lane - 16] = sp[lane - 16];
    action = PCL_ACTION_NONE;
    __syncwarp();
    if (lane == 0) { r_plot[PCL_P_EPISODES] = episodes + 1; r_plot[PCL_P_ERROR] = error; }
  }
  mbar_wait(bar, 0);
  __syncwarp();

  Plot plot;
  plot.frame = r_plot[PCL_P_FRAME] + 1;                    // engine.py:716
  plot.error = r_plot[PCL_P_ERROR];
  plot.order_frame = PCL_NEVER; plot.order_r = 0; plot.order_c = 0; plot.ego_mask = 0;
  Directives dir = fresh_directives();

  // Player as of the previous render; box i's record sits in lane i.
  Sprite player;
  {
    const int32_t* r = rec + NB * PCL_SPRITE_WORDS;
    player.row = r[PCL_S_ROW]; player.col = r[PCL_S_COL];
    player.vrow = r[PCL_S_VROW]; player.vcol = r[PCL_S_VCOL];
    player.flags = r[PCL_S_FLAGS]; player.aux0 = player.aux1 = player.aux2 = 0;
  }
  const bool pl_vis = visible(player);
  const int pl_row = player.row, pl_col = player.col;
  // Sprite characters -> smem (rec words 88..95 are padding); static indices only,
  // so the parameter block is never copied to local memory.
  uint8_t* s_chars = reinterpret_cast<uint8_t*>(rec + 88);
  {
    int ch = 0;
#pragma unroll
    for (int i = 0; i < kMaxS; ++i) if (i == lane) ch = p.sprite_char[i];
    if (lane < kMaxS) s_chars[lane] = (uint8_t)ch;
    __syncwarp();
  }
  const int P_CHAR = s_chars[NB];
  const bool is_box = lane < NB;
  const int32_t* mine = rec + (is_box ? lane : 0) * PCL_SPRITE_WORDS;
  int b_row = mine[PCL_S_ROW], b_col = mine[PCL_S_COL];

  // Character of the stale board (= previous final render) at (r, c): P on top,
  // then a box (drawn 'X' when the judge marked it), else the backdrop.
  // Box positions are still the old ones in `rec` while this is used.
  auto stale_cell = [&](int r, int c) -> int {
    if (pl_vis && r == pl_row && c == pl_col) return P_CHAR;
    int code = s_bd[r * pitch + c];
    for (int i = 0; i < NB; ++i) {
      const int32_t* b = rec + i * PCL_SPRITE_WORDS;
      if (b[PCL_S_ROW] == r && b[PCL_S_COL] == c) code = b[PCL_S_AUX0] ? 'X' : s_chars[i];
    }
    return code;
  };

  // ---- group 0: boxes (BoxSprite.update, warehouse_manager.py:208-226)
  if (action >= 0 && action <= 3) {
    // layers['P'][rows +- 1, cols +- 1] with NumPy index rules: -1 wraps, >= size raises.
    const int dr = action == 0 ? 1 : action == 1 ? -1 : 0;
    const int dc = action == 2 ? 1 : action == 3 ? -1 : 0;
    int rr = b_row + dr, cc = b_col + dc;
    if (rr < 0) rr += H;
    if (cc < 0) cc += W;
    const bool oob = is_box && (rr >= H || cc >= W);
    const bool pushed = is_box && !oob && pl_vis && rr == pl_row && cc == pl_col;
    if (__any_sync(PCL_FULL, oob)) plot.error |= PCL_ENV_ERR_INDEX;
    const unsigned who = __ballot_sync(PCL_FULL, pushed);
    if (who) {                               // at most one box can be next to P
      const int j = __ffs(who) - 1;
      const int32_t* r = rec + j * PCL_SPRITE_WORDS;
      Sprite box;
      box.row = r[PCL_S_ROW]; box.col = r[PCL_S_COL];
      box.vrow = r[PCL_S_VROW]; box.vcol = r[PCL_S_VCOL];
      box.flags = r[PCL_S_FLAGS]; box.aux0 = r[PCL_S_AUX0]; box.aux1 = box.aux2 = 0;
      uint32_t imp[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) imp[w] = p.impassable[0][w];
#pragma unroll
      for (int i = 1; i < kMaxS - 1; ++i)
        if (i == j) {
#pragma unroll
          for (int w = 0; w < 4; ++w) imp[w] = p.impassable[i][w];
        }
      walker_move(box, j, motion_of_action(action), plot, H, W, false, false, lane,
                  [&](int r2, int c2) { return in_set(imp, stale_cell(r2, c2)); });
      __syncwarp();
      if (lane == 0) {
        int32_t* w = rec + j * PCL_SPRITE_WORDS;
        w[PCL_S_ROW] = box.row; w[PCL_S_COL] = box.col;
        w[PCL_S_VROW] = box.vrow; w[PCL_S_VCOL] = box.vcol; w[PCL_S_FLAGS] = box.flags;
      }
      __syncwarp();
      if (lane == j) { b_row = box.row; b_col = box.col; }
    }
  }

  // ---- group 1: JudgeDrape.update (:245-266), one lane per box
  bool first = is_box, goal = false;
  if (is_box) {
    for (int i = 0; i < lane; ++i) {
      const int32_t* b = rec + i * PCL_SPRITE_WORDS;
      if (b[PCL_S_ROW] == b_row && b[PCL_S_COL] == b_col) first = false;
    }
    goal = s_bd[b_row * pitch + b_col] == '_';
  }
  const int num_boxes = __popc(__ballot_sync(PCL_FULL, first));
  const int on_goals = __popc(__ballot_sync(PCL_FULL, first && goal));
  if (is_box) rec[lane * PCL_SPRITE_WORDS + PCL_S_AUX0] = goal ? 1 : 0;
  add_reward(dir, on_goals - r_judge[PCL_D_AUX0]);
  if (action == 5 || on_goals == num_boxes) terminate(dir);
  __syncwarp();

  // ---- group 2: PlayerSprite.update (:284-295); board = boxes moved, X redrawn
  if (action >= 0 && action <= 3) {
    uint32_t imp[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) imp[w] = p.impassable[kMaxS - 1][w];
#pragma unroll
    for (int i = 1; i < kMaxS - 1; ++i)
      if (i == NB) {
#pragma unroll
        for (int w = 0; w < 4; ++w) imp[w] = p.impassable[i][w];
      }
    walker_move(player, NB, motion_of_action(action), plot, H, W, false, false, lane,
                [&](int r2, int c2) { return in_set(imp, stale_cell(r2, c2)); });
  }

  // ---- _apply_and_clear_plot + records back
  __syncwarp();
  if (lane == 0) {
    int32_t* w = rec + NB * PCL_SPRITE_WORDS;
    w[PCL_S_ROW] = player.row; w[PCL_S_COL] = player.col;
    w[PCL_S_VROW] = player.vrow; w[PCL_S_VCOL] = player.vcol; w[PCL_S_FLAGS] = player.flags;
    r_judge[PCL_D_AUX0] = on_goals;
    r_plot[PCL_P_FRAME] = plot.frame; r_plot[PCL_P_GAME_OVER] = dir.game_over;
    r_plot[PCL_P_ERROR] = plot.error;
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
  }
  __syncwarp();
  for (int i = lane; i < S * PCL_SPRITE_WORDS; i += 32) g_sprites[i] = rec[i];
  if (lane < PCL_DRAPE_WORDS) g_drapes[lane] = r_judge[lane];
  if (lane >= 16) g_plot[lane - 16] = r_plot[lane - 16];

  // ---- final render: patch the sprite cells into the staged tile, stream it out.
  // Boxes never share a cell with each other or with P (each is impassable to
  // the others), so the patches are independent; P goes last to keep z-order.
  if (is_box && (mine[PCL_S_FLAGS] & 1))
    s_bd[b_row * pitch + b_col] = goal ? 'X' : s_chars[lane];
  __syncwarp();
  if (lane == 0 && visible(player)) s_bd[player.row * pitch + player.col] = (uint8_t)P_CHAR;
  __syncwarp();
  uint8_t* board = p.out.d_board + (int64_t)env * tile;
  tile_store_fence();
  __syncwarp();
  if (lane == 0) {
    tile_store_bulk(board, s_bd, (uint32_t)tile);
    tile_store_wait();                       // the tile must outlive the copy's reads
  }
}

}  // namespace

cudaError_t launch_warehouse(const StepParams& p, cudaStream_t s) {
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const size_t smem = (kRecWords * 4 + (size_t)p.H * p.pitch) * kWarpsPerBlock;
  if (smem > 48 * 1024) {   // opt in per launch: the attribute is per device, handles are not
    cudaError_t e = cudaFuncSetAttribute(warehouse_step,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  warehouse_step<<<blocks, kWarpsPerBlock * 32, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pcl
