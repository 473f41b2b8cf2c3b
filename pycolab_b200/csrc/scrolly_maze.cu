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
// it.
//
// Memory schedule (one warp per env, everything staged through shared memory):
//   1. records (sprites/drapes/plot, 64 words) -> smem with two coalesced loads;
//      only the fields this game uses are pulled into registers;
//   2. group 0 ('#' MazeDrape) is pure register arithmetic and fixes BOTH
//      final window corners (the '@' drape can only obey an order, never issue
//      one: by the time it runs, the player's permit is already for frame+1);
//   3. one batch of loads is issued for everything else the step will read:
//        - cp.async: the backdrop tile (issued before anything else) and the
//          two windows of the bit-packed patterns (4 words per row, two 8-byte
//          copies) -> smem, no registers held;
//        - plain loads: a 5x5 patch of wall bits around each of the 4 walkers
//          (covers every cell any _check_motion of this step can consult,
//          wherever the scroll order moves the walker first) and the 3x3 patch
//          of coin bits around the player, one cell per lane, 4 per lane;
//   4. groups 1 and 2 run on registers + ballots of those bits;
//   5. each lane shifts whole window rows once into one word per 16-cell board
//      segment (wall16 << 16 | coin16); the paint loop then composes 16-byte
//      segments from smem (prmt with a 256-entry selector table) and streams them
//      out with uint4 stores; records go back with two coalesced stores.
// So a step costs ~two dependent DRAM round trips (records, then everything).
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
constexpr int kRecWords = 64;     // 4 sprites * 8 + 2 drapes * 8 + plot 16

__device__ __forceinline__ int action_to_motion(int a) {   // scrolly_maze.py:262-271
  return a == 0 ? PCL_M_N : a == 1 ? PCL_M_S : a == 2 ? PCL_M_W
       : a == 3 ? PCL_M_E : a == 4 ? PCL_M_STAY : PCL_M_NONE;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::
               "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::
               "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}

// A 64-cell window row starts at bit corner_c of its pattern row; the four words
// from the even word at or below corner_c >> 5 always cover it (<= 31 + 32 + 64
// bits), and pattern rows are 8-byte aligned (pattern_words is even), so a row is
// staged with two 8-byte cp.async into a 16-byte smem slot.

// prmt.b32 without __byte_perm's selector masking (the table holds nibbles 0..5).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

__host__ __device__ __forceinline__ size_t warp_smem_bytes(int H, int pitch) {
  // records, backdrop tile, two 4-word window rows per board row, one word per segment
  return kRecWords * 4 + (size_t)H * pitch + 2 * ((size_t)H * 16) +
         (((size_t)H * (pitch >> 2) + 15) & ~(size_t)15);      // keep every warp's slice 16-byte aligned
}

// Programmatic dependent launch: let the next kernel of the stream begin its
// launch/prologue while this one runs, and wait for everything earlier in the
// stream before touching global memory.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grids() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 7)
scrolly_maze_step(const StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // Byte-permute selectors for 4 cells at once: index = wall nibble << 4 | coin
  // nibble; selector nibble k picks byte 5 ('#') if wall_k, else byte 4 ('@') if
  // coin_k, else byte k of the backdrop word (z-order ... '@' '#' ...).
  __shared__ uint16_t s_sel[256];
  for (int idx = threadIdx.x; idx < 256; idx += blockDim.x) {
    unsigned sel = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned nib = ((idx >> (4 + k)) & 1) ? 5u : ((idx >> k) & 1) ? 4u : (unsigned)k;
      sel |= nib << (4 * k);
    }
    s_sel[idx] = (uint16_t)sel;
  }
  pdl_launch_dependents();
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerBlock + warp;
  if (env >= p.B) return;
  const int H = p.H, W = p.W, PWW = p.PWW;

  uint8_t* my = smem_raw + warp * warp_smem_bytes(H, p.pitch);
  int32_t* rec = reinterpret_cast<int32_t*>(my);
  uint8_t* s_bd = my + kRecWords * 4;
  uint32_t* s_wall = reinterpret_cast<uint32_t*>(s_bd + (size_t)H * p.pitch);
  uint32_t* s_coin = s_wall + H * 4;
  // Everything above ran without touching global memory; from here on the
  // kernel reads state earlier work in the stream may have produced.
  pdl_wait_prior_grids();
  const int64_t lvl = p.st.d_level ? p.st.d_level[env] : env;   // index of static level data

  int32_t* g_sprites = p.st.d_sprites + (int64_t)env * kS * PCL_SPRITE_WORDS;
  int32_t* g_drapes = p.st.d_drapes + (int64_t)env * 2 * PCL_DRAPE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  const uint32_t* wall_pat = p.st.d_pattern[0] + lvl * p.st.pattern_bstride[0];
  uint32_t* coin_pat = p.st.d_pattern[1] + (int64_t)env * p.st.pattern_bstride[1];
  const uint8_t* backdrop = p.st.d_backdrop + lvl * p.st.backdrop_bstride;

  // ---- 0. the backdrop tile depends on nothing: get it moving first --------
  {
    const int n16 = (H * p.pitch) >> 4;
    for (int i = lane; i < n16; i += 32) cp_async16(s_bd + i * 16, backdrop + i * 16);
  }
  // ---- 1. records -> smem (coalesced) ------------------------------------
  rec[lane] = g_sprites[lane];
  rec[32 + lane] = lane < 16 ? g_drapes[lane] : g_plot[lane - 16];
  __syncwarp();
  const int was_over = rec[48 + PCL_P_GAME_OVER];
  bool restart;                              // engine.py:520-581, 619-624
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) { cp_async_wait_all(); return; }
  } else {
    restart = was_over && p.auto_reset;
    if (was_over && !p.auto_reset) {         // reference raises; env stays frozen
      cp_async_wait_all();
      return;
    }
  }
  int action;
  if (restart) {
    const int episodes = rec[48 + PCL_P_EPISODES], error = rec[48 + PCL_P_ERROR];
    __syncwarp();
    const int32_t* si = p.st.d_sprites_init + lvl * p.st.sprites_init_bstride;
    const int32_t* di = p.st.d_drapes_init + lvl * p.st.drapes_init_bstride;
    const int32_t* pi = p.st.d_plot_init + lvl * p.st.plot_init_bstride;
    rec[lane] = __ldg(si + lane);
    rec[32 + lane] = lane < 16 ? __ldg(di + lane) : __ldg(pi + lane - 16);
    // Fresh coins: restore the mutable pattern (one Engine per episode).
    const uint32_t* src = p.st.d_pattern_init[1] + lvl * p.st.pattern_init_bstride[1];
    const int n = p.PH * PWW;
    for (int i = lane; i < n; i += 32) coin_pat[i] = __ldg(src + i);
    __syncwarp();
    if (lane == 0) { rec[48 + PCL_P_EPISODES] = episodes + 1; rec[48 + PCL_P_ERROR] = error; }
    __syncwarp();
    action = PCL_ACTION_NONE;
  } else {
    action = p.actions[(int64_t)env * p.actions_per_env];
  }

  // ---- only the fields this game uses live in registers ------------------
  Sprite sp[kS];
#pragma unroll
  for (int i = 0; i < kS; ++i) {
    const int32_t* r = rec + i * PCL_SPRITE_WORDS;
    sp[i].row = r[PCL_S_ROW]; sp[i].col = r[PCL_S_COL];
    sp[i].vrow = r[PCL_S_VROW]; sp[i].vcol = r[PCL_S_VCOL];
    sp[i].flags = r[PCL_S_FLAGS]; sp[i].aux0 = r[PCL_S_AUX0];
    sp[i].aux1 = (i == 0) ? r[PCL_S_AUX1] : 0; sp[i].aux2 = 0;
  }
  Drape walls, coins;
  {
    const int32_t* r = rec + 32;
    walls.corner_r = r[PCL_D_CORNER_R]; walls.corner_c = r[PCL_D_CORNER_C];
    walls.pre_r = r[PCL_D_PRE_R]; walls.pre_c = r[PCL_D_PRE_C];
    walls.last_frame = r[PCL_D_LAST_FRAME];
    r += PCL_DRAPE_WORDS;
    coins.corner_r = r[PCL_D_CORNER_R]; coins.corner_c = r[PCL_D_CORNER_C];
    coins.pre_r = r[PCL_D_PRE_R]; coins.pre_c = r[PCL_D_PRE_C];
    coins.last_frame = r[PCL_D_LAST_FRAME];
    coins.aux0 = r[PCL_D_AUX0]; coins.aux1 = r[PCL_D_AUX1];
  }
  Plot plot;
  {
    const int32_t* r = rec + 48;
    plot.frame = r[PCL_P_FRAME]; plot.error = r[PCL_P_ERROR];
    plot.order_r = r[PCL_P_ORDER_R]; plot.order_c = r[PCL_P_ORDER_C];
    plot.order_frame = r[PCL_P_ORDER_FRAME]; plot.ego_mask = r[PCL_P_EGO_MASK];
    plot.aux0 = r[PCL_P_AUX0];
  }

  const ScrollyCfg wcfg = scrolly_cfg(H, W, p.PH, p.PW, p.margin[0][0], p.margin[0][1]);
  const ScrollyCfg ccfg = scrolly_cfg(H, W, p.PH, p.PW, p.margin[1][0], p.margin[1][1]);
  Directives dir = fresh_directives();
  const int motion = action_to_motion(action);

  plot.frame += 1;                                           // engine.py:716

  // ---- 2. update group 0: '#' MazeDrape (scrolly_maze.py:308-329) --------
  if (motion != PCL_M_NONE) scrolly_move(walls, wcfg, motion, plot, sp);
  const bool ordered = plot.order_frame == plot.frame;
  const int wr = walls.corner_r, wc = walls.corner_c;
  // Where the '@' window will be after it obeys the same order (checked below).
  const int cr_pred = coins.corner_r + (ordered && motion != PCL_M_NONE ? plot.order_r : 0);
  const int cc_pred = coins.corner_c + (ordered && motion != PCL_M_NONE ? plot.order_c : 0);

  // ---- 3. one batch of loads ---------------------------------------------
  const int we = (wc >> 5) & ~1, ce = (cc_pred >> 5) & ~1;   // first staged word (even)
  {
    const int nhalf = H * 2;                 // two 8-byte halves per window row
    for (int i = lane; i < nhalf; i += 32) {
      const int r = i >> 1, k = (i & 1) * 2;
      cp_async8(s_wall + i * 2, wall_pat + (int64_t)(wr + r) * PWW + we + k);
      cp_async8(s_coin + i * 2, coin_pat + (int64_t)(cr_pred + r) * PWW + ce + k);
    }
  }
  // Walker start positions: every look-up below is relative to these.
  int vr0[kS], vc0[kS];
#pragma unroll
  for (int i = 0; i < kS; ++i) { vr0[i] = sp[i].vrow; vc0[i] = sp[i].vcol; }
  scrolly_touch_prescroll(coins, plot);      // '@' has not moved yet this frame
  unsigned patch[4];                         // 128 look-up bits, one per (lane, round)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int j = q * 32 + lane;
    bool bit = false;
    if (j < 100) {                           // wall patch of walker j / 25
      const int i = j / 25, k = j - i * 25;
      int vr = vr0[0], vc = vc0[0];
#pragma unroll
      for (int t = 1; t < kS; ++t) if (i == t) { vr = vr0[t]; vc = vc0[t]; }
      const int pr = wr + vr + k / 5 - 2, pc = wc + vc + k % 5 - 2;
      if ((unsigned)pr < (unsigned)p.PH && (unsigned)pc < (unsigned)p.PW)
        bit = bit_at(wall_pat + (int64_t)pr * PWW, pc);
    } else if (j < 110) {                    // coin patch around the player
      const int k = j - 100;
      int r = k < 9 ? vr0[0] + k / 3 - 1 : 0, c = k < 9 ? vc0[0] + k % 3 - 1 : 0;
      if (on_board(r, c, H, W))
        bit = bit_at(coin_pat + (int64_t)(coins.pre_r + r) * PWW, coins.pre_c + c);
    }
    patch[q] = __ballot_sync(PCL_FULL, bit);
  }
  unsigned wall5[kS];                        // 25-bit 5x5 patches
  wall5[0] = patch[0] & 0x1ffffffu;
  wall5[1] = __funnelshift_r(patch[0], patch[1], 25) & 0x1ffffffu;
  wall5[2] = __funnelshift_r(patch[1], patch[2], 18) & 0x1ffffffu;
  wall5[3] = __funnelshift_r(patch[2], patch[3], 11) & 0x1ffffffu;
  const unsigned coin9 = (patch[3] >> 4) & 0x3ffu;     // 3x3 + the (0,0) cell

  // Board of render #1 as far as MazeWalkers care: is cell (r, c) a '#'?
  // P is still painted where the previous render put it.
  const bool p_vis = visible(sp[0]);
  const int p_row = sp[0].row, p_col = sp[0].col;

  // ---- 4a. update group 1: patrollers a, b, c then P ---------------------
  const int p_vrow = sp[0].vrow, p_vcol = sp[0].vcol;   // P moves after them
#pragma unroll
  for (int i = 0; i < kS; ++i) {
    const int idx = (i + 1) & 3;             // order a, b, c, P = 1, 2, 3, 0
    const unsigned field = wall5[idx];
    const int r0 = vr0[idx], c0 = vc0[idx];
    auto is_wall = [&](int r, int c) -> bool {
      if (p_vis && r == p_row && c == p_col) return false;
      const int dr = r - r0 + 2, dc = c - c0 + 2;
      if ((unsigned)dr > 4u || (unsigned)dc > 4u) return false;   // cannot happen
      return (field >> (dr * 5 + dc)) & 1u;
    };
    if (idx != 0) {                          // PatrollerSprite :284-305
      if (plot.frame % 2) {
        walker_move(sp[idx], idx, PCL_M_STAY, plot, H, W, false, false, lane, is_wall);
      } else {
        scrolly_touch_prescroll(walls, plot);
        const int step = sp[idx].aux0 ? 1 : -1;
        int pr = r0 + walls.pre_r, pc = c0 + walls.pre_c + step;
        bool next_to_wall = false;
        if ((unsigned)pr < (unsigned)p.PH && (unsigned)pc < (unsigned)p.PW) {
          // Same cell seen from the post-scroll corner: inside the 5x5 patch.
          const int dr = pr - wr - r0 + 2, dc = pc - wc - c0 + 2;
          next_to_wall = (field >> (dr * 5 + dc)) & 1u;
        } else {
          // NumPy indexing: negatives wrap once, anything else is an IndexError.
          if (pr < 0) pr += p.PH;
          if (pc < 0) pc += p.PW;
          if ((unsigned)pr < (unsigned)p.PH && (unsigned)pc < (unsigned)p.PW)
            next_to_wall = bit_at(wall_pat + (int64_t)pr * PWW, pc);
          else
            plot.error |= PCL_ENV_ERR_INDEX;
        }
        if (next_to_wall) sp[idx].aux0 = !sp[idx].aux0;
        walker_move(sp[idx], idx, sp[idx].aux0 ? PCL_M_E : PCL_M_W, plot, H, W, false,
                    false, lane, is_wall);
        if (sp[idx].vrow == p_vrow && sp[idx].vcol == p_vcol) terminate(dir);
      }
    } else if (motion != PCL_M_NONE) {       // PlayerSprite :258-271
      walker_move(sp[0], 0, motion, plot, H, W, false, true, lane, is_wall);
    }
  }

  // ---- 4b. update group 2: '@' CashDrape (scrolly_maze.py:341-364) -------
  int picked_r = -1, picked_c = -1;          // pattern cell cleared this frame
  {
    const int dr = sp[0].row - vr0[0], dc = sp[0].col - vc0[0];
    bool coin;
    const int pr = coins.pre_r + sp[0].row, pc = coins.pre_c + sp[0].col;
    if (sp[0].row == 0 && sp[0].col == 0 && !on_board(sp[0].vrow, sp[0].vcol, H, W))
      coin = (coin9 >> 9) & 1u;              // off-board player sits at (0, 0)
    else if ((unsigned)(dr + 1) <= 2u && (unsigned)(dc + 1) <= 2u)
      coin = (coin9 >> ((dr + 1) * 3 + dc + 1)) & 1u;
    else
      coin = bit_at(coin_pat + (int64_t)pr * PWW, pc);   // cannot happen
    if (coin) {
      add_reward(dir, 100);
      if (lane == 0) coin_pat[(int64_t)pr * PWW + (pc >> 5)] &= ~(1u << (pc & 31));
      picked_r = pr; picked_c = pc;
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
  cp_async_wait_all();
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kS; ++i) {
      int32_t* r = rec + i * PCL_SPRITE_WORDS;
      r[PCL_S_ROW] = sp[i].row; r[PCL_S_COL] = sp[i].col;
      r[PCL_S_VROW] = sp[i].vrow; r[PCL_S_VCOL] = sp[i].vcol;
      r[PCL_S_FLAGS] = sp[i].flags; r[PCL_S_AUX0] = sp[i].aux0;
      if (i == 0) r[PCL_S_AUX1] = sp[i].aux1;
    }
    int32_t* r = rec + 32;
    r[PCL_D_CORNER_R] = walls.corner_r; r[PCL_D_CORNER_C] = walls.corner_c;
    r[PCL_D_PRE_R] = walls.pre_r; r[PCL_D_PRE_C] = walls.pre_c;
    r[PCL_D_LAST_FRAME] = walls.last_frame;
    r += PCL_DRAPE_WORDS;
    r[PCL_D_CORNER_R] = coins.corner_r; r[PCL_D_CORNER_C] = coins.corner_c;
    r[PCL_D_PRE_R] = coins.pre_r; r[PCL_D_PRE_C] = coins.pre_c;
    r[PCL_D_LAST_FRAME] = coins.last_frame;
    r[PCL_D_AUX0] = coins.aux0; r[PCL_D_AUX1] = coins.aux1;
    r = rec + 48;
    r[PCL_P_FRAME] = plot.frame; r[PCL_P_GAME_OVER] = dir.game_over;
    r[PCL_P_ERROR] = plot.error;
    r[PCL_P_ORDER_R] = plot.order_r; r[PCL_P_ORDER_C] = plot.order_c;
    r[PCL_P_ORDER_FRAME] = plot.order_frame; r[PCL_P_EGO_MASK] = plot.ego_mask;
    r[PCL_P_AUX0] = plot.aux0;
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
    // The coin window was staged before the pick-up: clear the bit there too.
    if (picked_r >= 0) {
      const int r2 = picked_r - cr_pred, b = picked_c - (ce << 5);
      if ((unsigned)r2 < (unsigned)H && (unsigned)b < 128u)
        s_coin[r2 * 4 + (b >> 5)] &= ~(1u << (b & 31));
    }
  }
  const int cr = coins.corner_r, cc = coins.corner_c;
  int ce_final = ce;
  if (cr != cr_pred || cc != cc_pred) {      // '@' issued its own order: restage
    __syncwarp();
    ce_final = (cc >> 5) & ~1;
    for (int i = lane; i < H * 4; i += 32)
      s_coin[i] = coin_pat[(int64_t)(cr + (i >> 2)) * PWW + ce_final + (i & 3)];
  }
  __syncwarp();
  g_sprites[lane] = rec[lane];
  if (lane < 16) g_drapes[lane] = rec[32 + lane];
  else g_plot[lane - 16] = rec[32 + lane];

  // ---- 5. final render, z-order a b c @ # P (engine.py:737-759) ----------
  // 5a. Window rows -> ONE word per 16-cell board segment (wall16 << 16 | coin16),
  // aligned to the board: each lane shifts whole rows once, so the streaming loop
  // below does no bit addressing at all.  Cells past W and the stale coin
  // (drapes.py:689 has not refreshed the curtain yet) are folded in here.
  const int pitch = p.pitch;
  const int spr = pitch >> 4;                // 16-byte segments per row
  uint32_t* s_seg = s_coin + H * 4;
  {
    const int wsh = wc - (we << 5), csh = cc - (ce_final << 5);   // 0..63 into the staged row
    const uint32_t m_lo = W >= 32 ? 0xffffffffu : (1u << W) - 1u;
    const uint32_t m_hi = W >= 64 ? 0xffffffffu : W > 32 ? (1u << (W - 32)) - 1u : 0u;
    for (int r = lane; r < H; r += 32) {
      const uint4 wv = *reinterpret_cast<const uint4*>(s_wall + r * 4);
      const uint4 cv = *reinterpret_cast<const uint4*>(s_coin + r * 4);
      const uint32_t wa = (wsh & 32) ? wv.y : wv.x, wb = (wsh & 32) ? wv.z : wv.y,
                     wd = (wsh & 32) ? wv.w : wv.z;
      const uint32_t ca = (csh & 32) ? cv.y : cv.x, cb = (csh & 32) ? cv.z : cv.y,
                     cd = (csh & 32) ? cv.w : cv.z;
      const uint32_t w_lo = __funnelshift_r(wa, wb, wsh & 31) & m_lo;
      const uint32_t w_hi = __funnelshift_r(wb, wd, wsh & 31) & m_hi;
      const uint32_t c_lo = __funnelshift_r(ca, cb, csh & 31) & m_lo;
      const uint32_t c_hi = __funnelshift_r(cb, cd, csh & 31) & m_hi;
      uint32_t* out = s_seg + r * spr;
      out[0] = __byte_perm(c_lo, w_lo, 0x5410);
      if (spr > 1) out[1] = __byte_perm(c_lo, w_lo, 0x7632);
      if (spr > 2) out[2] = __byte_perm(c_hi, w_hi, 0x5410);
      if (spr > 3) out[3] = __byte_perm(c_hi, w_hi, 0x7632);
    }
  }
  // a, b, c lie under both drapes, so they are patched into the staged backdrop
  // up front (in z-order, by one lane); the player is the top layer and is
  // patched into the one segment that holds it.
  __syncwarp();
  if (lane == 0) {
    if (coins.aux0 >= 0) s_seg[coins.aux0 * spr + (coins.aux1 >> 4)] |= 1u << (coins.aux1 & 15);
#pragma unroll
    for (int i = 1; i < kS; ++i)
      if (visible(sp[i])) s_bd[sp[i].row * pitch + sp[i].col] = p.sprite_char[i];
  }
  __syncwarp();
  // 5b. The streaming loop: 16 cells per lane per iteration, segment index ==
  // 16-byte index into both the staged tile and the board (pitch = 16 * spr).
  const int total = H * spr;
  const int p_seg = visible(sp[0]) ? sp[0].row * spr + (sp[0].col >> 4) : -1;
  const int p_word = (sp[0].col & 15) >> 2;
  const uint32_t p_keep = ~(0xffu << ((sp[0].col & 3) * 8));
  const uint32_t p_char = (uint32_t)p.sprite_char[0] << ((sp[0].col & 3) * 8);
  const unsigned drape_chars = ('#' << 8) | '@';           // bytes 4 and 5 of the permute
  const uint4* src = reinterpret_cast<const uint4*>(s_bd);
  uint4* dst = reinterpret_cast<uint4*>(p.out.d_board + (int64_t)env * H * pitch);
  for (int seg = lane; seg < total; seg += 32) {
    uint4 px = src[seg];
    const uint32_t bits = s_seg[seg];
    px.x = prmt(px.x, drape_chars, s_sel[((bits >> 12) & 0xf0u) | (bits & 0xfu)]);
    px.y = prmt(px.y, drape_chars, s_sel[((bits >> 16) & 0xf0u) | ((bits >> 4) & 0xfu)]);
    px.z = prmt(px.z, drape_chars, s_sel[((bits >> 20) & 0xf0u) | ((bits >> 8) & 0xfu)]);
    px.w = prmt(px.w, drape_chars, s_sel[((bits >> 24) & 0xf0u) | ((bits >> 12) & 0xfu)]);
    if (seg == p_seg) {
      if (p_word == 0) px.x = (px.x & p_keep) | p_char;
      else if (p_word == 1) px.y = (px.y & p_keep) | p_char;
      else if (p_word == 2) px.z = (px.z & p_keep) | p_char;
      else px.w = (px.w & p_keep) | p_char;
    }
    dst[seg] = px;
  }
}

}  // namespace

cudaError_t launch_scrolly_maze(const StepParams& p, cudaStream_t s) {
  if (p.W > 64 || (p.PWW & 1)) return cudaErrorInvalidValue;   // 4-word staged window rows
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const size_t smem = warp_smem_bytes(p.H, p.pitch) * kWarpsPerBlock;
  if (smem > 48 * 1024) {   // opt in per launch: the attribute is per device, handles are not
    cudaError_t e = cudaFuncSetAttribute(scrolly_maze_step,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // Programmatic dependent launch: this kernel may start (prologue only) before
  // the previous kernel of the stream has drained; it calls griddepcontrol.wait
  // before its first global access.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(kWarpsPerBlock * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, scrolly_maze_step, p);
}

}  // namespace pcl
