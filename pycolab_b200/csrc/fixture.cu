// fixture.cu — the general step program: any mix of MazeWalkers, Scrollys and
// plain drapes, any update schedule, per-env dynamic z-order.
//
// This is the device counterpart of the reference's own test fixtures
// (tests/test_things.py: TestMazeWalker :203-250, TestScrolly :253-295,
// TestDrape :178-200): every entity performs the motion its action slot names
// (walkers/Scrollys call the matching motion helper, `_stay` by default), and
// Plot directives (add_reward, terminate_episode, change_z_order — injected
// upstream with test_things.post_update) arrive as extra action words.  It
// exists so that the prefab semantics (sprites.py, drapes.py, scrolling.py) and
// the engine's staging/z-order rules (engine.py:698-847) are exercised on the
// GPU in their full generality — diagonal moves, EDGE, confined walkers,
// arbitrary impassable sets, several egocentric walkers, margin-less
// scrolling — not only in the shapes the three example games use.
//
// Unlike the game kernels it keeps a real board: the render after every update
// group (engine.py:735) is materialised in shared memory, because an arbitrary
// walker may test any character.  One warp per env; speed is not the point.
//
// Action row (i32 [n_entities + 2 * PCL_FIXTURE_DIRECTIVES]): motion code per entity
// in UPDATE order, then up to PCL_FIXTURE_DIRECTIVES (opcode, argument) pairs applied
// in order, i.e. in the order the entities issued them (include/pcl.h PCL_DIR_*):
// add_reward(int), terminate_episode(f32 discount), change_default_discount(f32),
// change_z_order(move_this | in_front_of << 8).
#include "pcl_device.cuh"
#include "pcl_kernels.cuh"

namespace pcl {

namespace {

constexpr int kWarpsPerBlock = 2;
constexpr int kMaxEnt = PCL_MAX_SPRITES + PCL_MAX_DRAPES;

struct WarpState {                   // lives in shared memory, one per warp
  int32_t sprites[PCL_MAX_SPRITES][PCL_SPRITE_WORDS];
  int32_t drapes[PCL_MAX_DRAPES][PCL_DRAPE_WORDS];
  int32_t plot[PCL_PLOT_WORDS];
  uint32_t impassable[PCL_MAX_SPRITES][4];
  // One record per scrolling group (protocols/scrolling.py:198-241): the group an
  // entity belongs to is swapped into the `Plot` registers around its update.
  int32_t groups[PCL_MAX_SCROLL_GROUPS][PCL_GROUP_WORDS];
  uint8_t z[kMaxEnt + 8];
};

// The order / egocentric-set registers of scrolling group g <-> `plot`.
__device__ __forceinline__ void group_in(Plot& plot, const WarpState* st, int g) {
  plot.order_r = st->groups[g][PCL_G_ORDER_R]; plot.order_c = st->groups[g][PCL_G_ORDER_C];
  plot.order_frame = st->groups[g][PCL_G_ORDER_FRAME];
  plot.ego_mask = st->groups[g][PCL_G_EGO_MASK];
}
__device__ __forceinline__ void group_out(const Plot& plot, WarpState* st, int g, int lane) {
  __syncwarp();
  if (lane == 0) {
    st->groups[g][PCL_G_ORDER_R] = plot.order_r; st->groups[g][PCL_G_ORDER_C] = plot.order_c;
    st->groups[g][PCL_G_ORDER_FRAME] = plot.order_frame;
    st->groups[g][PCL_G_EGO_MASK] = plot.ego_mask;
  }
  __syncwarp();
}

__device__ __forceinline__ bool in_set(const uint32_t* set, int code) {
  return (set[(code >> 5) & 3] >> (code & 31)) & 1u;
}

struct Ctx {
  const StepParams* p;
  WarpState* st;
  uint8_t* board;                    // smem, H * pitch
  const uint8_t* backdrop;
  int env, lane;
  int64_t lvl;                       // index of static level data
};

__device__ __forceinline__ bool drape_bit(const Ctx& c, int d, int r, int col) {
  const StepParams& p = *c.p;
  if (p.drape_kind[d]) {             // Scrolly: window of the pattern (drapes.py:689-695)
    const uint32_t* pat = p.st.d_pattern[d] + c.lvl * p.st.pattern_bstride[d];
    const int pr = c.st->drapes[d][PCL_D_CORNER_R] + r, pc = c.st->drapes[d][PCL_D_CORNER_C] + col;
    return bit_at(pat + (int64_t)pr * p.PWW, pc);
  }
  const uint32_t* bits = p.st.d_bits[d] + (int64_t)c.env * p.st.bits_bstride[d];
  return bit_at(bits + (int64_t)r * p.BW, col);
}

// engine.py:737-759 + rendering.py:98-160 into the smem board.
__device__ void render(const Ctx& c) {
  const StepParams& p = *c.p;
  const int n = p.S + p.D, cells = p.H * p.W;
  for (int i = c.lane; i < cells; i += 32) {
    const int r = i / p.W, col = i - r * p.W;
    int code = c.backdrop[(int64_t)r * p.pitch + col];
    for (int k = 0; k < n; ++k) {
      const int ch = c.st->z[k];
      for (int s = 0; s < p.S; ++s) {
        if (p.sprite_char[s] == ch) {
          const int32_t* rec = c.st->sprites[s];
          if ((rec[PCL_S_FLAGS] & 1) && rec[PCL_S_ROW] == r && rec[PCL_S_COL] == col) code = ch;
        }
      }
      for (int d = 0; d < p.D; ++d)
        if (p.drape_char[d] == ch && drape_bit(c, d, r, col)) code = ch;
    }
    c.board[(int64_t)r * p.pitch + col] = (uint8_t)code;
  }
  __syncwarp();
}

__device__ __forceinline__ Sprite load_sprite(const int32_t* r) {
  Sprite s;
  s.row = r[0]; s.col = r[1]; s.vrow = r[2]; s.vcol = r[3];
  s.flags = r[4]; s.aux0 = r[5]; s.aux1 = r[6]; s.aux2 = r[7];
  return s;
}
__device__ __forceinline__ void store_sprite(int32_t* r, const Sprite& s, int lane) {
  __syncwarp();
  if (lane == 0) {
    r[0] = s.row; r[1] = s.col; r[2] = s.vrow; r[3] = s.vcol;
    r[4] = s.flags; r[5] = s.aux0; r[6] = s.aux1; r[7] = s.aux2;
  }
  __syncwarp();
}

// scrolling.py:437-482 over the sprites in shared memory.
__device__ bool is_possible(const Ctx& c, const Plot& plot, int motion) {
  bool ok = true;
  for (int i = 0; i < c.p->S; ++i) {
    if ((plot.ego_mask >> i) & 1) {
      const int32_t* r = c.st->sprites[i];
      ok = ok && (r[PCL_S_AUX1] == plot.frame) && ((r[PCL_S_AUX0] >> motion) & 1);
    }
  }
  return ok;
}

// drapes.py:487-659 for drape d (same logic as pcl::scrolly_move, dynamic S).
__device__ void scrolly_move_dyn(const Ctx& c, int d, int motion, Plot& plot) {
  const StepParams& p = *c.p;
  int32_t* rec = c.st->drapes[d];
  int corner_r = rec[PCL_D_CORNER_R], corner_c = rec[PCL_D_CORNER_C];
  int pre_r = rec[PCL_D_PRE_R], pre_c = rec[PCL_D_PRE_C], last = rec[PCL_D_LAST_FRAME];
  const ScrollyCfg cfg = scrolly_cfg(p.H, p.W, p.PH, p.PW, p.margin[d][0], p.margin[d][1]);
  if (last < plot.frame) { last = plot.frame; pre_r = corner_r; pre_c = corner_c; }
  const int dr = motion_dr(motion), dc = motion_dc(motion);
  if (plot.order_frame == plot.frame) {
    if (dr != plot.order_r && dc != plot.order_c) plot.error |= PCL_ENV_ERR_ORDER_MISMATCH;
    corner_r += plot.order_r; corner_c += plot.order_c;
  } else if (motion != PCL_M_STAY) {
    if (!cfg.have_margins) {
      if (is_possible(c, plot, motion)) {
        const int nr = corner_r + dr, nc = corner_c + dc;
        const int orr = (0 <= nr && nr <= cfg.limit_r) ? dr : 0;
        const int occ = (0 <= nc && nc <= cfg.limit_c) ? dc : 0;
        corner_r += orr; corner_c += occ;
        plot.order_r = orr; plot.order_c = occ; plot.order_frame = plot.frame;
      }
    } else {
      bool want_v = false, want_h = false;
      for (int i = 0; i < p.S; ++i) {
        if ((plot.ego_mask >> i) & 1) {
          const int32_t* s = c.st->sprites[i];
          const int row = s[PCL_S_ROW], col = s[PCL_S_COL];
          const int nr = row + dr, nc = col + dc;
          want_v |= (row > nr && nr <= cfg.m_north) || (row < nr && nr >= cfg.m_south);
          want_h |= (col > nc && nc <= cfg.m_west) || (col < nc && nc >= cfg.m_east);
        }
      }
      if (want_v || want_h) {
        const int orr = want_v ? dr : 0, occ = want_h ? dc : 0;
        const int nr = corner_r + orr, nc = corner_c + occ;
        bool can = (0 <= nr && nr <= cfg.limit_r) && (0 <= nc && nc <= cfg.limit_c);
        can = can && is_possible(c, plot, motion);
        if (can) {
          corner_r = nr; corner_c = nc;
          plot.order_r = orr; plot.order_c = occ; plot.order_frame = plot.frame;
        }
      }
    }
  }
  __syncwarp();
  if (c.lane == 0) {
    rec[PCL_D_CORNER_R] = corner_r; rec[PCL_D_CORNER_C] = corner_c;
    rec[PCL_D_PRE_R] = pre_r; rec[PCL_D_PRE_C] = pre_c; rec[PCL_D_LAST_FRAME] = last;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
fixture_step(const StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int env = blockIdx.x * kWarpsPerBlock + warp;
  if (env >= p.B) return;
  const int64_t lvl = p.st.d_level ? p.st.d_level[env] : env;   // index of static level data
  const int H = p.H, W = p.W, S = p.S, D = p.D, n = S + D;
  const size_t board_bytes = ((size_t)H * p.pitch + 15) & ~(size_t)15;
  uint8_t* my = smem_raw + warp * (sizeof(WarpState) + board_bytes);
  WarpState* st = reinterpret_cast<WarpState*>(my);
  Ctx c;
  c.p = &p; c.st = st; c.board = my + sizeof(WarpState);
  c.backdrop = p.st.d_backdrop + lvl * p.st.backdrop_bstride;
  c.env = env; c.lane = lane; c.lvl = lvl;

  int32_t* g_sprites = p.st.d_sprites + (int64_t)env * S * PCL_SPRITE_WORDS;
  int32_t* g_drapes = p.st.d_drapes + (int64_t)env * D * PCL_DRAPE_WORDS;
  int32_t* g_plot = p.st.d_plot + (int64_t)env * PCL_PLOT_WORDS;
  uint8_t* g_z = p.st.d_z_order + (int64_t)env * n;
  uint8_t* g_board = p.out.d_board + (int64_t)env * H * p.pitch;

  const int was_over = g_plot[PCL_P_GAME_OVER];
  bool restart;
  if (p.mode == MODE_RESET) {
    restart = (p.env_mask == nullptr) || (p.env_mask[env] != 0);
    if (!restart) return;
  } else {
    restart = was_over && p.auto_reset;
    if (was_over && !p.auto_reset) return;
  }
  const int32_t* src_s = restart ? p.st.d_sprites_init + lvl * p.st.sprites_init_bstride
                                 : g_sprites;
  const int32_t* src_d = restart ? p.st.d_drapes_init + lvl * p.st.drapes_init_bstride
                                 : g_drapes;
  const int32_t* src_p = restart ? p.st.d_plot_init + lvl * p.st.plot_init_bstride
                                 : g_plot;
  const uint8_t* src_z = restart ? p.st.d_z_order_init + lvl * p.st.z_order_init_bstride
                                 : g_z;
  const int episodes = g_plot[PCL_P_EPISODES], old_error = g_plot[PCL_P_ERROR];
  for (int i = lane; i < S * PCL_SPRITE_WORDS; i += 32) (&st->sprites[0][0])[i] = src_s[i];
  for (int i = lane; i < D * PCL_DRAPE_WORDS; i += 32) (&st->drapes[0][0])[i] = src_d[i];
  if (lane < PCL_PLOT_WORDS) st->plot[lane] = src_p[lane];
  if (lane < n) st->z[lane] = src_z[lane];
  for (int i = lane; i < S * 4; i += 32) (&st->impassable[0][0])[i] = p.impassable[i >> 2][i & 3];
  for (int i = lane; i < (int)board_bytes; i += 32) c.board[i] = 0;
  __syncwarp();
  // Scrolling groups: group 0 is the plot record's own order / egocentric words,
  // the others come from (or are restored into) pcl_state.d_groups.
  const int n_sg = p.n_scroll_groups;
  int32_t* g_groups = p.st.d_groups
      ? p.st.d_groups + (int64_t)env * PCL_MAX_SCROLL_GROUPS * PCL_GROUP_WORDS : nullptr;
  if (lane < PCL_GROUP_WORDS) st->groups[0][lane] = st->plot[PCL_P_ORDER_R + lane];
  if (g_groups && n_sg > 1) {
    const int32_t* src_g = restart ? p.st.d_groups_init + lvl * p.st.groups_init_bstride : g_groups;
    if (lane >= PCL_GROUP_WORDS && lane < n_sg * PCL_GROUP_WORDS)
      (&st->groups[0][0])[lane] = src_g[lane];
  }
  __syncwarp();
  if (restart) {
    if (lane == 0) { st->plot[PCL_P_EPISODES] = episodes + 1; st->plot[PCL_P_ERROR] = old_error; }
    __syncwarp();
    render(c);                               // the pre-initial render, engine.py:572-578
  } else {
    // The board every entity of the first group reads = last step's final board.
    const int n16 = (H * p.pitch) >> 4;
    for (int i = lane; i < n16; i += 32)
      reinterpret_cast<uint4*>(c.board)[i] = reinterpret_cast<const uint4*>(g_board)[i];
    __syncwarp();
  }

  Plot plot;
  plot.frame = st->plot[PCL_P_FRAME] + 1;    // engine.py:716
  plot.error = st->plot[PCL_P_ERROR];
  group_in(plot, st, 0);
  Directives dir = fresh_directives();
  const int32_t* act = restart ? nullptr : p.actions + (int64_t)env * p.actions_per_env;

  // ---- update groups (engine.py:725-735)
  int k = 0;
  for (int g = 0; g < p.n_groups; ++g) {
    for (int e = 0; e < p.group_len[g]; ++e, ++k) {
      const int ch = p.group_chars[k];
      int motion = act ? act[k] : PCL_M_STAY;
      if (motion < 0 || motion > PCL_M_STAY) motion = PCL_M_STAY;   // test_things.py:247
      for (int s = 0; s < S; ++s) {
        if (p.sprite_char[s] != ch) continue;
        Sprite sp = load_sprite(st->sprites[s]);
        const uint32_t* imp = st->impassable[s];
        const uint8_t* board = c.board;
        const int pitch = p.pitch;
        group_in(plot, st, p.sprite_group[s]);
        walker_move(sp, s, motion, plot, H, W, p.confined[s] != 0, p.egocentric[s] != 0, lane,
                    [&](int r, int col) { return in_set(imp, board[r * pitch + col]); });
        group_out(plot, st, p.sprite_group[s], lane);
        store_sprite(st->sprites[s], sp, lane);
      }
      for (int d = 0; d < D; ++d) {
        if (p.drape_char[d] == ch && p.drape_kind[d]) {
          group_in(plot, st, p.drape_group[d]);
          scrolly_move_dyn(c, d, motion, plot);
          group_out(plot, st, p.drape_group[d], lane);
        }
      }
    }
    render(c);
  }

  // ---- Plot directives (plot.py:136-260) + _apply_and_clear_plot (engine.py:761-847)
  if (act) {
    bool z_changed = false;
    for (int i = 0; i < PCL_FIXTURE_DIRECTIVES; ++i) {
      const int op = act[n + 2 * i], arg = act[n + 2 * i + 1];
      if (op == PCL_DIR_ADD_REWARD) {
        add_reward(dir, arg);                                    // plot.py:201-214
      } else if (op == PCL_DIR_TERMINATE) {
        terminate(dir, __int_as_float(arg));                     // plot.py:176-199
      } else if (op == PCL_DIR_DEFAULT_DISCOUNT) {
        change_default_discount(dir, __int_as_float(arg));       // plot.py:247-260
      } else if (op == PCL_DIR_Z_ORDER) {                        // plot.py:136-174
        const int z_this = arg & 0xff, z_that = (arg >> 8) & 0xff;   // 0 = None (rearmost)
        bool have_this = false, have_that = (z_that == 0);
        for (int k2 = 0; k2 < n; ++k2) {
          have_this |= st->z[k2] == z_this;
          have_that |= st->z[k2] == z_that;
        }
        // Moving an entity in front of itself makes upstream DROP it from the
        // catalogue (engine.py:826-832 skips it and never re-inserts it); that is
        // reported as a bad directive here instead of corrupting the z-order.
        if (!have_this || !have_that || z_this == z_that) {
          plot.error |= PCL_ENV_ERR_BAD_Z;   // engine.py:802-812 raises RuntimeError
        } else {
          __syncwarp();
          if (lane == 0) {
            uint8_t fresh[kMaxEnt];
            int m = 0;
            if (z_that == 0) fresh[m++] = (uint8_t)z_this;
            for (int k2 = 0; k2 < n; ++k2) {
              const uint8_t chz = st->z[k2];
              if (chz == z_this) continue;
              fresh[m++] = chz;
              if (chz == z_that) fresh[m++] = (uint8_t)z_this;
            }
            for (int k2 = 0; k2 < n; ++k2) st->z[k2] = fresh[k2];
          }
          __syncwarp();
          z_changed = true;
        }
      }
    }
    if (z_changed) render(c);                // should_rerender, engine.py:636
  }

  __syncwarp();
  if (lane == 0) {
    st->plot[PCL_P_FRAME] = plot.frame; st->plot[PCL_P_GAME_OVER] = dir.game_over;
    st->plot[PCL_P_ERROR] = plot.error;
    for (int w = 0; w < PCL_GROUP_WORDS; ++w) st->plot[PCL_P_ORDER_R + w] = st->groups[0][w];
    p.out.d_reward[env] = dir.reward;
    p.out.d_has_reward[env] = (uint8_t)dir.has_reward;
    p.out.d_discount[env] = dir.discount;
    p.out.d_done[env] = (uint8_t)dir.game_over;
  }
  __syncwarp();
  for (int i = lane; i < S * PCL_SPRITE_WORDS; i += 32) g_sprites[i] = (&st->sprites[0][0])[i];
  for (int i = lane; i < D * PCL_DRAPE_WORDS; i += 32) g_drapes[i] = (&st->drapes[0][0])[i];
  if (lane < PCL_PLOT_WORDS) g_plot[lane] = st->plot[lane];
  if (g_groups && n_sg > 1 && lane >= PCL_GROUP_WORDS && lane < n_sg * PCL_GROUP_WORDS)
    g_groups[lane] = (&st->groups[0][0])[lane];
  if (lane < n) g_z[lane] = st->z[lane];
  const int n16 = (H * p.pitch) >> 4;
  for (int i = lane; i < n16; i += 32)
    reinterpret_cast<uint4*>(g_board)[i] = reinterpret_cast<const uint4*>(c.board)[i];
}

}  // namespace

cudaError_t launch_fixture(const StepParams& p, cudaStream_t s) {
  const int blocks = (p.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const size_t board_bytes = ((size_t)p.H * p.pitch + 15) & ~(size_t)15;
  const size_t smem = (sizeof(WarpState) + board_bytes) * kWarpsPerBlock;
  if (smem > 48 * 1024) {   // opt in per launch: the attribute is per device, handles are not
    cudaError_t e = cudaFuncSetAttribute(fixture_step,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  fixture_step<<<blocks, kWarpsPerBlock * 32, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pcl
