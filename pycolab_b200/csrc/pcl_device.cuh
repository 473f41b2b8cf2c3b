// pcl_device.cuh — device building blocks shared by the fused step kernels.
//
// Execution model: ONE WARP PER ENV.  Entity registers (sprite / drape / plot
// records) are loaded once, held REDUNDANTLY in every lane's registers, and
// all game logic is warp-uniform scalar code (no divergence); the lanes split
// up only for (a) board look-ups around a sprite (one lane per neighbour cell,
// combined with __ballot_sync) and (b) the paint loop (one lane per 16-byte
// row segment).  Everything here restates reference semantics; the citations
// are to /root/reference/pycolab.
#pragma once

#include <stdint.h>
#include <limits.h>

#include "../../include/pcl.h"

#define PCL_FULL 0xffffffffu
#define PCL_NEVER INT_MIN          // "-inf" frame (drapes.py:371) / None frame

namespace pcl {

// ---------------------------------------------------------------- records --

struct Sprite {                     // things.py:339-391 + sprites.py:153-205
  int row, col, vrow, vcol, flags, aux0, aux1, aux2;
};
struct Drape {                      // drapes.py:293-376
  int corner_r, corner_c, pre_r, pre_c, last_frame, aux0, aux1, aux2;
};
struct Plot {                       // plot.py:69-104 + scrolling.py:198-241
  int frame, game_over, error, episodes;
  int order_r, order_c, order_frame, ego_mask;
  int aux0, aux1, aux2, aux3;
  int crop_r, crop_c, crop_init, reserved;
};
// Engine directives accumulated during one step (plot.py:69-104).
struct Directives {
  int reward;
  int has_reward;
  int game_over;
  float discount;
};

__device__ __forceinline__ Directives fresh_directives() {
  Directives d;
  d.reward = 0; d.has_reward = 0; d.game_over = 0; d.discount = 1.0f;
  return d;
}
__device__ __forceinline__ void add_reward(Directives& d, int r) {  // plot.py:201
  d.reward += r; d.has_reward = 1;
}
__device__ __forceinline__ void terminate(Directives& d, float discount = 0.0f) {  // plot.py:176-199
  d.game_over = 1; d.discount = discount;
}
// plot.py:247-260.  Upstream rebuilds the directives (discount 1.0) after every
// step (plot.py:345-356, engine.py:845), so the "default" lasts for this step only.
__device__ __forceinline__ void change_default_discount(Directives& d, float discount) {
  d.discount = discount;
}

// ------------------------------------------------------------- MazeWalker --

__device__ __forceinline__ bool visible(const Sprite& s) { return s.flags & 1; }
__device__ __forceinline__ bool on_board(int r, int c, int H, int W) {
  return (unsigned)r < (unsigned)H && (unsigned)c < (unsigned)W;  // sprites.py:548
}

// sprites.py:315-352 (+ _on_board_exit/_on_board_enter :223-275).
__device__ __forceinline__ void walker_teleport(Sprite& s, int H, int W, int vr, int vc) {
  const bool was_on = on_board(s.vrow, s.vcol, H, W);
  const bool now_on = on_board(vr, vc, H, W);
  if (was_on && !now_on) {                     // exit: stash visibility, hide
    const int prior = (s.flags & 1) ? 2 : 1;
    s.flags = (prior << 1);                    // visible = 0
  }
  s.vrow = vr; s.vcol = vc;
  s.row = now_on ? vr : 0;
  s.col = now_on ? vc : 0;
  if (!was_on && now_on) {                     // enter: restore visibility
    const int prior = (s.flags >> 1) & 3;      // 0 None, 1 False, 2 True
    // `_visible = _prior_visible`; None is falsy for the renderer.
    s.flags = (s.flags & ~1) | (prior == 2 ? 1 : 0);
  }
}

// (drow, dcol) of a motion code, two bits per code packed in a constant:
// N NE E SE S SW W NW STAY -> drow+1 = 0 0 1 2 2 2 1 0 1, dcol+1 = 1 2 2 2 1 0 0 0 1.
__device__ __forceinline__ int motion_dr(int m) { return ((0x11A90 >> (2 * m)) & 3) - 1; }
__device__ __forceinline__ int motion_dc(int m) { return ((0x101A9 >> (2 * m)) & 3) - 1; }

// 3x3 neighbourhood "blocked" mask around the walker's VIRTUAL position: bit
// (dr+1)*3 + (dc+1).  Lane k < 9 evaluates one neighbour with `cell_blocked(r,
// c)` (the board of the last render vs the walker's impassable set,
// sprites.py:495-507); off-board neighbours are EDGE, which blocks only walkers
// confined to the board (sprites.py:503-506).
template <typename CellBlocked>
__device__ __forceinline__ unsigned neighbourhood(const Sprite& s, int H, int W,
                                                  bool confined, int lane,
                                                  CellBlocked cell_blocked) {
  const int k = lane < 9 ? lane : 4;
  const int r = s.vrow + k / 3 - 1;
  const int c = s.vcol + k % 3 - 1;
  bool blocked = false;
  if (lane < 9) {
    if (on_board(r, c, H, W)) blocked = cell_blocked(r, c);
    else blocked = confined;
  }
  return __ballot_sync(PCL_FULL, blocked) & 0x1ffu;
}

// sprites.py:479-546: is `motion` legal given the neighbourhood mask?
__device__ __forceinline__ bool motion_legal(unsigned blk, int motion) {
  const int dr = motion_dr(motion), dc = motion_dc(motion);
  if (dr == 0 && dc == 0) return true;
  const unsigned dest = (blk >> ((dr + 1) * 3 + (dc + 1))) & 1u;
  if (dr != 0 && dc != 0) {
    const unsigned f1 = (blk >> ((dr + 1) * 3 + 1)) & 1u;   // (dr, 0)
    const unsigned f2 = (blk >> (3 + (dc + 1))) & 1u;        // (0, dc)
    return !(dest || (f1 && f2));                            // sprites.py:539-543
  }
  return !dest;
}

// scrolling.py:437-482 over the single scrolling group '' the configured
// games use: every egocentric sprite must hold a permit for THIS frame that
// lists `motion`.  Permits live in the sprite record: aux0 = 9-bit motion mask,
// aux1 = frame the permit is valid at.
template <int S>
__device__ __forceinline__ bool scroll_is_possible(const Plot& plot, const Sprite (&sp)[S],
                                                   int motion) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    if ((plot.ego_mask >> i) & 1) {
      ok = ok && (sp[i].aux1 == plot.frame) && ((sp[i].aux0 >> motion) & 1);
    }
  }
  return ok;
}

// sprites.py:356-389 `_move` for sprite `idx`.
template <typename CellBlocked>
__device__ __forceinline__ bool walker_move(Sprite& s, int idx, int motion, Plot& plot,
                                            int H, int W, bool confined, bool egocentric,
                                            int lane, CellBlocked cell_blocked) {
  const int dr = motion_dr(motion), dc = motion_dc(motion);
  // _obey_scrolling_order, sprites.py:413-454
  if (egocentric) plot.ego_mask |= (1 << idx);
  if (plot.order_frame == plot.frame) {
    walker_teleport(s, H, W, s.vrow - plot.order_r, s.vcol - plot.order_c);
    if (egocentric && plot.order_r != dr && plot.order_c != dc)
      plot.error |= PCL_ENV_ERR_ORDER_MISMATCH;
  }
  bool legal = true;
  unsigned blk = 0;
  if (motion != PCL_M_STAY || egocentric) {
    blk = neighbourhood(s, H, W, confined, lane, cell_blocked);
    legal = motion_legal(blk, motion);
  }
  if (legal && motion != PCL_M_STAY) {
    walker_teleport(s, H, W, s.vrow + dr, s.vcol + dc);      // _raw_move :391
    if (egocentric) blk = neighbourhood(s, H, W, confined, lane, cell_blocked);
  }
  if (egocentric) {                                          // :456-477 + scrolling.py:373-434
    int mask = 1 << PCL_M_STAY;
#pragma unroll
    for (int m = 0; m < 8; ++m) mask |= (motion_legal(blk, m) ? 1 : 0) << m;
    const int valid_at = plot.frame + 1;
    if (s.aux1 != valid_at) { s.aux1 = valid_at; s.aux0 = 0; }
    s.aux0 |= mask;
  }
  return legal;
}

// ---------------------------------------------------------------- Scrolly --

struct ScrollyCfg {               // drapes.py:338-364
  int limit_r, limit_c;           // _northwest_corner_limit
  int have_margins;
  int m_north, m_south, m_west, m_east;
};

__device__ __forceinline__ ScrollyCfg scrolly_cfg(int H, int W, int PH, int PW,
                                                  int margin_r, int margin_c) {
  ScrollyCfg c;
  c.limit_r = PH - H; c.limit_c = PW - W;
  c.have_margins = margin_r >= 0;
  c.m_north = margin_r - 1; c.m_south = H - margin_r;
  c.m_west = margin_c - 1;  c.m_east = W - margin_c;
  return c;
}

// drapes.py:378-411 pattern_position_prescroll: refresh the pre-scroll corner
// when this Scrolly has not moved yet in this frame.
__device__ __forceinline__ void scrolly_touch_prescroll(Drape& d, const Plot& plot) {
  if (d.last_frame < plot.frame) { d.pre_r = d.corner_r; d.pre_c = d.corner_c; }
}

// drapes.py:487-659 `_maybe_move`.  Returns true when the curtain must be
// re-derived from the pattern (always, upstream: every path ends in
// _update_curtain); the window itself is never materialised here.
template <int S>
__device__ __forceinline__ void scrolly_move(Drape& d, const ScrollyCfg& cfg, int motion,
                                             Plot& plot, const Sprite (&sp)[S]) {
  if (d.last_frame < plot.frame) {
    d.last_frame = plot.frame;
    d.pre_r = d.corner_r; d.pre_c = d.corner_c;
  }
  const int dr = motion_dr(motion), dc = motion_dc(motion);
  if (plot.order_frame == plot.frame) {          // obey an existing order :513-535
    if (dr != plot.order_r && dc != plot.order_c) plot.error |= PCL_ENV_ERR_ORDER_MISMATCH;
    d.corner_r += plot.order_r; d.corner_c += plot.order_c;
    return;
  }
  if (motion == PCL_M_STAY) return;
  if (!cfg.have_margins) {                       // :598-623
    if (scroll_is_possible(plot, sp, motion)) {
      const int nr = d.corner_r + dr, nc = d.corner_c + dc;
      const int orr = (0 <= nr && nr <= cfg.limit_r) ? dr : 0;
      const int occ = (0 <= nc && nc <= cfg.limit_c) ? dc : 0;
      d.corner_r += orr; d.corner_c += occ;
      plot.order_r = orr; plot.order_c = occ; plot.order_frame = plot.frame;
    }
    return;
  }
  bool want_v = false, want_h = false;           // :625-642 + :661-687
#pragma unroll
  for (int i = 0; i < S; ++i) {
    if ((plot.ego_mask >> i) & 1) {
      const int nr = sp[i].row + dr, nc = sp[i].col + dc;   // TRUE position
      want_v |= (sp[i].row > nr && nr <= cfg.m_north) || (sp[i].row < nr && nr >= cfg.m_south);
      want_h |= (sp[i].col > nc && nc <= cfg.m_west) || (sp[i].col < nc && nc >= cfg.m_east);
    }
  }
  if (!(want_v || want_h)) return;
  const int orr = want_v ? dr : 0, occ = want_h ? dc : 0;
  const int nr = d.corner_r + orr, nc = d.corner_c + occ;
  bool can = (0 <= nr && nr <= cfg.limit_r) && (0 <= nc && nc <= cfg.limit_c);
  can = can && scroll_is_possible(plot, sp, motion);   // full motion, :650-651
  if (can) {
    d.corner_r = nr; d.corner_c = nc;
    plot.order_r = orr; plot.order_c = occ; plot.order_frame = plot.frame;
  }
}

// ------------------------------------------------------------ bit helpers --

// 16 consecutive bits starting at bit `off` of a bit-packed row (rows carry one
// zero pad word so the second load never leaves the row).
__device__ __forceinline__ unsigned bits16(const uint32_t* row, int off) {
  const int w = off >> 5, sh = off & 31;
  const uint32_t lo = row[w];
  const uint32_t hi = row[w + 1];
  return __funnelshift_r(lo, hi, sh) & 0xffffu;
}
__device__ __forceinline__ bool bit_at(const uint32_t* row, int c) {
  return (row[c >> 5] >> (c & 31)) & 1u;
}
// 4 mask bits -> 4 mask bytes (0x00 / 0xff), cell j in byte j.
__device__ __forceinline__ uint32_t spread4(unsigned nib) {
  return ((nib * 0x00204081u) & 0x01010101u) * 0xffu;
}
// Paint `ch` into the 16-byte segment `seg` wherever `bits` is set.
__device__ __forceinline__ void paint_bits(uint4& seg, unsigned bits, uint32_t ch) {
  const uint32_t c4 = ch * 0x01010101u;
  uint32_t m;
  m = spread4(bits & 15u);         seg.x = (seg.x & ~m) | (c4 & m);
  m = spread4((bits >> 4) & 15u);  seg.y = (seg.y & ~m) | (c4 & m);
  m = spread4((bits >> 8) & 15u);  seg.z = (seg.z & ~m) | (c4 & m);
  m = spread4((bits >> 12) & 15u); seg.w = (seg.w & ~m) | (c4 & m);
}
// One-cell sprite paint (rendering.py:139): bit mask for the sprite inside the
// segment that starts at (r, c0), or 0.
__device__ __forceinline__ unsigned sprite_bit(const Sprite& s, int r, int c0) {
  const int dc = s.col - c0;
  return (visible(s) && s.row == r && (unsigned)dc < 16u) ? (1u << dc) : 0u;
}

// ------------------------------------------------- TMA (1-D bulk) tile moves --
// A board tile is a contiguous, 16-byte aligned run of H * pitch bytes, so it
// moves between HBM and shared memory as ONE bulk-copy instruction issued by
// one lane (SASS UBLKCP), completion tracked by an mbarrier (loads) or a bulk
// group (stores), instead of H * pitch / 512 vector load/store rounds per warp.

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(arrivals));
  // Make the initialised barrier visible to the async proxy (the bulk copy
  // engine).  A CTA-scope proxy fence is enough for a barrier only this CTA
  // uses; the cluster-scope mbarrier_init fence would also flush L1 (CCTL.IVALL).
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Global -> shared; `bytes` a multiple of 16.  Call from ONE thread.
__device__ __forceinline__ void tile_load_bulk(void* smem_dst, const void* gmem_src,
                                               uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
               "[%0], [%1], %2, [%3];"
               :: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
// Shared -> global; make the generic-proxy writes to the tile visible to the
// async proxy first.  Call tile_store_bulk from ONE thread after a warp/block
// barrier; that thread must tile_store_wait() before the CTA can retire.
__device__ __forceinline__ void tile_store_fence() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tile_store_bulk(void* gmem_dst, const void* smem_src,
                                                uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tile_store_wait() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------- env scoping --

enum { MODE_STEP = 0, MODE_RESET = 1 };

}  // namespace pcl
