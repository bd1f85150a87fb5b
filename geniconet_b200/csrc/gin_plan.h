// Plan blobs shared between the host builders (gin_host.cpp) and the device kernels.
// A plan is a position-independent array of int32 words; offsets are in words from the
// start of the blob.  The caller uploads it once per (layer geometry) and passes the
// device pointer back to the op entry points (include/geniconet_b200.h).
#pragma once
#include <stdint.h>

#define GIN_MAGIC 0x47494E31  // "GIN1"
#define GIN_TILE_M 128        // rows (pixels) per gather-GEMM tile == UMMA M
#define GIN_MAX_SLOTS 24      // (bank, tap) slots a tile may carry

// Source codes inside the gather tables:
//   c >= 0  : pixel index inside the tile's sample group (sg*P_src + p)
//   c == -1 : contributes zero
//   c <= -2 : pole mean;  q = -2 - c,  sample-in-group = q >> 1,  pole = q & 1  (0 north, 1 south)
#define GIN_SRC_ZERO (-1)

struct GinTileDesc {          // 8 words
  int32_t nslots;             // number of (tap) slots this tile accumulates over
  int32_t src_off;            // word offset (relative to side.src_off) of int32 src[nslots][128]
  int8_t tap[GIN_MAX_SLOTS];  // weight (tap) index used by each slot
};

struct GinSide {              // one gather-GEMM problem: rows of dst gathered from src
  int32_t ntiles;             // tiles per sample group
  int32_t tiles_off;          // GinTileDesc[ntiles]
  int32_t src_off;            // base of all src tables
  int32_t rows_off;           // int32 dst_row[ntiles*128]: dst pixel inside group or -1
  int32_t P_src, P_dst;       // pixels per sample on the gathered / written side
  int32_t ring_off;           // int32 ring[2][5]: the pole rings at the src level
  int32_t max_slots;
};

// Patch ("P") tiles for stride-1 convolutions: a tile is R chart rows x Q octets (8 consecutive pixels of one
// chart row), R*Q = 16.  The kernel stages the (R+2) x (8Q+2) padded neighbourhood ONCE per 64-channel chunk as
// three column-shifted copies; the seven taps are then just different start addresses of the UMMA descriptor.
//   src[tile][((i' * Q + q) * 10 + c)]  = source code of padded cell (row i0-1+i', col j0_q-1+c) of octet-column q
//   rows[tile][(r * Q + q) * 8 + px]    = destination pixel (inside the sample group)
struct GinPSide {
  int32_t R, Q, U;            // U = (R+2)*Q*10 source rows per tile
  int32_t ntiles;             // per sample group (0 = patch mode not available for this plan)
  int32_t src_off, rows_off;
  int32_t ring_off, pad_;
};

// Patch tiles for STRIDE-2 convolutions.  Tiles live on the COARSE (output) lattice: R coarse rows x Q octets, R*Q = 16, and the
// cells of the coarse padded patch are numbered exactly as in GinPSide (cell (i', q, c) -> row (i'*Q + q)*10 + c).  The fine
// input is seen as four PARITY PLANES pl = 2*pr + pc: plane pl at coarse cell (I', J') is the fine padded cell (2I'+pr, 2J'+pc).
// A forward tap (di, dj) reads fine cell (2I+1+di, 2J+dj) = plane ((1+di)&1, dj&1) at coarse offset (a, b):
//     tap 0 (0,0): plane 2 (0,0)     tap 1 (-1,0): plane 0 (0,0)    tap 2 (1,0): plane 0 (+1,0)    tap 3 (0,-1): plane 3 (0,-1)
//     tap 4 (0,1): plane 3 (0,0)     tap 5 (-1,1): plane 1 (0,0)    tap 6 (1,-1): plane 1 (+1,-1)
// so per plane the conv is a 1-2 tap patch conv over one staged image (forward: the planes extend K; wgrad: one tap pair per
// plane).  dgrad runs the adjoint per plane: the fine pixels of plane pl gather the SAME coarse dy image with offsets (-a, -b).
struct GinP2Side {
  int32_t R, Q, U;            // U = (R+2)*Q*10
  int32_t ntiles;             // per sample group (0 = not available)
  int32_t src_off;            // int32 src[ntiles][4][U]: fine source code of (plane, cell); GIN_SRC_ZERO where no tap reads it
  int32_t dsrc_off;           // int32 dsrc[ntiles][U]: coarse dy pixel of the cell, GIN_SRC_ZERO outside the chart (in-chart part of dgrad)
  int32_t rows_off;           // int32 rows[ntiles][128]: coarse pixel of tile row (r*Q + q)*8 + px
  int32_t frows_off;          // int32 frows[ntiles][Q]: fine pixel (2*I0, 2*J0_q) of each octet column
};

// Cross-seam / pole remainder of dgrad in REGULAR form for the patch kernel: the boundary pixels of a sample group, 128 rows per
// tile, and one global list of slots = the taps that occur.  src[tile][slot][128] is the dy row the tile's r-th pixel receives
// through that tap (GIN_SRC_ZERO for most: a pixel has 1-5 entries).  Every tile runs all slots -- a few wasted MMAs buy a kernel
// without per-tile control flow; zero rows cost no memory traffic (cp.async zero-fill).  A pixel that uses one tap twice (the
// stitched corners) has one more row per further entry directly below its first row, inside the same 32-row group, with
// dst = -3: the epilogue adds such rows to the row above before the store (no atomics).
#define GIN_MAX_XSLOTS 16
struct GinPxSide {
  int32_t ntiles;             // per sample group (0 = not available)
  int32_t nslots;             // <= GIN_MAX_XSLOTS
  int32_t src_off;            // int32 src[ntiles][nslots][128]
  int32_t dst_off;            // int32 dst[ntiles][128]: pixel inside the sample group, -1 (no row), or -3 (extra row: add to the nearest pixel row above)
  int8_t tap[GIN_MAX_XSLOTS]; // weight index of each slot
};

// dgrad in ONE launch (r02): the patch kernel runs the in-chart tiles (GinPSide pdg / GinP2Side) and, appended to the same tile
// list, the BOUNDARY tiles below.  A boundary pixel (any pixel with a cross-seam or pole entry) gets its COMPLETE gradient here --
// in-chart and cross-seam entries alike, one slot per tap that occurs -- and the in-chart tiles skip its store (mask), so the two
// tile classes write disjoint pixels: no read-modify-write, no ordering between tiles, no second launch.  Layout of src / dst as
// in GinPxSide (extra rows with dst = -3 for a tap used more than once).
//   mask[tile][fl][4]: 128 bits per output tile of the in-chart pass (fl = 0 for stride 1; the four parity classes of a
//   stride-2 tile), bit r set = tile row r is a boundary pixel and is NOT stored by the in-chart pass.
struct GinPfSide {
  int32_t ntiles;             // boundary tiles per sample group (0 = not available: fall back to in-chart pass + GinPxSide pass)
  int32_t nslots;             // <= GIN_MAX_XSLOTS
  int32_t src_off;            // int32 src[ntiles][nslots][128]
  int32_t dst_off;            // int32 dst[ntiles][128]
  int32_t mask_off;           // uint32 mask[in-chart tiles per group][nfl][4]
  int32_t nfl;                // 1 (stride 1) or 4 (stride 2)
  int32_t all;                // 1: the boundary tiles cover EVERY pixel (small levels); the in-chart tiles are not run at all
  int32_t pad_;
  int8_t tap[GIN_MAX_XSLOTS];
};

struct GinConvPlanHdr {
  int32_t magic, kind, level_in, level_out, stride, corner_mode, group, total_words;
  GinSide fwd;                // y rows gathered from x   (also drives wgrad)
  GinSide dg;                 // dx rows gathered from dy (adjoint of pad o conv)
  GinPSide pfwd;              // stride 1 only: forward in patch mode (halo cells come through the chart stitching)
  GinPSide pdg;               // stride 1 only: in-chart part of dgrad in patch mode (halo cells are zero) ...
  GinSide dgx;                // ... plus the cross-seam / pole entries, ACCUMULATED on top by a gather-mode pass (both strides)
  GinP2Side p2;               // stride 2 only: forward / wgrad / in-chart dgrad in patch mode
  GinPxSide px;               // the same remainder as dgx, regular form (patch kernel, read-modify-write after the in-chart pass)
  GinPfSide pf;               // boundary pixels with ALL their entries + the in-chart store mask: dgrad in one launch
};

struct GinUpPlanHdr {
  int32_t magic, kind, level, corner_mode, Pc, Pf, total_words;
  int32_t fwd_off;            // int32 src[Pf][2]   (codes as above, sample-in-group = 0)
  int32_t ring_off;           // int32 ring[2][5] at the coarse level
  int32_t bwd_deg;            // ELL width of the adjoint
  int32_t bwd_idx_off;        // int32 idx[Pc][deg] fine pixel or -1
  int32_t bwd_w_off;          // float w[Pc][deg]
};

struct GinLossPlanHdr {
  int32_t magic, kind, level, P, V, total_words;
  int32_t ring_off;           // int32 nb[V][6] one-ring, counter-clockwise seen from outside, -1 padded
  int32_t pole_off;           // int32 ring[2][5] pixels averaged into the poles
  int32_t flag_off;           // int32 poleflag[P]: 1 = in north ring, 2 = in south ring
};
