// Host side of the C ABI: icosahedral chart geometry and the index tables ("plans")
// that drive every device kernel.  Everything here is table generation -- a different
// stitching convention (SURVEY.md 9.2 option B / other handedness) is a change to
// source_vertex() only, never to a kernel.
//
// Replaces (reference call sites; the implementation itself is the absent icocnn package):
//   chart padding inside IcoConvS2S / IcoUpsampleS2S      models.py:13-15,25-33,45-55
//   icocnn.utils.ico_geometry.get_ico_faces               losses.py:34
//   icocnn.utils.ico_geometry.get_icosahedral_grid        generate.py:151
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/geniconet_b200.h"
#include "gin_plan.h"

void gin_set_error(const char* fmt, ...);  // gin_api.cu

namespace {

// 7 live taps of the hex-masked 3x3 stencil; index == weight[..., t] index.
const int kTap[7][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, 1}, {1, -1}};

struct Geo {
  int s, n, W, P;
  explicit Geo(int level) : s(level), n(1 << level), W(2 << level), P(10 << (2 * level)) {}
  int vid(int k, int i, int j) const { return ((k % 5 + 5) % 5) * n * W + i * W + j; }
  // vertex feeding padded cell (i,j), i in [-1,n], j in [-1,2n] of chart k.
  // P = north pole, P+1 = south pole, -1 = never read.
  int source(int k, int i, int j) const {
    if (i >= 0 && i < n && j >= 0 && j < W) return vid(k, i, j);
    if (i == -1 && j == 0) return P;
    if (i == n - 1 && j == W) return P + 1;
    if ((i == -1 && j == -1) || (i == n && j == W)) return -1;
    if (i == -1) return (j <= n) ? vid(k - 1, j - 1, 0) : vid(k - 1, n - 1, j - n);
    if (j == W) return vid(k - 1, n - 1, n + i + 1);
    if (i == n) return (j <= n - 1) ? vid(k + 1, 0, j + n) : vid(k + 1, j - n, W - 1);
    /* j == -1 */ return vid(k + 1, 0, i);
  }
  void rings(int32_t out[10]) const {
    for (int k = 0; k < 5; ++k) {
      out[k] = vid(k, 0, 0);
      out[5 + k] = vid(k, n - 1, W - 1);
    }
  }
};

void vertices(const Geo& g, std::vector<double>& v) {
  const double zc = 1.0 / std::sqrt(5.0), rc = 2.0 / std::sqrt(5.0), PI = 3.14159265358979323846;
  double N[3] = {0, 0, 1}, S[3] = {0, 0, -1}, U[5][3], L[5][3];
  for (int k = 0; k < 5; ++k) {
    U[k][0] = rc * std::cos(2 * PI * k / 5); U[k][1] = rc * std::sin(2 * PI * k / 5); U[k][2] = zc;
    L[k][0] = rc * std::cos(2 * PI * (k - 0.5) / 5); L[k][1] = rc * std::sin(2 * PI * (k - 0.5) / 5); L[k][2] = -zc;
  }
  v.assign((size_t)(g.P + 2) * 3, 0.0);
  for (int k = 0; k < 5; ++k) {
    const double *Uk = U[k], *Um = U[(k + 4) % 5], *Lk = L[k], *Lm = L[(k + 4) % 5];
    for (int i = 0; i < g.n; ++i)
      for (int j = 0; j < g.W; ++j) {
        double a = double(i + 1) / g.n, b = double(j) / g.n, p[3];
        for (int d = 0; d < 3; ++d) {
          if (b <= 1.0) {
            p[d] = (a + b <= 1.0) ? N[d] + a * (Uk[d] - N[d]) + b * (Um[d] - N[d])
                                  : (1 - b) * Uk[d] + (a + b - 1) * Lk[d] + (1 - a) * Um[d];
          } else {
            double bb = b - 1.0;
            p[d] = (a + bb <= 1.0) ? Um[d] + a * (Lk[d] - Um[d]) + bb * (Lm[d] - Um[d])
                                   : (1 - bb) * Lk[d] + (a + bb - 1) * S[d] + (1 - a) * Lm[d];
          }
        }
        double nr = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        for (int d = 0; d < 3; ++d) v[(size_t)g.vid(k, i, j) * 3 + d] = p[d] / nr;
      }
  }
  for (int d = 0; d < 3; ++d) { v[(size_t)g.P * 3 + d] = N[d]; v[(size_t)(g.P + 1) * 3 + d] = S[d]; }
}

// one-ring of every vertex, counter-clockwise seen from outside; valence-5 rings are -1 padded
void one_rings(const Geo& g, std::vector<int32_t>& nb) {
  // hexagon walked in lattice order; handedness fixed below with the vertex positions
  static const int cyc[6] = {1, 5, 4, 2, 6, 3};  // (-1,0) (-1,+1) (0,+1) (+1,0) (+1,-1) (0,-1)
  std::vector<double> pos;
  vertices(g, pos);
  const int V = g.P + 2;
  nb.assign((size_t)V * 6, -1);
  auto orient = [&](int v, std::vector<int>& r) {
    double acc[3] = {0, 0, 0};
    for (size_t a = 0; a < r.size(); ++a) {
      const double* x = &pos[(size_t)r[a] * 3];
      const double* y = &pos[(size_t)r[(a + 1) % r.size()] * 3];
      acc[0] += x[1] * y[2] - x[2] * y[1]; acc[1] += x[2] * y[0] - x[0] * y[2]; acc[2] += x[0] * y[1] - x[1] * y[0];
    }
    const double* c = &pos[(size_t)v * 3];
    if (acc[0] * c[0] + acc[1] * c[1] + acc[2] * c[2] < 0) std::reverse(r.begin(), r.end());
    for (size_t a = 0; a < r.size(); ++a) nb[(size_t)v * 6 + a] = r[a];
  };
  for (int k = 0; k < 5; ++k)
    for (int i = 0; i < g.n; ++i)
      for (int j = 0; j < g.W; ++j) {
        std::vector<int> r;
        for (int c = 0; c < 6; ++c) {
          int u = g.source(k, i + kTap[cyc[c]][0], j + kTap[cyc[c]][1]);
          if (r.empty() || (r.back() != u)) r.push_back(u);
        }
        if (r.size() > 1 && r.front() == r.back()) r.pop_back();
        orient(g.vid(k, i, j), r);
      }
  int32_t rg[10];
  g.rings(rg);
  for (int pole = 0; pole < 2; ++pole) {
    std::vector<int> r(rg + 5 * pole, rg + 5 * pole + 5);
    orient(g.P + pole, r);
  }
}

void faces(const Geo& g, std::vector<int32_t>& f) {
  std::vector<int32_t> nb;
  one_rings(g, nb);
  const int V = g.P + 2;
  f.clear();
  for (int v = 0; v < V; ++v) {
    int deg = 0;
    while (deg < 6 && nb[(size_t)v * 6 + deg] >= 0) ++deg;
    for (int a = 0; a < deg; ++a) {
      int u = nb[(size_t)v * 6 + a], w = nb[(size_t)v * 6 + (a + 1) % deg];
      if (v < u && v < w) { f.push_back(v); f.push_back(u); f.push_back(w); }  // emit once, from its smallest vertex
    }
  }
}

// ---------------------------------------------------------------- conv plan -----------
struct Entry { int src; int tap; };  // src: pixel of the gathered tensor, or -2/-3 pole mean

struct SideBuild {
  int P_src = 0, P_dst = 0, level_src = 0;
  std::vector<GinTileDesc> tiles;
  std::vector<int32_t> src;   // concatenated [nslots][128] blocks
  std::vector<int32_t> rows;  // [ntiles*128]
  int max_slots = 0;
};

int group_size(int P) {
  int g = 1;
  while ((g * P) % GIN_TILE_M) ++g;
  return g;
}

// rows: for each dst pixel of ONE sample, its list of (src, tap) entries.
bool build_side(const std::vector<std::vector<Entry>>& rows_in, int P_src, int P_dst, int level_src, int group,
                bool sort_rows, SideBuild& out) {
  out.P_src = P_src; out.P_dst = P_dst; out.level_src = level_src;
  // slot id = bank*7 + tap, bank = how many earlier entries of this row used the same tap
  struct Row { uint32_t sig; int sg; int p; std::vector<std::pair<int, int>> slots; };  // (slot id, src)
  std::vector<Row> rows;
  rows.reserve((size_t)group * P_dst);
  for (int sg = 0; sg < group; ++sg)
    for (int p = 0; p < P_dst; ++p) {
      Row r; r.sig = 0; r.sg = sg; r.p = p;
      int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
      for (const Entry& e : rows_in[p]) {
        int slot = cnt[e.tap]++ * 7 + e.tap;
        if (slot >= 32) { gin_set_error("plan: tap multiplicity too high"); return false; }
        int code = (e.src >= 0) ? sg * P_src + e.src : -2 - (2 * sg + (-2 - e.src));
        r.slots.push_back({slot, code});
        r.sig |= 1u << slot;
      }
      rows.push_back(std::move(r));
    }
  if (sort_rows)
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.sig < b.sig; });
  const int ntiles = (int)(rows.size() + GIN_TILE_M - 1) / GIN_TILE_M;
  out.rows.assign((size_t)ntiles * GIN_TILE_M, -1);
  for (int t = 0; t < ntiles; ++t) {
    uint32_t uni = 0;
    for (int r = 0; r < GIN_TILE_M; ++r) {
      size_t idx = (size_t)t * GIN_TILE_M + r;
      if (idx < rows.size()) uni |= rows[idx].sig;
    }
    GinTileDesc d;
    std::memset(&d, 0, sizeof(d));
    int slot_of[32];
    for (int b = 0; b < 32; ++b) {
      slot_of[b] = -1;
      if (uni >> b & 1) {
        if (d.nslots >= GIN_MAX_SLOTS) { gin_set_error("plan: too many slots in a tile"); return false; }
        slot_of[b] = d.nslots;
        d.tap[d.nslots++] = (int8_t)(b % 7);
      }
    }
    d.src_off = (int32_t)out.src.size();
    out.src.resize(out.src.size() + (size_t)d.nslots * GIN_TILE_M, GIN_SRC_ZERO);
    for (int r = 0; r < GIN_TILE_M; ++r) {
      size_t idx = (size_t)t * GIN_TILE_M + r;
      if (idx >= rows.size()) continue;
      out.rows[idx] = rows[idx].sg * P_dst + rows[idx].p;
      for (auto& sl : rows[idx].slots) out.src[(size_t)d.src_off + (size_t)slot_of[sl.first] * GIN_TILE_M + r] = sl.second;
    }
    out.max_slots = std::max(out.max_slots, (int)d.nslots);
    out.tiles.push_back(d);
  }
  return true;
}

bool build_conv(int level, int stride, int corner_mode, std::vector<int32_t>& blob) {
  if (level < 0 || level > 9 || (stride != 1 && stride != 2) || (stride == 2 && level < 1) ||
      (corner_mode != 0 && corner_mode != 1)) {
    gin_set_error("hexconv plan: bad level/stride/corner_mode (%d,%d,%d)", level, stride, corner_mode);
    return false;
  }
  Geo gi(level), go(stride == 2 ? level - 1 : level);
  // forward gather table: out pixel -> 7 sources
  std::vector<std::vector<Entry>> fwd(go.P), adj(gi.P), adjx(gi.P);   // adjx: entries that cross a chart seam / a pole
  std::vector<std::pair<int, int>> pole_readers[2];  // (out pixel, tap)
  std::vector<char> is_boundary;                      // input-level pixels that own a cross-seam / pole entry of dgrad (filled with the seam plans)
  for (int k = 0; k < 5; ++k)
    for (int I = 0; I < go.n; ++I)
      for (int J = 0; J < go.W; ++J) {
        int po = go.vid(k, I, J);
        int ci = (stride == 1) ? I : 2 * I + 1, cj = (stride == 1) ? J : 2 * J;
        for (int t = 0; t < 7; ++t) {
          int v = gi.source(k, ci + kTap[t][0], cj + kTap[t][1]);
          if (v < 0) { gin_set_error("hexconv plan: live tap reads a dead cell"); return false; }
          if (v >= gi.P) {
            int pole = v - gi.P;
            if (corner_mode == GIN_CORNER_AVERAGE) {
              fwd[po].push_back({-2 - pole, t});
              pole_readers[pole].push_back({po, t});
            }  // 'zeros': contributes nothing
          } else {
            fwd[po].push_back({v, t});
            adj[v].push_back({po, t});
            const int ti = ci + kTap[t][0], tj = cj + kTap[t][1];
            if (!(ti >= 0 && ti < gi.n && tj >= 0 && tj < gi.W)) adjx[v].push_back({po, t});
          }
        }
      }
  // adjoint of the pole average: dx[ring_j] += (1/5) W_t^T sum_k dy[reader_k].  When the readers are
  // exactly dy's own pole ring with one tap (true for stride 1) this is W_t^T * mean over dy's ring.
  int32_t ring_in[10], ring_out[10];
  gi.rings(ring_in);
  go.rings(ring_out);
  for (int pole = 0; pole < 2; ++pole) {
    auto& rd = pole_readers[pole];
    if (rd.empty()) continue;
    bool ok = rd.size() == 5 && stride == 1;
    for (size_t a = 0; ok && a < rd.size(); ++a) {
      ok = rd[a].second == rd[0].second &&
           std::find(ring_out + 5 * pole, ring_out + 5 * pole + 5, rd[a].first) != ring_out + 5 * pole + 5;
    }
    if (!ok) { gin_set_error("hexconv plan: unexpected pole reader set"); return false; }
    for (int j = 0; j < 5; ++j) {
      adj[ring_in[5 * pole + j]].push_back({-2 - pole, rd[0].second});
      adjx[ring_in[5 * pole + j]].push_back({-2 - pole, rd[0].second});
    }
  }
  const int group = std::max(group_size(gi.P), group_size(go.P));
  SideBuild F, D;
  if (!build_side(fwd, gi.P, go.P, level, group, false, F)) return false;
  if (!build_side(adj, go.P, gi.P, go.s, group, true, D)) return false;

  GinConvPlanHdr h;
  std::memset(&h, 0, sizeof(h));
  const int hdr_words = (int)(sizeof(h) / 4);
  blob.assign(hdr_words, 0);
  auto emit_side = [&](SideBuild& sb, const int32_t* ring, GinSide& s) {
    s.ntiles = (int)sb.tiles.size();
    s.P_src = sb.P_src; s.P_dst = sb.P_dst; s.max_slots = sb.max_slots;
    s.tiles_off = (int)blob.size();
    blob.resize(blob.size() + sb.tiles.size() * (sizeof(GinTileDesc) / 4));
    std::memcpy(&blob[s.tiles_off], sb.tiles.data(), sb.tiles.size() * sizeof(GinTileDesc));
    s.src_off = (int)blob.size();
    blob.insert(blob.end(), sb.src.begin(), sb.src.end());
    s.rows_off = (int)blob.size();
    blob.insert(blob.end(), sb.rows.begin(), sb.rows.end());
    s.ring_off = (int)blob.size();
    blob.insert(blob.end(), ring, ring + 10);
  };
  emit_side(F, ring_in, h.fwd);
  emit_side(D, ring_out, h.dg);
  if (stride == 1 && level >= 2) {
    // ---- patch tiles
    const int R = (level >= 3) ? 8 : 4, Q = 16 / R, octs = gi.W / 8, rblocks = gi.n / R;
    struct OctCol { int sg, k, i0, j0; };
    std::vector<OctCol> cols;
    for (int sg = 0; sg < group; ++sg)
      for (int k = 0; k < 5; ++k)
        for (int bi = 0; bi < rblocks; ++bi)
          for (int jo = 0; jo < octs; ++jo) cols.push_back({sg, k, bi * R, jo * 8});
    if (cols.size() % Q) { gin_set_error("hexconv plan: octet columns do not tile"); return false; }
    const int ntiles = (int)cols.size() / Q, U = (R + 2) * Q * 10;
    std::vector<int32_t> psrc((size_t)ntiles * U), pdsrc((size_t)ntiles * U), prows((size_t)ntiles * GIN_TILE_M);
    for (int t = 0; t < ntiles; ++t)
      for (int q = 0; q < Q; ++q) {
        const OctCol& oc = cols[(size_t)t * Q + q];
        for (int ip = 0; ip < R + 2; ++ip)
          for (int c = 0; c < 10; ++c) {
            const int i = oc.i0 - 1 + ip, j = oc.j0 - 1 + c;
            const int v = gi.source(oc.k, i, j);
            int code = GIN_SRC_ZERO;
            if (v >= 0 && v < gi.P) code = oc.sg * gi.P + v;
            else if (v >= gi.P && corner_mode == GIN_CORNER_AVERAGE) code = -2 - (2 * oc.sg + (v - gi.P));
            const bool inside = i >= 0 && i < gi.n && j >= 0 && j < gi.W;
            const size_t e = (size_t)t * U + ((size_t)ip * Q + q) * 10 + c;
            psrc[e] = code;
            pdsrc[e] = inside ? code : GIN_SRC_ZERO;
          }
        for (int r = 0; r < R; ++r)
          for (int px = 0; px < 8; ++px)
            prows[(size_t)t * GIN_TILE_M + ((size_t)r * Q + q) * 8 + px] = oc.sg * gi.P + gi.vid(oc.k, oc.i0 + r, oc.j0 + px);
      }
    auto emit_p = [&](const std::vector<int32_t>& src, const int32_t* ring, GinPSide& ps) {
      ps.R = R; ps.Q = Q; ps.U = U; ps.ntiles = ntiles;
      ps.src_off = (int)blob.size();
      blob.insert(blob.end(), src.begin(), src.end());
      ps.rows_off = (int)blob.size();
      blob.insert(blob.end(), prows.begin(), prows.end());
      ps.ring_off = (int)blob.size();
      blob.insert(blob.end(), ring, ring + 10);
    };
    emit_p(psrc, ring_in, h.pfwd);
    emit_p(pdsrc, ring_out, h.pdg);
  }
  if (level >= 2 && (stride == 1 || go.s >= 2)) {
    // ---- the cross-seam remainder of dgrad: one row per BOUNDARY PIXEL carrying all of its cross-seam / pole entries as
    // slots (rows sorted by slot signature so a tile needs few slots).  A pixel appears in exactly one row, so the pass adds
    // into dx with a plain read-modify-write after the in-chart pass -- no atomics.
    SideBuild X;
    std::vector<std::vector<Entry>> only;
    std::vector<int> pix;
    for (int v = 0; v < gi.P; ++v)
      if (!adjx[v].empty()) { only.push_back(adjx[v]); pix.push_back(v); }
    if (!build_side(only, go.P, (int)only.size(), go.s, group, true, X)) return false;
    // build_side numbered the rows 0..only.size()-1 per sample; translate back to pixels of the full map
    for (auto& r : X.rows)
      if (r >= 0) { const int sg = r / (int)only.size(), idx = r % (int)only.size(); r = sg * gi.P + pix[idx]; }
    X.P_dst = gi.P;
    emit_side(X, ring_out, h.dgx);
    // ---- the same remainder in regular form (GinPxSide).  Slots = the TAPS that occur (bank 0).  The few pixels that use a tap
    // twice (the stitched corners) get one extra ROW per further entry, placed directly below the pixel's first row and inside
    // the same 32-row group (dst = -3): the epilogue warp that owns the group adds them to the row above before its single
    // read-modify-write, so the pass needs no atomics and its result does not depend on any execution order.
    auto regular_form = [&](const std::vector<std::vector<Entry>>& only, int& out_ntiles, int& out_nslots, int8_t* out_tap,
                            std::vector<int32_t>& xsrc, std::vector<int32_t>& xdst) -> bool {
      int slot_of_tap[7], nslots = 0;
      uint32_t taps_used = 0;
      for (const auto& row : only)
        for (const Entry& e : row) taps_used |= 1u << e.tap;
      for (int t = 0; t < 7; ++t) slot_of_tap[t] = (taps_used >> t & 1) ? nslots++ : -1;
      struct XRow { int pix; int span; std::vector<Entry> ent; };      // span: rows of this pixel (on its first row), 0 on extra rows
      std::vector<XRow> xr;
      bool ok = !only.empty();
      for (size_t i = 0; i < only.size() && ok; ++i) {
        int cnt[7] = {0, 0, 0, 0, 0, 0, 0}, maxbank = 0;
        for (const Entry& e : only[i]) maxbank = std::max(maxbank, cnt[e.tap]++);
        if (maxbank > 2) ok = false;                                   // the epilogue folds at most two extra rows
        for (int bank = 0; bank <= maxbank; ++bank) {
          XRow r{pix[i], bank == 0 ? maxbank + 1 : 0, {}};
          int c2[7] = {0, 0, 0, 0, 0, 0, 0};
          for (const Entry& e : only[i])
            if (c2[e.tap]++ == bank) r.ent.push_back(e);
          xr.push_back(std::move(r));
        }
      }
      if (ok) {
        std::vector<std::pair<int, int>> seq;                          // (sample in group, index into xr) or (-1, -1): padding row
        for (int sg = 0; sg < group; ++sg)
          for (size_t i = 0; i < xr.size(); ++i) {
            if (xr[i].span > 0)
              while ((int)(seq.size() % 32) + xr[i].span > 32) seq.emplace_back(-1, -1);
            seq.emplace_back(sg, (int)i);
          }
        const int rows_total = (int)seq.size(), ntiles = (rows_total + GIN_TILE_M - 1) / GIN_TILE_M;
        xsrc.assign((size_t)ntiles * nslots * GIN_TILE_M, GIN_SRC_ZERO);
        xdst.assign((size_t)ntiles * GIN_TILE_M, -1);
        for (int r = 0; r < rows_total; ++r) {
          if (seq[r].first < 0) continue;
          const int sg = seq[r].first, t = r / GIN_TILE_M, rr = r % GIN_TILE_M;
          const XRow& row = xr[seq[r].second];
          xdst[(size_t)t * GIN_TILE_M + rr] = row.span > 0 ? sg * gi.P + row.pix : -3;
          for (const Entry& e : row.ent) {
            const int code = (e.src >= 0) ? sg * go.P + e.src : -2 - (2 * sg + (-2 - e.src));
            xsrc[((size_t)t * nslots + slot_of_tap[e.tap]) * GIN_TILE_M + rr] = code;
          }
        }
        out_ntiles = ntiles; out_nslots = nslots;
        for (int t = 0; t < 7; ++t)
          if (slot_of_tap[t] >= 0) out_tap[slot_of_tap[t]] = (int8_t)t;
      }
      return ok;
    };
    {
      std::vector<int32_t> xsrc, xdst;
      int nt = 0, ns = 0;
      if (regular_form(only, nt, ns, h.px.tap, xsrc, xdst)) {
        h.px.ntiles = nt; h.px.nslots = ns;
        h.px.src_off = (int)blob.size();
        blob.insert(blob.end(), xsrc.begin(), xsrc.end());
        h.px.dst_off = (int)blob.size();
        blob.insert(blob.end(), xdst.begin(), xdst.end());
      }
    }
    // ---- one-launch dgrad (GinPfSide): the same boundary pixels carrying ALL their entries, and the store mask of the in-chart
    // tiles (filled in below, once the in-chart tiles of this stride exist)
    // At levels <= 3 (stride 1) more than a third of all pixels are boundary pixels: there the boundary form takes EVERY pixel and
    // the in-chart tiles are dropped altogether (pf.all = 1) -- 5 tile-units per 640 pixels instead of 5 + 3.
    const bool take_all = stride == 1 && level <= 3;
    if (take_all) {
      pix.clear();
      for (int v = 0; v < gi.P; ++v) pix.push_back(v);
    }
    {
      std::vector<std::vector<Entry>> full;
      for (int v : pix) full.push_back(adj[v]);
      std::vector<int32_t> xsrc, xdst;
      int nt = 0, ns = 0;
      if (regular_form(full, nt, ns, h.pf.tap, xsrc, xdst)) {
        h.pf.ntiles = nt; h.pf.nslots = ns;
        h.pf.src_off = (int)blob.size();
        blob.insert(blob.end(), xsrc.begin(), xsrc.end());
        h.pf.dst_off = (int)blob.size();
        blob.insert(blob.end(), xdst.begin(), xdst.end());
        h.pf.all = take_all ? 1 : 0;
      }
    }
    is_boundary.assign((size_t)gi.P, 0);
    for (int v : pix) is_boundary[v] = 1;
  }
  if (stride == 2 && go.s >= 2) {
    // ---- stride-2 patch tiles on the coarse lattice (see GinP2Side)
    const int R = (go.s >= 3) ? 8 : 4, Q = 16 / R, octs = go.W / 8, rblocks = go.n / R;
    struct OctCol { int sg, k, i0, j0; };
    std::vector<OctCol> cols;
    for (int sg = 0; sg < group; ++sg)
      for (int k = 0; k < 5; ++k)
        for (int bi = 0; bi < rblocks; ++bi)
          for (int jo = 0; jo < octs; ++jo) cols.push_back({sg, k, bi * R, jo * 8});
    if (cols.size() % Q) { gin_set_error("hexconv plan: stride-2 octet columns do not tile"); return false; }
    const int ntiles = (int)cols.size() / Q, U = (R + 2) * Q * 10;
    std::vector<int32_t> src((size_t)ntiles * 4 * U, GIN_SRC_ZERO), dsrc((size_t)ntiles * U, GIN_SRC_ZERO);
    std::vector<int32_t> rows((size_t)ntiles * GIN_TILE_M), frows((size_t)ntiles * Q);
    for (int t = 0; t < ntiles; ++t)
      for (int q = 0; q < Q; ++q) {
        const OctCol& oc = cols[(size_t)t * Q + q];
        for (int ip = 0; ip < R + 2; ++ip)
          for (int c = 0; c < 10; ++c) {
            const int I = oc.i0 - 1 + ip, J = oc.j0 - 1 + c;
            const size_t cell = ((size_t)ip * Q + q) * 10 + c;
            for (int pl = 0; pl < 4; ++pl) {
              const int pr = pl >> 1, pc = pl & 1;
              // cells some forward tap of this plane reads: rows I0 .. I0+R-1 (+1 more for the even-row planes), columns J0 .. J0+7
              // (-1 more for the odd-column planes)
              const bool used = ip >= 1 && ip <= (pr == 0 ? R + 1 : R) && c >= (pc == 1 ? 0 : 1) && c <= 8;
              if (!used) continue;
              const int fi = 2 * I + pr, fj = 2 * J + pc;
              if (fi < -1 || fi > gi.n || fj < -1 || fj > gi.W) continue;
              const int v = gi.source(oc.k, fi, fj);
              int code = GIN_SRC_ZERO;
              if (v >= 0 && v < gi.P) code = oc.sg * gi.P + v;
              else if (v >= gi.P && corner_mode == GIN_CORNER_AVERAGE) code = -2 - (2 * oc.sg + (v - gi.P));
              src[((size_t)t * 4 + pl) * U + cell] = code;
            }
            if (I >= 0 && I < go.n && J >= 0 && J < go.W) dsrc[(size_t)t * U + cell] = oc.sg * go.P + go.vid(oc.k, I, J);
          }
        for (int r = 0; r < R; ++r)
          for (int px = 0; px < 8; ++px)
            rows[(size_t)t * GIN_TILE_M + ((size_t)r * Q + q) * 8 + px] = oc.sg * go.P + go.vid(oc.k, oc.i0 + r, oc.j0 + px);
        frows[(size_t)t * Q + q] = oc.sg * gi.P + gi.vid(oc.k, 2 * oc.i0, 2 * oc.j0);
      }
    GinP2Side& p2 = h.p2;
    p2.R = R; p2.Q = Q; p2.U = U; p2.ntiles = ntiles;
    p2.src_off = (int)blob.size();
    blob.insert(blob.end(), src.begin(), src.end());
    p2.dsrc_off = (int)blob.size();
    blob.insert(blob.end(), dsrc.begin(), dsrc.end());
    p2.rows_off = (int)blob.size();
    blob.insert(blob.end(), rows.begin(), rows.end());
    p2.frows_off = (int)blob.size();
    blob.insert(blob.end(), frows.begin(), frows.end());
  }
  if (h.pf.ntiles > 0) {
    // ---- store mask of the in-chart dgrad tiles: the boundary pixels belong to the GinPfSide tiles
    std::vector<uint32_t> mask;
    if (stride == 1 && h.pdg.ntiles > 0) {
      h.pf.nfl = 1;
      mask.assign((size_t)h.pdg.ntiles * 4, 0u);
      for (int t = 0; t < h.pdg.ntiles; ++t)
        for (int r = 0; r < GIN_TILE_M; ++r) {
          const int d = blob[(size_t)h.pdg.rows_off + (size_t)t * GIN_TILE_M + r];
          if (d >= 0 && is_boundary[d % gi.P]) mask[(size_t)t * 4 + r / 32] |= 1u << (r % 32);
        }
    } else if (stride == 2 && h.p2.ntiles > 0) {
      h.pf.nfl = 4;
      const int Q = h.p2.Q;
      mask.assign((size_t)h.p2.ntiles * 16, 0u);
      for (int t = 0; t < h.p2.ntiles; ++t)
        for (int pl = 0; pl < 4; ++pl)
          for (int r = 0; r < GIN_TILE_M; ++r) {
            const int g = r >> 3, r_in = g / Q, q = g % Q, px = r & 7;
            const int fine = blob[(size_t)h.p2.frows_off + (size_t)t * Q + q] + r_in * 2 * gi.W + px * 2 + (pl >> 1) * gi.W + (pl & 1);
            if (is_boundary[fine % gi.P]) mask[((size_t)t * 4 + pl) * 4 + r / 32] |= 1u << (r % 32);
          }
    }
    if (mask.empty()) h.pf.ntiles = 0;            // no in-chart patch tiles at this size: the two-pass path stays
    else {
      h.pf.mask_off = (int)blob.size();
      for (uint32_t w : mask) blob.push_back((int32_t)w);
    }
  }
  h.magic = GIN_MAGIC; h.kind = GIN_PLAN_HEXCONV; h.level_in = level; h.level_out = go.s; h.stride = stride;
  h.corner_mode = corner_mode; h.group = group; h.total_words = (int)blob.size();
  std::memcpy(blob.data(), &h, sizeof(h));
  return true;
}

// ---------------------------------------------------------------- upsample plan -------
bool build_upsample(int level, int corner_mode, std::vector<int32_t>& blob) {
  if (level < 0 || level > 8 || (corner_mode != 0 && corner_mode != 1)) {
    gin_set_error("upsample plan: bad level/corner_mode (%d,%d)", level, corner_mode);
    return false;
  }
  Geo gc(level), gf(level + 1);
  std::vector<int32_t> fwd((size_t)gf.P * 2);
  std::vector<std::vector<std::pair<int, float>>> adj(gc.P);
  std::vector<int> pole_users[2];
  for (int k = 0; k < 5; ++k)
    for (int i = 0; i < gf.n; ++i)
      for (int j = 0; j < gf.W; ++j) {
        int ai, aj, bi, bj;
        if ((i & 1) && !(j & 1)) { ai = bi = (i - 1) / 2; aj = bj = j / 2; }
        else if (!(i & 1) && !(j & 1)) { ai = i / 2 - 1; aj = j / 2; bi = i / 2; bj = j / 2; }            // vertical edge
        else if ((i & 1) && (j & 1)) { ai = bi = (i - 1) / 2; aj = (j - 1) / 2; bj = (j + 1) / 2; }        // horizontal edge
        else { ai = i / 2 - 1; aj = (j + 1) / 2; bi = i / 2; bj = (j - 1) / 2; }                          // diagonal edge
        int pf = gf.vid(k, i, j);
        int s2[2] = {gc.source(k, ai, aj), gc.source(k, bi, bj)};
        for (int e = 0; e < 2; ++e) {
          int v = s2[e];
          if (v < 0) { gin_set_error("upsample plan: reads a dead cell"); return false; }
          if (v >= gc.P) {
            int pole = v - gc.P;
            fwd[(size_t)pf * 2 + e] = (corner_mode == GIN_CORNER_AVERAGE) ? -2 - pole : GIN_SRC_ZERO;
            if (corner_mode == GIN_CORNER_AVERAGE) pole_users[pole].push_back(pf);
          } else {
            fwd[(size_t)pf * 2 + e] = v;
            adj[v].push_back({pf, 0.5f});
          }
        }
      }
  int32_t ring[10];
  gc.rings(ring);
  for (int pole = 0; pole < 2; ++pole)
    for (int j = 0; j < 5; ++j)
      for (int pf : pole_users[pole]) adj[ring[5 * pole + j]].push_back({pf, 0.1f});
  // merge duplicate fine pixels (coarse copy appears twice with 0.5 each)
  int deg = 0;
  for (auto& a : adj) {
    std::map<int, float> m;
    for (auto& e : a) m[e.first] += e.second;
    a.assign(m.begin(), m.end());
    deg = std::max(deg, (int)a.size());
  }
  GinUpPlanHdr h;
  std::memset(&h, 0, sizeof(h));
  blob.assign(sizeof(h) / 4, 0);
  h.fwd_off = (int)blob.size();
  blob.insert(blob.end(), fwd.begin(), fwd.end());
  h.ring_off = (int)blob.size();
  blob.insert(blob.end(), ring, ring + 10);
  h.bwd_deg = deg;
  h.bwd_idx_off = (int)blob.size();
  blob.resize(blob.size() + (size_t)gc.P * deg, -1);
  h.bwd_w_off = (int)blob.size();
  blob.resize(blob.size() + (size_t)gc.P * deg, 0);
  for (int c = 0; c < gc.P; ++c)
    for (size_t e = 0; e < adj[c].size(); ++e) {
      blob[(size_t)h.bwd_idx_off + (size_t)c * deg + e] = adj[c][e].first;
      float w = adj[c][e].second;
      std::memcpy(&blob[(size_t)h.bwd_w_off + (size_t)c * deg + e], &w, 4);
    }
  h.magic = GIN_MAGIC; h.kind = GIN_PLAN_UPSAMPLE; h.level = level; h.corner_mode = corner_mode;
  h.Pc = gc.P; h.Pf = gf.P; h.total_words = (int)blob.size();
  std::memcpy(blob.data(), &h, sizeof(h));
  return true;
}

// ---------------------------------------------------------------- loss plan -----------
bool build_loss(int level, std::vector<int32_t>& blob) {
  if (level < 0 || level > 9) { gin_set_error("loss plan: bad level %d", level); return false; }
  Geo g(level);
  std::vector<int32_t> nb;
  one_rings(g, nb);
  GinLossPlanHdr h;
  std::memset(&h, 0, sizeof(h));
  blob.assign(sizeof(h) / 4, 0);
  h.ring_off = (int)blob.size();
  blob.insert(blob.end(), nb.begin(), nb.end());
  int32_t ring[10];
  g.rings(ring);
  h.pole_off = (int)blob.size();
  blob.insert(blob.end(), ring, ring + 10);
  h.flag_off = (int)blob.size();
  blob.resize(blob.size() + g.P, 0);
  for (int j = 0; j < 5; ++j) { blob[h.flag_off + ring[j]] |= 1; blob[h.flag_off + ring[5 + j]] |= 2; }
  h.magic = GIN_MAGIC; h.kind = GIN_PLAN_LOSS; h.level = level; h.P = g.P; h.V = g.P + 2; h.total_words = (int)blob.size();
  std::memcpy(blob.data(), &h, sizeof(h));
  return true;
}

bool build_any(int kind, int level, int stride, int corner_mode, std::vector<int32_t>& blob) {
  switch (kind) {
    case GIN_PLAN_HEXCONV: return build_conv(level, stride, corner_mode, blob);
    case GIN_PLAN_UPSAMPLE: return build_upsample(level, corner_mode, blob);
    case GIN_PLAN_LOSS: return build_loss(level, blob);
  }
  gin_set_error("unknown plan kind %d", kind);
  return false;
}

}  // namespace

extern "C" {

int gin_index_map_len(int level) {
  if (level < 0 || level > 12) return 0;
  Geo g(level);
  return 5 * (g.n + 2) * (g.W + 2);
}

int gin_index_map(int level, int32_t* out) {
  if (level < 0 || level > 12 || !out) { gin_set_error("gin_index_map: bad argument"); return GIN_ERR_ARG; }
  Geo g(level);
  size_t o = 0;
  for (int k = 0; k < 5; ++k)
    for (int i = -1; i <= g.n; ++i)
      for (int j = -1; j <= g.W; ++j) out[o++] = g.source(k, i, j);
  return GIN_OK;
}

int gin_ico_faces_len(int level) { return (level < 0 || level > 9) ? 0 : 60 << (2 * level); }

int gin_ico_faces(int level, int32_t* out) {
  if (level < 0 || level > 9 || !out) { gin_set_error("gin_ico_faces: bad argument"); return GIN_ERR_ARG; }
  std::vector<int32_t> f;
  faces(Geo(level), f);
  if ((int)f.size() != gin_ico_faces_len(level)) { gin_set_error("gin_ico_faces: face count %zu", f.size()); return GIN_ERR_PLAN; }
  std::memcpy(out, f.data(), f.size() * 4);
  return GIN_OK;
}

int gin_ico_vertices(int level, float* out) {
  if (level < 0 || level > 9 || !out) { gin_set_error("gin_ico_vertices: bad argument"); return GIN_ERR_ARG; }
  std::vector<double> v;
  vertices(Geo(level), v);
  for (size_t i = 0; i < v.size(); ++i) out[i] = (float)v[i];
  return GIN_OK;
}

size_t gin_plan_bytes(int kind, int level, int stride, int corner_mode) {
  std::vector<int32_t> blob;
  if (!build_any(kind, level, stride, corner_mode, blob)) return 0;
  return blob.size() * 4;
}

int gin_plan_build(int kind, int level, int stride, int corner_mode, void* host_buf, size_t bytes) {
  std::vector<int32_t> blob;
  if (!host_buf) { gin_set_error("gin_plan_build: null buffer"); return GIN_ERR_ARG; }
  if (!build_any(kind, level, stride, corner_mode, blob)) return GIN_ERR_ARG;
  if (bytes < blob.size() * 4) { gin_set_error("gin_plan_build: buffer too small (%zu < %zu)", bytes, blob.size() * 4); return GIN_ERR_ARG; }
  std::memcpy(host_buf, blob.data(), blob.size() * 4);
  return GIN_OK;
}

}  // extern "C"
