// Host check of the persistent tile schedule (dicp_b200/csrc/tile_stream.cuh): TEST INFRASTRUCTURE ONLY.
// Walks every CTA's TileCursor and verifies that each tile is visited exactly once, that every (chunk, pair)
// segment gets its own row below plan.rows, that rows of a pair are used contiguously from 0 (the segment ending
// the pair zero-fills the rest), and returns the schedule's imbalance.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../dicp_b200/csrc/tile_stream.cuh"

static int g_sms = 148;
namespace dicp { int device_sm_count() { return g_sms; } }

extern "C" int dicp_plan_check(int N, int n, int tile, int ctas_per_sm, int k, int sms, int* out /* [ctas, rows, max_tiles, min_tiles] */) {
  using namespace dicp;
  g_sms = sms;
  const TilePlan p = make_tile_plan(N, n, tile, ctas_per_sm, k);
  std::vector<unsigned char> seen((size_t)p.tiles, 0);
  std::vector<int> rows_used((size_t)N * p.rows, 0);
  std::vector<int> rows_count(N, -1);
  int max_t = 0, min_t = 1 << 30;
  for (int c = 0; c < p.ctas; ++c) {
    TileWalk walk;
    walk.init(p, c);
    const int nt = walk.ntiles;
    if (nt > max_t) max_t = nt;
    if (nt < min_t) min_t = nt;
    TileCursor cur;
    int seg_b = -1;
    long long seg_tile = 0, last = -1;
    auto flush = [&]() -> int {
      const int row = tp_segment_row(p, seg_b, seg_tile);
      if (row < 0 || row >= p.rows) return 1;
      if (rows_used[(size_t)seg_b * p.rows + row]++) return 2;
      if (last == (long long)seg_b * p.tpp + p.tpp - 1) rows_count[seg_b] = row + 1;
      return 0;
    };
    for (cur.init(p, walk); cur.L < nt; cur.next(p, walk)) {
      const long long tile = cur.tile(p);
      if (tile < 0 || tile >= p.tiles) return 3;
      if (cur.j < 0 || cur.j >= p.tpp || cur.b < 0 || cur.b >= N) return 4;
      if (seen[(size_t)tile]++) return 5;
      if (cur.seg_first) {
        if (cur.L > 0) { int rc = flush(); if (rc) return rc; }
        seg_b = cur.b; seg_tile = tile;
      } else if (cur.b != seg_b || tile != last + 1) return 6;
      last = tile;
    }
    if (nt > 0) { int rc = flush(); if (rc) return rc; }
  }
  for (long long t = 0; t < p.tiles; ++t) if (seen[(size_t)t] != 1) return 7;
  for (int b = 0; b < N; ++b) {
    if (rows_count[b] < 1) return 8;
    for (int r = 0; r < p.rows; ++r)
      if (rows_used[(size_t)b * p.rows + r] != (r < rows_count[b] ? 1 : 0)) return 9;   // contiguous from 0; the rest is zero-filled
  }
  out[0] = p.ctas; out[1] = p.rows; out[2] = max_t; out[3] = min_t;
  return 0;
}
