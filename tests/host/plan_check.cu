// Host-only check of the work decomposition of the two contractions (no GPU needed): for many shapes, every
// (dictionary-row group, frame tile) gets its whole K range exactly once from plan_c1 + decode_item, no work item is
// empty (an empty item would publish an accumulator nobody wrote), and the half-width tail items of contraction 2
// tile each frame tile exactly.  Built and run by tests/test_host_logic.py.
#include "../../exemplars_vc_b200/csrc/evc_common.cuh"
#include "../../exemplars_vc_b200/csrc/simt_kernels.cuh"
#include "../../exemplars_vc_b200/csrc/tc_kernels.cuh"

#include <cstdio>
#include <map>
#include <random>
#include <vector>

using namespace evc;
using namespace evc::tc;

static int fails = 0;
#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      if (fails < 20) { printf("FAIL %s: ", #cond); printf(__VA_ARGS__); printf("\n"); } \
      ++fails;                                            \
    }                                                     \
  } while (0)

// contraction 1: the parameters contract_wh_t derives from the plan
// `pairs` CTA pairs per cluster share the dictionary tile and take neighbouring frame tiles: the plan is made over
// frame super-tiles and `slots` clusters (B200: 74 / 33 / 15 for 1 / 2 / 4 pairs, tools/probe/cluster_probe.cu)
static void check_c1(int F_main, int N, int T, int bke, int pairs = 1, int slots = 0) {
  const C1Plan pl = plan_c1(F_main, N, T, bke, pairs, slots);
  CHECK(pl.t_tiles == ceil_div(ceil_div(T, kC1BlockT), pairs), "super-tiles");
  CHECK(slots == 0 || (long long)pl.t_tiles * ((pl.splits_last ? pl.m_groups - 1 : pl.m_groups) * pl.splits + pl.splits_last) <= slots ||
            (pl.splits <= 1 && pl.splits_last <= 1), "one wave of clusters");
  GemmParams p{};
  p.M_total = F_main; p.T = T; p.K = N;
  p.num_m_groups = pl.m_groups; p.num_t_tiles = pl.t_tiles; p.num_splits = pl.splits;
  p.kblocks_per_split = pl.kb_per_split; p.kblocks_total = pl.kb_total;
  p.splits_last = pl.splits_last; p.kblocks_per_split_last = pl.kb_per_split_last;
  p.items_main = (pl.splits_last ? pl.m_groups - 1 : pl.m_groups) * pl.t_tiles * pl.splits;
  p.half_from = p.items_main; p.tail_parts = 1;
  const int items = num_items_of(p);
  CHECK(items > 0, "F=%d N=%d T=%d", F_main, N, T);
  CHECK(pl.kb_total == ceil_div(N, bke), "kb_total");
  std::map<std::pair<int, int>, std::vector<char>> cover;
  for (int it = 0; it < items; ++it) {
    const WorkItem w = decode_item(p, it, kC1BlockT);
    CHECK(w.kb0 < w.kb1, "empty item %d (F=%d N=%d T=%d bk=%d): kb [%d,%d)", it, F_main, N, T, bke, w.kb0, w.kb1);
    CHECK(w.m_group >= 0 && w.m_group < pl.m_groups && w.t_tile >= 0 && w.t_tile < pl.t_tiles, "range");
    CHECK(w.t_off == 0 && w.t_cols == kC1BlockT, "full-width items only");
    auto& v = cover[{w.m_group, w.t_tile}];
    v.resize(pl.kb_total, 0);
    for (int kb = w.kb0; kb < w.kb1 && kb < pl.kb_total; ++kb) v[kb]++;
    // the split index addresses the partial buffer: it must stay below the buffer's depth
    CHECK(w.split < pl.max_splits, "split %d >= max_splits %d", w.split, pl.max_splits);
  }
  CHECK((int)cover.size() == pl.m_groups * pl.t_tiles, "tiles covered %d of %d", (int)cover.size(), pl.m_groups * pl.t_tiles);
  for (auto& kv : cover)
    for (int kb = 0; kb < pl.kb_total; ++kb)
      CHECK(kv.second[kb] == 1, "K-block %d of tile (%d,%d) covered %d times (F=%d N=%d T=%d bk=%d)", kb, kv.first.first,
            kv.first.second, (int)kv.second[kb], F_main, N, T, bke);
  // rows of the last group start where the reduction expects them
  const int sub_rows = 128 * cta_group();
  CHECK(pl.f_last == F_main || (pl.f_last < F_main && pl.f_last % sub_rows == 0), "f_last %d (F_main %d)", pl.f_last, F_main);
}

// contraction 2: the parameters contract2_cg derives (no split-K, half-width tail items)
static void check_c2(int N, int T, int F, int bke) {
  const int kCG = cta_group();
  GemmParams p{};
  p.M_total = N; p.T = T; p.K = F;
  p.num_m_groups = ceil_div(N, 128 * kC2MTiles * kCG); p.num_t_tiles = ceil_div(T, kC2BlockT); p.num_splits = 1;
  p.kblocks_total = ceil_div(F, bke); p.kblocks_per_split = p.kblocks_total;
  p.items_main = p.num_m_groups * p.num_t_tiles; p.splits_last = 0; p.kblocks_per_split_last = 0;
  const int slots = num_sms() / kCG;
  plan_tail(p, slots, kC2BlockT, true);
  for (int m_fastest = 0; m_fastest < 2; ++m_fastest) {
    p.m_fastest = m_fastest;
    const int items = num_items_of(p);
    std::map<std::pair<int, int>, int> cols;
    std::map<std::pair<int, int>, unsigned> chunk_mask;
    for (int it = 0; it < items; ++it) {
      const WorkItem w = decode_item(p, it, kC2BlockT);
      CHECK(w.kb0 == 0 && w.kb1 == p.kblocks_total, "whole K per item");
      CHECK(w.m_group >= 0 && w.m_group < p.num_m_groups && w.t_tile >= 0 && w.t_tile < p.num_t_tiles, "range");
      CHECK(w.t_cols >= 64 && w.t_cols <= kC2BlockT && w.t_cols % 32 == 0, "t_cols %d", w.t_cols);
      CHECK(w.t_off % 32 == 0 && w.t_off + w.t_cols <= kC2BlockT, "t_off %d", w.t_off);
      CHECK(w.t_cols % kHChunkT == 0, "items are whole H chunks");
      cols[{w.m_group, w.t_tile}] += w.t_cols;
      for (int c = w.t_off / 32; c < (w.t_off + w.t_cols) / 32; ++c) {  // every chunk of a tile exactly once
        CHECK(!(chunk_mask[{w.m_group, w.t_tile}] & (1u << c)), "chunk %d of tile (%d,%d) twice", c, w.m_group, w.t_tile);
        chunk_mask[{w.m_group, w.t_tile}] |= 1u << c;
      }
    }
    CHECK((int)cols.size() == p.items_main, "tiles covered %d of %d (N=%d T=%d)", (int)cols.size(), p.items_main, N, T);
    for (auto& kv : cols) CHECK(kv.second == kC2BlockT, "tile (%d,%d) got %d frame columns", kv.first.first, kv.first.second, kv.second);
    // balance: no CTA (pair) carries more than one tile above the mean
    std::vector<double> load(slots, 0.0);
    for (int it = 0; it < items; ++it) load[it % slots] += decode_item(p, it, kC2BlockT).t_cols / (double)kC2BlockT;
    double mx = 0, sum = 0;
    for (double l : load) { mx = std::max(mx, l); sum += l; }
    CHECK(mx <= sum / slots + 1.0, "imbalance max %.2f mean %.2f", mx, sum / slots);
  }
}

int main() {
  std::mt19937 rng(20190123);
  const int bkes[2] = {32, 64};  // K elements per K-block: split / tf32, bf16
  // the BASELINE shapes and their neighbours, then random ones
  const int fixed[][3] = {{513, 20000, 1000}, {512, 20000, 1000}, {513, 200000, 2000}, {2565, 50000, 1000}, {513, 20000, 129857},
                          {13, 32, 8}, {201, 777, 37}, {513, 768, 64}, {1, 5, 1}, {128, 128, 256}, {129, 129, 257},
                          {640, 4100, 300}, {513, 25000, 2000}, {1024, 16, 3}};
  int n = 0;
  for (auto& s : fixed)
    for (int bke : bkes) {
      const int F = s[0], N = s[1], T = s[2];
      const int n_left = (F > 128 && (F % 128) <= 8) ? F % 128 : 0;
      check_c1(F - n_left, N, T, bke); check_c2(N, T, F, bke); ++n;
      check_c1(F - n_left, N, T, bke, 2, 33); check_c1(F - n_left, N, T, bke, 4, 15);
    }
  for (int i = 0; i < 3000; ++i) {
    const int F = 1 + (int)(rng() % 3000), N = 1 + (int)(rng() % (i % 7 == 0 ? 300000 : 30000)), T = 1 + (int)(rng() % (i % 11 == 0 ? 140000 : 3000));
    const int bke = bkes[rng() % 2];
    const int n_left = (F > 128 && (F % 128) <= 8) ? F % 128 : 0;
    if (F - n_left < 1) continue;
    check_c1(F - n_left, N, T, bke); check_c2(N, T, F, bke); ++n;
    if (i % 3 == 0) { check_c1(F - n_left, N, T, bke, 2, 33); check_c1(F - n_left, N, T, bke, 4, 15); }
  }
  printf("plan_check: %d shapes, %d failures (cta_group %d, %d SMs assumed)\n", n, fails, cta_group(), num_sms());
  return fails ? 1 : 0;
}
