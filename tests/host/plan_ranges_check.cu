// Host-side check of the range plan shared by msda_bwd_vec_kernel (hit masks) and msda_scatter_mma2_kernel (work units):
// every pixel of the owned levels lies in exactly one range, the scatter kernel's pixel -> range formula
// (rbase[l] + offset / kR2Px) names that range, and [lv0, lv1) lists exactly the levels intersecting it.
// Usage: plan_ranges_check max_levels H0 W0 H1 W1 ...   (prints a JSON summary; exit code 1 on any violation)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../ziragroundingdino_b200/csrc/msda_common.cuh"

int main(int argc, char** argv) {
  if (argc < 4 || (argc - 2) % 2) return 2;
  const int max_levels = atoi(argv[1]);
  const int L = (argc - 2) / 2;
  int H[MSDA_MAX_LEVELS], W[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS], S = 0;
  for (int l = 0; l < L; ++l) { H[l] = atoi(argv[2 + 2 * l]); W[l] = atoi(argv[3 + 2 * l]); start[l] = S; S += H[l] * W[l]; }
  msda::RangePlan p;
  msda::plan_ranges(p, H, W, start, L, S, max_levels);
  int bad = 0;
  if (p.nranges > msda::kR2MaxRanges) ++bad;
  std::vector<int> owner(S, -1);
  for (int r = 0; r < p.nranges; ++r) {
    if (p.hi[r] - p.lo[r] <= 0 || p.hi[r] - p.lo[r] > msda::kR2Px) ++bad;
    for (int px = p.lo[r]; px < p.hi[r]; ++px) { if (owner[px] != -1) ++bad; owner[px] = r; }
  }
  for (int l = 0; l < L; ++l)
    for (int o = 0; o < H[l] * W[l]; ++o) {
      const int px = start[l] + o;
      if (l < p.first_level) { if (owner[px] != -1) ++bad; continue; }
      const int rid = p.rbase[l] + o / msda::kR2Px;
      if (owner[px] != rid) ++bad;
      if (!(p.lv0[rid] <= l && l < p.lv1[rid])) ++bad;
    }
  for (int r = 0; r < p.nranges; ++r)
    for (int l = p.lv0[r]; l < p.lv1[r]; ++l)
      if (!(start[l] < p.hi[r] && start[l] + H[l] * W[l] > p.lo[r])) ++bad;     // listed level really intersects
  const int tail = msda::coarse_first_level(H, W, start, L, S);
  printf("{\"first_level\": %d, \"nranges\": %d, \"bad\": %d, \"tail_first\": %d}\n", p.first_level, p.nranges, bad, tail);
  return bad ? 1 : 0;
}
