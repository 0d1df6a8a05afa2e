// phmm_registry.cpp — gathers the tier kernels and picks a class (lanes per read G, rows per
// lane R) for a read length.
#include "phmm_registry.h"

#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <vector>

namespace fcsphmm {

extern const TierKernel kTierF32T0, kTierF32T1, kTierF32T2, kTierF32UT0, kTierF32UT1, kTierF32UT2, kTierF32AT1, kTierF32AT2, kTierF32PT1, kTierF32PT2;
extern const TierKernel kTierF64T0, kTierF64T1, kTierF64T2, kTierF64UT0, kTierF64UT1, kTierF64UT2;

namespace {
const TierKernel* g_kernels[] = {&kTierF32T0, &kTierF32T1, &kTierF32T2, &kTierF32UT0, &kTierF32UT1, &kTierF32UT2, &kTierF32AT1, &kTierF32AT2, &kTierF32PT1, &kTierF32PT2,
                                 &kTierF64T0, &kTierF64T1, &kTierF64T2, &kTierF64UT0, &kTierF64UT1, &kTierF64UT2};
constexpr int kNumKernels = 16;
constexpr int kForms = 4;  // general, uniform-GCP, all-uniform, haplotype pairs (uniform GCP)
inline int fidx(bool f64, int form) { return (f64 ? kForms : 0) + form; }
constexpr int kMaxSelLen = 1024;
std::vector<ClassRef> g_classes[2 * kForms];    // [fidx(f64, form)]
std::vector<const ClassRef*> g_sel[2 * kForms];  // by read length
std::vector<const ClassRef*> g_sel_coarse[2 * kForms];  // same, restricted to the coarse grid of rows per lane
// Coarse grid: rows per lane up to 8, then multiples of 4.  A ragged chunk that spreads its reads over every
// row count runs dozens of different unrolled loop bodies at once (several chunks and tiers are in flight):
// config 3 end to end went from 1.6 to 2.2 TCUPS on the coarse grid although the tiles sweep ~5 % more
// cells.  The batcher therefore uses a fine class only for read lengths that are popular in the chunk.
inline bool on_coarse_grid(int R) {
  static const int step = [] { const char* e = std::getenv("FCS_PHMM_COARSE_STEP"); const int v = e ? std::atoi(e) : 4; return v >= 1 ? v : 4; }();  // developer knob
  return R <= 8 || (R % step) == 0;
}
std::vector<std::pair<int, int>> g_f64_queues;  // (G, R) of the general-form FP64 classes
std::once_flag g_once;

// Issue slots per read and haplotype column (x32): 8 FMA-pipe instructions per cell (x2 in double)
// plus the per-step overhead spread over the R rows of a lane, stretched by the wavefront
// fill/drain (G-1 extra steps on a ~300-column haplotype); all G*R rows of the tile are paid for.
// issue slots of one wavefront step of a warp: R cells x 8 FMA-pipe instructions + per-step overhead
double step_cost(bool f64, int R) {
  const int esz = f64 ? 8 : 4;
  const double nv = (R * esz + 15) / 16;
  return 8.0 * (f64 ? 2.0 : 1.0) * R + 7.5 + nv;
}
double class_cost(bool f64, int G, int R) { return step_cost(f64, R) * G * (1.0 + (G - 1) / 300.0); }

void build() {
  for (int i = 0; i < kNumKernels; ++i) {
    const TierKernel* tk = g_kernels[i];
    auto& v = g_classes[fidx(tk->f64, tk->form)];
    for (int c = 0; c < tk->n_classes; ++c) v.push_back(ClassRef{tk, c, tk->classes[c].G, tk->classes[c].R, nullptr});
  }
  for (int f = 0; f < 2 * kForms; ++f) {
    g_sel[f].assign(kMaxSelLen + 1, nullptr);
    for (int len = 1; len <= kMaxSelLen; ++len) {
      const ClassRef* best = nullptr;
      double bc = 0;
      for (const ClassRef& k : g_classes[f]) {
        if (k.G * k.R < len + 1) continue;
        const double c = class_cost(f >= kForms, k.G, k.R);
        if (!best || c < bc) { best = &k; bc = c; }
      }
      g_sel[f][len] = best;
    }
    g_sel_coarse[f].assign(kMaxSelLen + 1, nullptr);
    for (int len = 1; len <= kMaxSelLen; ++len) {
      const ClassRef* best = nullptr;
      double bc = 0;
      for (const ClassRef& k : g_classes[f]) {
        if (k.G * k.R < len + 1 || !on_coarse_grid(k.R)) continue;
        const double c = class_cost(f >= kForms, k.G, k.R);
        if (!best || c < bc) { best = &k; bc = c; }
      }
      g_sel_coarse[f][len] = best ? best : g_sel[f][len];
    }
  }
  for (const ClassRef& k : g_classes[fidx(true, 0)]) g_f64_queues.emplace_back(k.G, k.R);
  for (int f = 0; f < 2 * kForms; ++f) {
    if (f % kForms >= 2) continue;  // the all-uniform and the haplotype-pair form have their own class lists, no twin
    for (ClassRef& k : g_classes[f])
      for (const ClassRef& o : g_classes[f - f % kForms + (1 - f % kForms)])
        if (o.G == k.G && o.R == k.R) { k.twin = &o; break; }
  }
}
}  // namespace

const TierKernel* const* tier_kernels(int* n) {
  std::call_once(g_once, build);
  if (n) *n = kNumKernels;
  return g_kernels;
}

const ClassRef* select_class(bool f64, int form, int read_len, bool coarse) {
  std::call_once(g_once, build);
  if (read_len < 1 || read_len > kMaxSelLen) return nullptr;
  return (coarse ? g_sel_coarse : g_sel)[fidx(f64, form)][read_len];
}

const ClassRef* select_class_for(bool f64, int form, int read_len, int n_reads, int avg_hap_len, bool coarse) {
  std::call_once(g_once, build);
  if (read_len < 1 || read_len > kMaxSelLen) return nullptr;
  // memo: the choice depends on the haplotype length only weakly -> four length bins; benign races
  // (every thread computes the same pointer)
  static std::atomic<const ClassRef*> memo[2][2 * kForms][kMaxSelLen + 1][8][4];
  const int nb = n_reads < 1 ? 1 : (n_reads > 7 ? 7 : n_reads);
  const int hb = avg_hap_len < 150 ? 0 : (avg_hap_len < 300 ? 1 : (avg_hap_len < 600 ? 2 : 3));
  static const int hb_len[4] = {100, 220, 420, 900};
  std::atomic<const ClassRef*>& slot = memo[coarse ? 1 : 0][fidx(f64, form)][read_len][nb][hb];
  if (const ClassRef* hit = slot.load(std::memory_order_relaxed)) return hit;
  n_reads = nb;
  avg_hap_len = hb_len[hb];
  const ClassRef* best = nullptr;
  double bc = 0;
  const double lh = avg_hap_len > 0 ? avg_hap_len : 300;
  for (const ClassRef& k : g_classes[fidx(f64, form)]) {
    if (k.G * k.R < read_len + 1 || (coarse && !on_coarse_grid(k.R))) continue;
    const int served = std::min(n_reads, 32 / k.G);
    const double c = step_cost(f64, k.R) * (lh + k.G - 1) / served;
    if (!best || c < bc * 0.999) { best = &k; bc = c; }
  }
  slot.store(best, std::memory_order_relaxed);
  return best;
}

const ClassRef* select_class_wide(bool f64, int form, int read_len, int min_G) {
  std::call_once(g_once, build);
  if (read_len < 1 || read_len > kMaxSelLen) return nullptr;
  // memo per (form, length, min_G in {<=4, 8, 16, 32}); benign races (every thread computes the same pointer)
  static std::atomic<const ClassRef*> memo[2 * kForms][kMaxSelLen + 1][4];
  const int gb = min_G <= 4 ? 0 : (min_G <= 8 ? 1 : (min_G <= 16 ? 2 : 3));
  std::atomic<const ClassRef*>& slot = memo[fidx(f64, form)][read_len][gb];
  if (const ClassRef* hit = slot.load(std::memory_order_relaxed)) return hit;
  min_G = gb == 0 ? min_G : (4 << gb);
  const ClassRef* best = nullptr;
  for (const ClassRef& k : g_classes[fidx(f64, form)]) {
    if (k.G * k.R < read_len + 1 || k.G < min_G) continue;
    // fewest rows per lane first (shortest dependent chain per step), then fewest lanes
    if (!best || k.R < best->R || (k.R == best->R && k.G < best->G)) best = &k;
  }
  if (!best) best = select_class(f64, form, read_len);
  slot.store(best, std::memory_order_relaxed);  // (every class has G >= 4, so all min_G <= 4 are equivalent)
  return best;
}

const ClassRef* find_class(bool f64, int form, int G, int R) {
  std::call_once(g_once, build);
  for (const ClassRef& k : g_classes[fidx(f64, form)])
    if (k.G == G && k.R == R) return &k;
  return nullptr;
}

int f64_queue_count() {
  std::call_once(g_once, build);
  return (int)g_f64_queues.size();
}

int f64_queue_id(int G, int R) {
  std::call_once(g_once, build);
  for (size_t i = 0; i < g_f64_queues.size(); ++i)
    if (g_f64_queues[i].first == G && g_f64_queues[i].second == R) return (int)i;
  return -1;
}

const ClassRef* f64_queue_class(int qid, int form) {
  std::call_once(g_once, build);
  return find_class(true, form, g_f64_queues[qid].first, g_f64_queues[qid].second);
}

}  // namespace fcsphmm
