// phmm_registry.cpp — gathers the per-translation-unit kernel tables and picks a class
// (lanes per read G, rows per lane R) for a read length.
#include "phmm_registry.h"

#include <mutex>
#include <vector>

namespace fcsphmm {

extern const KernelEntry kEntriesF32G4[], kEntriesF32G8[], kEntriesF32G16[], kEntriesF32G32[];
extern const KernelEntry kEntriesF32UG4[], kEntriesF32UG8[], kEntriesF32UG16[], kEntriesF32UG32[];
extern const KernelEntry kEntriesF64G4[], kEntriesF64G8[], kEntriesF64G16[], kEntriesF64G32[];

namespace {
std::vector<KernelEntry> g_table;
std::vector<const KernelEntry*> g_sel[4];  // [f64 * 2 + ug] by read length
std::once_flag g_once;
constexpr int kMaxSelLen = 1024;

// Issue slots per useful cell: 8 FMA-pipe instructions + per-step overhead spread over the
// R rows of a lane, stretched by the wavefront fill/drain (G-1 extra steps on a ~300-column
// haplotype) and by the rows of the tile the read does not use.
double class_cost(const KernelEntry& k, int rows_needed) {
  const int esz = k.f64 ? 8 : 4;
  const double nv = (k.R * esz + 15) / 16;
  const double per_cell = 8.0 * (k.f64 ? 2.0 : 1.0) + (7.5 + nv) / k.R;
  const double skew = 1.0 + (k.G - 1) / 300.0;
  (void)rows_needed;
  return (double)(k.G * k.R) * per_cell * skew;  // issue slots per read and haplotype column, x32
}

void build() {
  const KernelEntry* lists[] = {kEntriesF32G4,  kEntriesF32G8,  kEntriesF32G16,  kEntriesF32G32,
                                kEntriesF32UG4, kEntriesF32UG8, kEntriesF32UG16, kEntriesF32UG32,
                                kEntriesF64G4,  kEntriesF64G8,  kEntriesF64G16,  kEntriesF64G32};
  for (const KernelEntry* l : lists)
    for (; l->G != 0; ++l) g_table.push_back(*l);
  KernelEntry end = {0, 0, false, false, nullptr, nullptr, nullptr, 0};
  g_table.push_back(end);
  for (int f = 0; f < 4; ++f) {
    g_sel[f].assign(kMaxSelLen + 1, nullptr);
    for (int len = 1; len <= kMaxSelLen; ++len) {
      const KernelEntry* best = nullptr;
      double bc = 0;
      for (const KernelEntry& k : g_table) {
        if (k.G == 0 || k.f64 != (f >= 2) || k.ug != ((f & 1) == 1) || k.G * k.R < len + 1) continue;
        const double c = class_cost(k, len + 1);
        if (!best || c < bc) { best = &k; bc = c; }
      }
      g_sel[f][len] = best;
    }
  }
}
}  // namespace

const KernelEntry* kernel_table() {
  std::call_once(g_once, build);
  return g_table.data();
}

const KernelEntry* find_kernel(bool f64, bool ug, int G, int R) {
  for (const KernelEntry* k = kernel_table(); k->G != 0; ++k)
    if (k->f64 == f64 && k->ug == ug && k->G == G && k->R == R) return k;
  return nullptr;
}

const KernelEntry* select_kernel(bool f64, bool ug, int read_len) {
  kernel_table();
  if (read_len < 1 || read_len > kMaxSelLen) return nullptr;
  return g_sel[(f64 ? 2 : 0) + (ug ? 1 : 0)][read_len];
}

}  // namespace fcsphmm
