// phmm_generic.cuh — striped (multi-pass) PairHMM kernel for shapes the single-pass tiles do not
// cover: reads longer than 32 lanes x R rows, or haplotypes too long for a shared-memory stream.
//
// Same tile arithmetic as phmm_kernel.cuh (Tile::step), one pair per warp, G = 32 lanes:
//  * the read is cut into stripes of 32*R rows, bottom-aligned (the first stripe starts with the
//    boundary-replica rows); stripes run one after the other over the whole haplotype;
//  * the bottom row of a stripe (M, X, Y per column) goes to a per-CTA scratch row in global memory
//    and comes back as the top boundary of the next stripe: every 32 steps the warp loads 32
//    consecutive columns coalesced, and lane 0 picks its column with a shuffle (in place: the
//    write index trails the read index by at least 31 columns);
//  * haplotype symbols are not staged in shared memory: a coalesced 32-column load every 32 steps,
//    then a one-lane-per-step shift register built from shuffles.
// Throughput is below the single-pass kernels (11 shuffles and ~8 selects per step); it exists so
// that no shape is refused.  Bit-identical results (same statement order per cell).
#pragma once
#include "phmm_kernel.cuh"

namespace fcsphmm {


template <typename T>
struct GenericCfg {
  static constexpr int G = 32;
  static constexpr int R = sizeof(T) == 4 ? 16 : 8;
  static constexpr int ROWS = G * R;
  static constexpr int STRIDE = tab_stride_bytes(R, (int)sizeof(T));
  static constexpr int HSCALE = 32 * STRIDE / 16;
};

// dynamic shared memory of the striped kernel: the prior table (the N row is always present here)
template <typename T>
inline size_t generic_smem_bytes(uint32_t n_sym) { return (size_t)(n_sym < (uint32_t)kTabRows ? (uint32_t)kTabRows : n_sym) * 32u * GenericCfg<T>::STRIDE; }

template <typename T, bool FROM_QUEUE>
__global__ void __launch_bounds__(32, 8) phmm_generic(const __grid_constant__ KParams p) {
  using C = GenericCfg<T>;
  using A = Ar<T>;
  constexpr int R = C::R;
  extern __shared__ __align__(128) uint8_t tab[];  // n_sym symbol rows of 32 * STRIDE bytes (generic_smem_bytes)
  __shared__ T lut[128];
  const int lane = threadIdx.x;
  uint8_t* tab_lane = tab + lane * C::STRIDE;
  const T* __restrict__ mm = reinterpret_cast<const T*>(p.mm);
  for (int i = lane; i < 128; i += 32) lut[i] = reinterpret_cast<const T*>(p.ph2pr)[i];
  __syncwarp();
  const uint32_t count = FROM_QUEUE ? p.rerun_count[kQueueGenericF64] : p.gen_count;
  const RerunEntry* list = FROM_QUEUE ? p.rerun + p.rerun_base[kQueueGenericF64] : p.gen_list;
  const size_t Lc = p.scratch_cols;
  T* bndM = reinterpret_cast<T*>(p.scratch) + (size_t)blockIdx.x * 3u * Lc;
  T* bndX = bndM + Lc;
  T* bndY = bndX + Lc;
  Tile<T, 32, R, false> tile;
  tile.cXX = T(0);
  tile.cGM = T(0);
  using State = typename Tile<T, 32, R, false>::State;

  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
    const RerunEntry e = list[i];
    const ReadMeta rm = p.rmeta[e.read];
    const HapMeta hm = p.hmeta[e.hap];
    const int Lr = (int)read_len_of(rm), Lh = (int)hm.len;
    const uint32_t layout = read_layout(rm);
    const uint8_t* rs = p.reads + (size_t)rm.data_off16 * 16u;
    const uint8_t* hap = p.haps + (size_t)hm.data_off16 * 16u;
    const int P = (Lr + 1 + C::ROWS - 1) / C::ROWS;
    const int npad = P * C::ROWS - Lr;
    const T y_init = A::div(A::K(), (T)Lh);
    const int nsteps = Lh + 31;
    T result = T(0);
    for (int s = 0; s < P; ++s) {
      __syncwarp();
      tile.build(rs, (uint32_t)Lr, lane, lut, mm, tab_lane, true, layout, s * C::ROWS, npad);
      if (p.n_sym > (uint32_t)kCodeOther) Tile<T, 32, R, false>::build_other_rows(rs, (uint32_t)Lr, lane, lut, tab_lane, p.n_sym, p.extra_bytes, tile.off_last, s * C::ROWS, npad);
      __syncwarp();
      State st;
      tile.init(st, y_init);
      // column 0 of the row above this stripe: K/Lh if that row is a boundary replica (Y[0][0] = K/Lh), else 0
      if (lane == 0 && s > 0) st.dY = (s * C::ROWS - 1 < npad) ? y_init : T(0);
      uint32_t sym = (uint32_t)(kCodePad * C::HSCALE), symbuf = sym;
      T bM = T(0), bX = T(0), bY = T(0), oM = T(0), oX = T(0), oY = T(0);
      const bool has_top = s > 0, has_bottom = s < P - 1;
#pragma unroll 2
      for (int t = 0; t < nsteps; ++t) {
        if ((t & 31) == 0) {
          const int col = t + lane;
          int c = kCodePad;
          if (col < Lh) c = hap_code(hap[col], p.n_sym, p.extra_bytes);
          symbuf = (uint32_t)(c * C::HSCALE);
          if (has_top) {
            bM = col < Lh ? bndM[col] : T(0);
            bX = col < Lh ? bndX[col] : T(0);
            bY = col < Lh ? bndY[col] : T(0);
          }
        }
        const uint32_t nsym = __shfl_sync(0xffffffffu, symbuf, t & 31);
        const uint32_t usym = __shfl_up_sync(0xffffffffu, sym, 1);
        sym = lane == 0 ? nsym : usym;
        T uM = __shfl_up_sync(0xffffffffu, st.M[R - 1], 1);
        T uX = __shfl_up_sync(0xffffffffu, st.X[R - 1], 1);
        T uY = __shfl_up_sync(0xffffffffu, st.Y[R - 1], 1);
        if (has_top) {
          const T cM = __shfl_sync(0xffffffffu, bM, t & 31);
          const T cX = __shfl_sync(0xffffffffu, bX, t & 31);
          const T cY = __shfl_sync(0xffffffffu, bY, t & 31);
          if (lane == 0) { uM = cM; uX = cX; uY = cY; }
        }
        tile.step(st, tab_lane + sym * 16u, uM, uX, uY);
        if (has_bottom) {
          const int c2 = t - 31;  // column lane 31 just finished
          const T vM = __shfl_sync(0xffffffffu, st.M[R - 1], 31);
          const T vX = __shfl_sync(0xffffffffu, st.X[R - 1], 31);
          const T vY = __shfl_sync(0xffffffffu, st.Y[R - 1], 31);
          if (c2 >= 0 && lane == (c2 & 31)) { oM = vM; oX = vX; oY = vY; }
          if (c2 >= 0 && ((c2 & 31) == 31 || t == nsteps - 1)) {
            const int col = (c2 & ~31) + lane;
            if (col <= c2) { bndM[col] = oM; bndX[col] = oX; bndY[col] = oY; }
          }
        }
      }
      result = st.acc;
    }
    if (lane == 31) {
      if constexpr (sizeof(T) == 4) emit_f32(p, rm, e.read, e.hap, result);
      else emit_f64(p, rm, e.hap, result);
    }
  }
}

}  // namespace fcsphmm
