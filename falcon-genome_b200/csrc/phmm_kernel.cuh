// phmm_kernel.cuh — sm_100a PairHMM forward kernels (FP32 wavefront + FP64 rerun).
//
// Replaces (behind the C ABI in include/fcs_pairhmm.h) the native computeLikelihoodsNative
// loop that the GATK JVMs spawned by /root/reference/src/workers/HTCWorker.cpp:48-113 and
// src/workers/Mutect2Worker.cpp:109-192 run; semantics per SURVEY.md Appendix A.
//
// Design (B200-first, not a translation of the AVX anti-diagonal code):
//  * One warp per CTA.  The warp is cut into 32/G lane groups; a group of G lanes owns ONE
//    read, each lane a register tile of R consecutive read rows (G*R >= len+1).  All groups
//    of the warp stream the SAME haplotypes, so one shared-memory haplotype stream feeds
//    the whole warp.
//  * Anti-diagonal wavefront across the G lanes: at step t lane l works on haplotype column
//    t-l.  Only the bottom row of a lane's tile crosses to the next lane: three
//    __shfl_up_sync per step (M, X, Y), amortised over R cells.
//  * Per-row transition terms (pMM, pGM, pMX, pXX, pMY) live in registers for the whole
//    task.  The per-cell prior (match ? 1-e : e/3) is NOT a compare+select: it is read
//    from a shared-memory table  prior[symbol][lane][row]  with LDS.128 (4 rows per
//    instruction), indexed by the haplotype symbol of the lane's current column.
//    The table lane stride is an odd multiple of 16 B and the symbol pitch a multiple of
//    128 B, so every quarter-warp phase is bank-conflict free whatever symbols the lanes see.
//  * No branch, predicate or boundary select in the inner loop.  Boundaries are data:
//      - rows above the read ("padding rows", tile top) carry constants that reproduce the
//        row-0 boundary exactly: M = X = 0, Y = K/Lh;
//      - columns outside the haplotype are the PAD symbol whose prior is 0, which makes
//        M and X exactly 0 there, so the running sum of the last row is unaffected.
//  * Reads, quals and haplotypes arrive in shared memory through cp.async.bulk (TMA bulk
//    copy, UBLKCP in SASS) completing on an mbarrier; ph2pr comes from a shared-memory LUT.
//  * Arithmetic is spelled with __fmul_rn/__fmaf_rn/__fadd_rn in the statement order of the
//    oracle's float twin, so the raw FP32 sums (and therefore the FP32->FP64 fallback
//    decisions) are bit-identical to it; same for the FP64 kernel vs the double oracle.
//
// Roofline: 8 FMA-pipe instructions per cell (4 FFMA + 4 FMUL); per step of R cells the
// overhead is 3 SHFL + 1 LDS.U16 + ceil(R/4) LDS.128 + 2 FADD + address/loop ~ 2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "phmm_types.h"

#ifndef PHMM_UNROLL_T
#define PHMM_UNROLL_T 4
#endif

namespace fcsphmm {

#ifndef PHMM_UNROLL2_ABOVE
#define PHMM_UNROLL2_ABOVE 24
#endif
constexpr int kUnrollT = PHMM_UNROLL_T;  // steps per loop trip (even, so the state arrays rotate without MOVs; 4 measured best: +4 % over 2)

// ----------------------------------------------------------------------------------------
// arithmetic with pinned rounding and no compiler contraction
template <typename T>
struct Ar;
template <>
struct Ar<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float K() { return 0x1p120f; }
};
template <>
struct Ar<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double K() { return 0x1p1020; }
};

// ----------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk) helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ASCII base -> symbol class.  Reads: anything outside ACGTN gets 7 (matches only an N hap).
__device__ __forceinline__ int base_code(uint32_t b) {
  int c = 7;
  c = (b == 'A') ? 0 : c;
  c = (b == 'C') ? 1 : c;
  c = (b == 'G') ? 2 : c;
  c = (b == 'T') ? 3 : c;
  c = (b == 'N') ? kCodeN : c;
  return c;
}

// Haplotype byte -> symbol row.  ACGT, N, then (only in launches whose haplotypes hold such bytes: n_sym > 6)
// the byte's own row if some read of the chunk contains the same byte, else the shared OTHER row.
__device__ __forceinline__ int hap_code(uint32_t b, uint32_t n_sym, uint64_t extra) {
  int c = base_code(b);
  if (c > kCodeN) {
    c = kCodeOther;
    for (uint32_t e = 0; e + (uint32_t)kCodeExtra0 < n_sym; ++e)
      if (b == (uint32_t)((extra >> (8u * e)) & 0xffu)) c = kCodeExtra0 + (int)e;
  }
  // the host sized the table from the caller's bytes before it copied them; if the caller's memory changed in
  // between (a shared daemon segment), stay inside the table instead of trusting the copy
  return (uint32_t)c < n_sym ? c : kCodePad;
}

// ----------------------------------------------------------------------------------------
// shared-memory layout of one CTA
//   [mbarrier 128 B][prior table: n_sym x 32 lanes x STRIDE][union][raw haplotype staging]
//   union = { ph2pr LUT + read staging }  (needed until the tile is built)
//         = { haplotype stream }          (written after the tile is built)
// The N symbol row exists only if a haplotype of the launch contains an N (n_sym = 6).  Together
// these keep the common 150-bp class at 12 one-warp CTAs per SM (8 for the all-uniform G=4, R=38 class,
// whose table alone is 24 KB).
template <typename T, int G, int R, bool LIST, int FORM = 0>
struct Layout {
  static constexpr int NG = 32 / G;
  static constexpr int ROWS = G * R;
  static constexpr int STRIDE = tab_stride_form(R, (int)sizeof(T), FORM);
  static constexpr int SYM_PITCH = 32 * STRIDE;     // bytes between symbol rows of the table
  static constexpr int HSCALE = SYM_PITCH / 16;     // value stored in the haplotype stream per symbol
  // read staging bytes per group; the all-uniform form reads bases and base qualities straight from
  // global memory when it builds its tile, so it stages nothing
  static constexpr int RSTAGE = FORM == 2 ? 0 : 5 * (int)round_up16(ROWS);
  static constexpr int OFF_BAR = 0;
  static constexpr int OFF_TAB = 128;
  static constexpr int LUT_BYTES = 128 * (int)sizeof(T);
  static constexpr int PRE_BYTES = LUT_BYTES + NG * RSTAGE;
  static_assert(OFF_TAB % 128 == 0, "table must start on a 128-byte line");
  PHMM_HD static constexpr uint32_t off_union(uint32_t n_sym) { return (uint32_t)OFF_TAB + n_sym * (uint32_t)SYM_PITCH; }
  // hs_cap / hap_stage_bytes are per group when LIST, per CTA otherwise
  PHMM_HD static constexpr uint32_t union_bytes(uint32_t hs_cap) {
    const uint32_t hs = (uint32_t)(LIST ? NG : 1) * round_up16(hs_cap * 2u);
    return hs > (uint32_t)PRE_BYTES ? hs : (uint32_t)PRE_BYTES;
  }
  static constexpr size_t smem_bytes(uint32_t hs_cap, uint32_t hap_stage_bytes, uint32_t n_sym) {
    return (size_t)off_union(n_sym) + union_bytes(hs_cap) + (size_t)(LIST ? NG : 1) * round_up16(hap_stage_bytes);
  }
};

// ----------------------------------------------------------------------------------------
// the register tile of one lane and its wavefront loop
//
// FORM 1, UG ("uniform GCP"): every read of the launch has one constant gap-continuation quality
// (GATK HaplotypeCaller / Mutect2 always pass a constant 10 [upstream, SURVEY A.6]).  Then
// pXX = pYY = ph2pr[gcp] and pGM = 1 - pXX are launch constants that the FMAs take from the
// constant bank instead of the register file.  The register file feeds about two 32-bit
// operands per clock per SM sub-partition (measured, tools/ubench/fp32_issue.cu), so the
// operand reads per cell, not the issue slots, bound this kernel: 20 reads/cell in the general
// form, 17 with UG (and 7 instead of 8 registers per row).  Only the Y update keeps a per-row
// multiplier (pYY[k] = 1 on the rows above the read, so their Y stays K/Lh).
//
// FORM 2, UA ("uniform all"): the insertion and deletion qualities are one constant over the reads of
// the launch as well (GATK without the PCR indel model, i.e. PCR-free libraries: 45/45; BASELINE
// config 2; GATK always writes equal insertion and deletion qualities).  Then pMM and pMX = pMY are
// launch constants too, and the product M[r][c] * pMX is needed twice -- by X[r+1][c] in this step
// and by Y[r][c+1] in the next one -- so it is computed once and kept (State::P): SEVEN FMA-pipe
// instructions per cell instead of eight, 13 register operands, 5 registers per row (M, X, Y, P,
// pYY), which lets a lane hold up to 38 rows -- a 150-bp read on FOUR lanes, eight reads per warp,
// half the wavefront skew.  Same values in the same operations as the other forms => same bits.  The rows above the read need no per-row
// zeros here: their prior is 0, so M = 0; X of tile row 0 is forced to 0 by its own register pair
// (pXX[0], pMX[0]) and every X below it is fma(0, cXX, 0 * cMX) = 0; Y keeps pYY[k] = 1.
template <typename T, int G, int R, int FORM>
struct Tile {
  using A = Ar<T>;
  static constexpr bool UG = FORM >= 1, UA = FORM == 2;
  static constexpr int STRIDE = tab_stride_form(R, (int)sizeof(T), FORM);
  static constexpr bool ROT = tab_rotated(R, (int)sizeof(T), FORM);  // rotated table rows, see tab_stride_form
  static constexpr int NV = (R * (int)sizeof(T) + 15) / 16;
  static constexpr int VW = 16 / (int)sizeof(T);
  // steps per loop trip: the unrolled body should stay near 650 instructions (instruction cache)
  static constexpr int UNROLL = (R > PHMM_UNROLL2_ABOVE) ? 2 : kUnrollT;
  static constexpr int RG = UG ? 1 : R;  // rows that keep their own pGM / pXX
  static constexpr int RA = UA ? 1 : R;  // rows that keep their own pMM / pMX / pMY

  T pMM[RA], pMX[RA], pMY[RA];  // UA: pMX[0] only (X-from-M multiplier of tile row 0)
  T pGM[RG], pXX[RG];  // general form: per row.  UG: [0] only (pXX[0] = X multiplier of tile row 0)
  T pYY[UG ? R : 1];   // UG: per-row Y multiplier.  general: [0] = Y multiplier of tile row 0
  T cXX, cGM;          // UG: launch constants
  T cMM, cMX;          // UA: launch constants (pMY == pMX: insertion and deletion quality are equal)
  int npl;             // rows of this lane that lie above the read (always its first rows)
  int off_last;        // ROT: byte offset of this lane's last 16-byte chunk from its (rotated) row base

  // this lane's row base inside a symbol row of the table, and the offset of chunk v from it
  static __device__ __forceinline__ int lane_base(int lane) { return lane * STRIDE + (ROT ? ((lane >> 2) & 1) * 16 : 0); }
  __device__ __forceinline__ int chunk_off(int v) const { return (ROT && v == NV - 1) ? off_last : v * 16; }

  __device__ __forceinline__ T gm(int k) const { return UG ? cGM : pGM[UG ? 0 : k]; }
  __device__ __forceinline__ T xx(int k) const { return UG ? (k == 0 ? pXX[0] : cXX) : pXX[UG ? 0 : k]; }
  __device__ __forceinline__ T yy(int k) const { return UG ? pYY[UG ? k : 0] : (k == 0 ? pYY[0] : pXX[UG ? 0 : k]); }
  __device__ __forceinline__ T mm_(int k) const { return UA ? cMM : pMM[UA ? 0 : k]; }
  __device__ __forceinline__ T mx(int k) const { return UA ? (k == 0 ? pMX[0] : cMX) : pMX[UA ? 0 : k]; }
  __device__ __forceinline__ T my(int k) const { return UA ? cMX : pMY[UA ? 0 : k]; }

  // Fill constants and this lane's slice of the prior table from the staged read.
  // tab_lane = table base + lane_base(lane).  rs points at the group's read blob (planes of Lp bytes; staged in shared memory, or in global
  // memory for the UA form, which touches only bases and base qualities); len == 0 => no read.
  // row0 / npad: tile row 0 of lane 0 is row `row0` of a striped read whose first `npad` rows are
  // boundary replicas (single pass: row0 = 0, npad = G*R - len).
  // layout: the blob's layout flags (phmm_types.h read_layout): which of the insertion / deletion / continuation planes exist; a
  // missing plane's value comes from the 16-byte trailer {ins, del, gcp} (or, for the deletion plane of a same-indel read, from
  // the insertion plane).
  __device__ __forceinline__ void build(const uint8_t* rs, uint32_t len, int lig, const T* lut, const T* __restrict__ mm,
                                        uint8_t* tab_lane, bool with_n, uint32_t layout, int row0 = 0, int npad_override = -1) {
    const uint32_t Lp = round_up16(len);
    // byte offset of the insertion / deletion / continuation quality of read position 0 and the stride per position (0 = constant)
    const uint32_t trailer = read_planes(layout) * Lp;
    const bool has_i = !(layout & kTwoPlaneBit), has_d = !(layout & (kTwoPlaneBit | kSameIndelBit)), has_c = !(layout & kLayoutMask);
    const uint32_t off_i = has_i ? 2u * Lp : trailer, off_d = has_d ? 3u * Lp : (has_i ? 2u * Lp : trailer + 1u);
    const uint32_t off_c = has_c ? 4u * Lp : trailer + 2u;
    const uint32_t str_i = has_i ? 1u : 0u, str_c = has_c ? 1u : 0u;
    const int npad = (npad_override >= 0 ? npad_override : G * R - (int)len) - row0;
    npl = min(max(npad - lig * R, 0), R);
    off_last = ((threadIdx.x >> 2) & 1) ? -16 : (NV - 1) * 16;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int pos = lig * R + k - npad;
      T pm = T(0), px = T(0);
      int rcode = 6;
      T xxk, yyk, gmk;
      if (pos >= 0) {
        const uint32_t b = rs[pos];
        const uint32_t q = rs[Lp + pos] & 127u;
        const T e = lut[q];
        pm = A::sub(T(1), e);
        px = A::div(e, T(3));
        if constexpr (!UA) {
          const uint32_t iq = rs[off_i + str_i * (uint32_t)pos] & 127u, dq = rs[off_d + str_i * (uint32_t)pos] & 127u;
          const uint32_t mn = min(iq, dq), mx = max(iq, dq);
          pMM[k] = mm[((mx * (mx + 1u)) >> 1) + mn];
          pMX[k] = lut[iq];
          pMY[k] = lut[dq];
        } else if (k == 0) {
          pMX[0] = cMX;
        }
        T pc = cXX;
        if constexpr (!UG) pc = lut[rs[off_c + str_c * (uint32_t)pos] & 127u];
        gmk = A::sub(T(1), pc);
        xxk = pc;
        yyk = pc;
        rcode = base_code(b);
      } else {
        // boundary replica: M = 0 (prior 0), X = 0, Y stays K/Lh.  Tile row 0 forces X to 0 whatever
        // the shuffle delivers; below it X_up is already 0.
        if constexpr (!UA) { pMM[k] = T(0); pMX[k] = T(0); pMY[k] = T(0); }
        else if (k == 0) pMX[0] = T(0);
        gmk = T(0);
        xxk = (k == 0) ? T(0) : T(1);
        yyk = T(1);
      }
      if constexpr (UG) {
        pYY[k] = yyk;
        if (k == 0) { pXX[0] = (pos >= 0) ? cXX : T(0); pGM[0] = cGM; }
      } else {
        pGM[k] = gmk;
        pXX[k] = (k == 0) ? xxk : ((pos >= 0) ? xxk : T(1));
        if (k == 0) pYY[0] = yyk;
      }
      const int koff = chunk_off(k / VW) + (k % VW) * (int)sizeof(T);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const bool m = (rcode == kCodeN) || (rcode == h);
        const T v = (pos >= 0) ? (m ? pm : px) : T(0);
        *reinterpret_cast<T*>(tab_lane + h * (32 * STRIDE) + koff) = v;
      }
      *reinterpret_cast<T*>(tab_lane + kCodePad * (32 * STRIDE) + koff) = T(0);
      if (with_n) *reinterpret_cast<T*>(tab_lane + kCodeN * (32 * STRIDE) + koff) = (pos >= 0) ? pm : T(0);
    }
  }

  // Symbol rows beyond N (launches whose haplotypes contain bytes outside ACGTN; rare): raw-byte equality as
  // GKL compares -- the OTHER row matches a read N only, row kCodeExtra0 + e a read N or the byte extra[e].
  // A plain loop kept apart from build() so that the common path pays nothing for it.
  // (static and fed by value: a non-inlined member would take the tile's address and push its arrays to local memory)
  static __device__ __noinline__ void build_other_rows(const uint8_t* rs, uint32_t len, int lig, const T* lut, uint8_t* tab_lane, uint32_t n_sym,
                                                       uint64_t extra, int off_last, int row0 = 0, int npad_override = -1) {
    const uint32_t Lp = round_up16(len);
    const int npad = (npad_override >= 0 ? npad_override : G * R - (int)len) - row0;
    for (int k = 0; k < R; ++k) {
      const int pos = lig * R + k - npad;
      T pm = T(0), px = T(0);
      uint32_t b = 0x100u;
      if (pos >= 0) {
        b = rs[pos];
        const T e = lut[rs[Lp + pos] & 127u];
        pm = A::sub(T(1), e);
        px = A::div(e, T(3));
      }
      const int v = k / VW;
      const int koff = ((ROT && v == NV - 1) ? off_last : v * 16) + (k % VW) * (int)sizeof(T);
      for (uint32_t sym = (uint32_t)kCodeOther; sym < n_sym; ++sym) {
        const bool m = b == (uint32_t)'N' || (sym >= (uint32_t)kCodeExtra0 && b == (uint32_t)((extra >> (8u * (sym - kCodeExtra0))) & 0xffu));
        *reinterpret_cast<T*>(tab_lane + sym * (32 * STRIDE) + koff) = (pos >= 0) ? (m ? pm : px) : T(0);
      }
    }
  }

  // DP state of the tile at the column processed last: R rows of (M, X, Y), the values received
  // from the lane above one step ago (= the diagonal inputs of tile row 0) and the running sum.
  struct State {
    T M[R], X[R], Y[R];
    T P[UA ? R : 1];  // UA: M[k] * cMX of the same column (made for X of the row below, reused for Y of the next column)
    T dM, dX, dY;
    T acc;
  };

  __device__ __forceinline__ void init(State& st, T y_init) const {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      st.M[k] = T(0);
      st.X[k] = T(0);
      st.Y[k] = (k < npl) ? y_init : T(0);
      if constexpr (UA) st.P[k] = T(0);
    }
    st.dM = T(0);
    st.dX = T(0);
    st.dY = __shfl_up_sync(0xffffffffu, st.Y[R - 1], 1, G);
    st.acc = T(0);
  }

  // One column.  prow = this lane's slice of the prior table for the symbol of its column;
  // (uM, uX, uY) = bottom row of the lane above at the same column (it finished it one step ago).
  __device__ __forceinline__ void step(State& st, const uint8_t* prow, T uM, T uX, T uY) const {
    T pr[NV * VW];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if constexpr (sizeof(T) == 4) {
        const float4 f = *reinterpret_cast<const float4*>(prow + chunk_off(v));
        pr[v * 4 + 0] = f.x; pr[v * 4 + 1] = f.y; pr[v * 4 + 2] = f.z; pr[v * 4 + 3] = f.w;
      } else {
        const double2 f = *reinterpret_cast<const double2*>(prow + v * 16);
        pr[v * 2 + 0] = f.x; pr[v * 2 + 1] = f.y;
      }
    }
    T nM[R], nX[R], nY[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const T md = k ? st.M[k - 1] : st.dM;
      const T xd = k ? st.X[k - 1] : st.dX;
      const T yd = k ? st.Y[k - 1] : st.dY;
      T s = A::mul(md, mm_(k));
      s = A::fma(xd, gm(k), s);
      s = A::fma(yd, gm(k), s);
      nM[k] = A::mul(s, pr[k]);
      if constexpr (UA) nY[k] = A::fma(st.Y[k], yy(k), st.P[k]);
      else nY[k] = A::fma(st.Y[k], yy(k), A::mul(st.M[k], my(k)));
    }
    nX[0] = A::fma(uX, xx(0), A::mul(uM, mx(0)));
    if constexpr (UA) {
#pragma unroll
      for (int k = 0; k < R; ++k) st.P[k] = A::mul(nM[k], cMX);
#pragma unroll
      for (int k = 1; k < R; ++k) nX[k] = A::fma(nX[k - 1], xx(k), st.P[k - 1]);
    } else {
#pragma unroll
      for (int k = 1; k < R; ++k) nX[k] = A::fma(nX[k - 1], xx(k), A::mul(nM[k - 1], mx(k)));
    }
    st.acc = A::add(st.acc, A::add(nM[R - 1], nX[R - 1]));
    st.dM = uM; st.dX = uX; st.dY = uY;
#pragma unroll
    for (int k = 0; k < R; ++k) { st.M[k] = nM[k]; st.X[k] = nX[k]; st.Y[k] = nY[k]; }
  }

  // One haplotype.  hs_lane[t] is the symbol offset (in 16-byte units) of the column this
  // lane sees at step t.  Returns sum over columns of (M + X) of this lane's bottom row.
  __device__ __forceinline__ T run(const uint8_t* tab_lane, const uint16_t* hs_lane, int nsteps, T y_init) const {
    State st;
    init(st, y_init);
#pragma unroll(UNROLL)
    for (int t = 0; t < nsteps; ++t) {
      const uint32_t hoff = hs_lane[t];
      const T uM = __shfl_up_sync(0xffffffffu, st.M[R - 1], 1, G);
      const T uX = __shfl_up_sync(0xffffffffu, st.X[R - 1], 1, G);
      const T uY = __shfl_up_sync(0xffffffffu, st.Y[R - 1], 1, G);
      step(st, tab_lane + hoff * 16u, uM, uX, uY);
    }
    return st.acc;
  }
};

// ----------------------------------------------------------------------------------------
// Haplotype-pair form (FP32, uniform gap-continuation quality): every lane runs its R read rows against TWO haplotypes
// of the region at once.  The two cells of a row share the row's transition terms, so the state lives in aligned
// register pairs and the recurrences run as packed f32x2 instructions (FFMA2 / FMUL2, sm_100a) whose per-row operand
// is a broadcast scalar register (R.F32) and whose launch constants come from uniform registers: 7 packed + 2 scalar
// issue slots per TWO cells instead of 16, no register-bank conflicts (a pair reads both banks once).  Each half is
// an IEEE fp32 operation with round-to-nearest and subnormals, in the statement order of Tile::step => same bits.
// Measured in isolation (tools/ubench/step_pairs.cu): 9.5-9.8 cycles per warp-cell against 10.4 for the scalar step;
// the FMA pipe itself (2 cycles per packed instruction) is then the limiter (ncu: 85-87 % busy, math_pipe_throttle).
// Only the prior multiply stays scalar: the two columns see different haplotype symbols, so their priors come from
// two table rows.
template <int G, int R>
struct PairTile {
  using T1 = Tile<float, G, R, 1>;
  static constexpr int NV = T1::NV;
  // steps per loop trip: 2, and 4 for the classes of the standard read lengths (100/101, 150/151 and 250 bp), which carry
  // whole chunks on their own (measured: config 4 +2.3 %, config 1 +1.4 %; unrolling every class by 4 costs the ragged
  // config 3, which keeps a dozen loop bodies in flight, 2.6 %; 1 loses 15 %: the state arrays then rotate through MOVs)
  static constexpr int UNROLL = ((G == 8 && (R == 19 || R == 13)) || (G == 16 && R == 16)) ? 4 : 2;
  static __device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

  struct State {
    float2 M[R], X[R], Y[R];
    float2 dM, dX, dY;
    float2 acc;
  };

  static __device__ __forceinline__ float2 shfl_up2(float2 v) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1, G), __shfl_up_sync(0xffffffffu, v.y, 1, G));
  }

  static __device__ __forceinline__ void init(const T1& t, State& st, float2 y_init) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      st.M[k] = bc(0.f);
      st.X[k] = bc(0.f);
      st.Y[k] = (k < t.npl) ? y_init : bc(0.f);
    }
    st.dM = bc(0.f);
    st.dX = bc(0.f);
    st.dY = shfl_up2(st.Y[R - 1]);
    st.acc = bc(0.f);
  }

  // one column of each haplotype; prow_a / prow_b = this lane's slice of the prior table for the two symbols
  static __device__ __forceinline__ void step(const T1& t, State& st, const uint8_t* prow_a, const uint8_t* prow_b, float2 uM, float2 uX, float2 uY) {
    float pa[NV * 4], pb[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 f = *reinterpret_cast<const float4*>(prow_a + t.chunk_off(v));
      pa[v * 4 + 0] = f.x; pa[v * 4 + 1] = f.y; pa[v * 4 + 2] = f.z; pa[v * 4 + 3] = f.w;
      const float4 g = *reinterpret_cast<const float4*>(prow_b + t.chunk_off(v));
      pb[v * 4 + 0] = g.x; pb[v * 4 + 1] = g.y; pb[v * 4 + 2] = g.z; pb[v * 4 + 3] = g.w;
    }
    float2 nM[R], nX[R], nY[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float2 md = k ? st.M[k - 1] : st.dM;
      const float2 xd = k ? st.X[k - 1] : st.dX;
      const float2 yd = k ? st.Y[k - 1] : st.dY;
      float2 s = __fmul2_rn(md, bc(t.pMM[k]));
      s = __ffma2_rn(xd, bc(t.cGM), s);
      s = __ffma2_rn(yd, bc(t.cGM), s);
      nM[k] = make_float2(__fmul_rn(s.x, pa[k]), __fmul_rn(s.y, pb[k]));
      nY[k] = __ffma2_rn(st.Y[k], bc(t.pYY[k]), __fmul2_rn(st.M[k], bc(t.pMY[k])));
    }
    nX[0] = __ffma2_rn(uX, bc(t.pXX[0]), __fmul2_rn(uM, bc(t.pMX[0])));
#pragma unroll
    for (int k = 1; k < R; ++k) nX[k] = __ffma2_rn(nX[k - 1], bc(t.cXX), __fmul2_rn(nM[k - 1], bc(t.pMX[k])));
    st.acc = __fadd2_rn(st.acc, __fadd2_rn(nM[R - 1], nX[R - 1]));
    st.dM = uM; st.dX = uX; st.dY = uY;
#pragma unroll
    for (int k = 0; k < R; ++k) { st.M[k] = nM[k]; st.X[k] = nX[k]; st.Y[k] = nY[k]; }
  }

  // One haplotype pair.  hs_lane[t] holds the table offsets (16-byte units) of the two columns this lane sees at
  // step t: low half = first haplotype, high half = second.  Returns the two row sums of this lane's bottom row.
  static __device__ __forceinline__ float2 run(const T1& t, const uint8_t* tab_lane, const uint32_t* hs_lane, int nsteps, float2 y_init) {
    State st;
    init(t, st, y_init);
#pragma unroll(UNROLL)
    for (int i = 0; i < nsteps; ++i) {
      const uint32_t h = hs_lane[i];
      const float2 uM = shfl_up2(st.M[R - 1]);
      const float2 uX = shfl_up2(st.X[R - 1]);
      const float2 uY = shfl_up2(st.Y[R - 1]);
      step(t, st, tab_lane + (h & 0xffffu) * 16u, tab_lane + (h >> 16) * 16u, uM, uX, uY);
    }
    return st.acc;
  }
};

// ----------------------------------------------------------------------------------------
// result emission
__device__ __forceinline__ void emit_f32(const KParams& p, const ReadMeta& rm, uint32_t read, uint32_t hap, float S) {
  const uint32_t oi = rm.out_off + p.hmeta[hap].col;  // the caller's haplotype order
  if (p.raw_f32) p.raw_f32[oi] = S;
  if (S < 1e-28f) {  // GKL MIN_ACCEPTED: queue the pair for the double-precision kernel
    const uint32_t cls = rm.len_cls >> 24;
    const uint32_t slot = atomicAdd(&p.rerun_count[cls], 1u);
    RerunEntry e;
    e.read = read;
    e.hap = hap;
    p.rerun[p.rerun_base[cls] + slot] = e;
    p.used_fp64[oi] = 1;
  } else {
    // (float)log10((double)S): the correctly rounded log10f(S); K = 2^120
    const float l = (float)log10((double)S);
    const float k = 0x1.20fd22p+5f;  // (float)log10(2^120)
    p.out[oi] = (double)__fsub_rn(l, k);
    p.used_fp64[oi] = 0;
  }
}
__device__ __forceinline__ void emit_f64(const KParams& p, const ReadMeta& rm, uint32_t hap, double S) {
  const uint32_t oi = rm.out_off + p.hmeta[hap].col;  // the caller's haplotype order
  p.out[oi] = log10(S) - 0x1.330cf3d4eda85p+8;  // log10(2^1020) as glibc rounds it (307.0505955772608)
  p.used_fp64[oi] = 1;
}

// ----------------------------------------------------------------------------------------
// Task form (FP32 main path): one CTA = one task = <= 32/G reads of a region x a run of its
// haplotypes, all lane groups streaming the same haplotypes.
template <typename T, int G, int R, int FORM>
__device__ __forceinline__ void run_task(const KParams& p, const Task task, uint8_t* smem) {
  using L = Layout<T, G, R, false, FORM>;
  constexpr bool UA = FORM == 2;
  using A = Ar<T>;
  const int lane = threadIdx.x;
  const int grp = lane / G, lig = lane % G;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint8_t* tab_lane = smem + L::OFF_TAB + Tile<T, G, R, FORM>::lane_base(lane);
  uint8_t* un = smem + L::off_union(p.n_sym);
  T* lut = reinterpret_cast<T*>(un);
  uint8_t* rstage = un + L::LUT_BYTES + grp * L::RSTAGE;
  uint16_t* hs = reinterpret_cast<uint16_t*>(un);  // aliases lut + rstage: written only after tile.build
  uint8_t* hstage = un + L::union_bytes(p.hs_cap);
  const T* __restrict__ mm = reinterpret_cast<const T*>(p.mm);

  for (int i = lane; i < 128; i += 32) lut[i] = reinterpret_cast<const T*>(p.ph2pr)[i];
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  Tile<T, G, R, FORM> tile;
  if constexpr (sizeof(T) == 4) { tile.cXX = p.c_xx_f; tile.cGM = p.c_gm_f; tile.cMM = p.c_mm_f; tile.cMX = p.c_mx_f; }
  else { tile.cXX = p.c_xx_d; tile.cGM = p.c_gm_d; tile.cMM = T(0); tile.cMX = T(0); }

  const bool active = grp < (int)task.n_reads;
  const uint32_t read = task.read0 + (active ? grp : 0);
  const ReadMeta rm = p.rmeta[read];
  const uint32_t rlen = active ? read_len_of(rm) : 0u;
  const uint32_t layout = read_layout(rm);
  const HapMeta h_first = p.hmeta[task.hap0];
  const HapMeta h_last = p.hmeta[task.hap0 + task.n_haps - 1];
  const uint32_t hap_bytes = (h_last.data_off16 - h_first.data_off16) * 16u + round_up16(h_last.len);
  // ---- stage reads + haplotypes with TMA bulk copies
  const uint32_t my_bytes = ((!UA && active && lig == 0) ? read_blob_bytes(rlen, layout) : 0u) + (lane == 0 ? hap_bytes : 0u);
  const uint32_t tot = __reduce_add_sync(0xffffffffu, my_bytes);
  if (lane == 0) mbar_expect_tx(bar, tot);
  __syncwarp();
  if (!UA && active && lig == 0) bulk_g2s(rstage, p.reads + (size_t)rm.data_off16 * 16u, read_blob_bytes(rlen, layout), bar);
  if (lane == 0) bulk_g2s(hstage, p.haps + (size_t)h_first.data_off16 * 16u, hap_bytes, bar);
  // ---- per-row constants + prior table (the UA form builds from global memory while the copy flies)
  if constexpr (UA) tile.build(p.reads + (size_t)rm.data_off16 * 16u, rlen, lig, lut, mm, tab_lane, p.n_sym > 5u, layout);
  mbar_wait(bar, 0u);
  if constexpr (!UA) tile.build(rstage, rlen, lig, lut, mm, tab_lane, p.n_sym > 5u, layout);
  if (p.n_sym > (uint32_t)kCodeOther)
    Tile<T, G, R, FORM>::build_other_rows(UA ? p.reads + (size_t)rm.data_off16 * 16u : rstage, rlen, lig, lut, tab_lane, p.n_sym, p.extra_bytes, tile.off_last);
  __syncwarp();  // every lane is done with the LUT and the read staging before the stream overwrites them
  // ---- haplotype stream: [G-1 PAD] hap0 [G-1 PAD] hap1 ... [G-1 PAD]
  uint32_t off = 0;
  for (uint32_t j = 0; j < task.n_haps; ++j) {
    const HapMeta hm = p.hmeta[task.hap0 + j];
    const uint8_t* src = hstage + (hm.data_off16 - h_first.data_off16) * 16u;
    if (lane < G - 1) hs[off + lane] = (uint16_t)(kCodePad * L::HSCALE);
    for (uint32_t x = lane; x < hm.len; x += 32) {
      const int c = hap_code(src[x], p.n_sym, p.extra_bytes);
      hs[off + (G - 1) + x] = (uint16_t)(c * L::HSCALE);
    }
    off += (G - 1) + hm.len;
  }
  if (lane < G - 1) hs[off + lane] = (uint16_t)(kCodePad * L::HSCALE);
  __syncwarp();
  // ---- wavefront over every haplotype of the task
  off = 0;
  for (uint32_t j = 0; j < task.n_haps; ++j) {
    const uint32_t Lh = p.hmeta[task.hap0 + j].len;
    const T y_init = A::div(A::K(), (T)(int)Lh);
    const T acc = tile.run(tab_lane, hs + off + (G - 1) - lig, (int)Lh + G - 1, y_init);
    off += (G - 1) + Lh;
    if (active && lig == G - 1) {
      if constexpr (sizeof(T) == 4) emit_f32(p, rm, read, task.hap0 + j, acc);
      else emit_f64(p, rm, task.hap0 + j, acc);
    }
  }
}

// Task form of the haplotype-pair kernels: the task's haplotypes are taken two at a time (an odd run ends with a
// pair whose second stream is all PAD: its result is dropped).  Stream entries are 32 bits: two table offsets.
template <int G, int R>
__device__ __forceinline__ void run_task_pairs(const KParams& p, const Task task, uint8_t* smem) {
  using L = Layout<float, G, R, false, 1>;
  using TT = Tile<float, G, R, 1>;
  using PT = PairTile<G, R>;
  const int lane = threadIdx.x;
  const int grp = lane / G, lig = lane % G;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint8_t* tab_lane = smem + L::OFF_TAB + TT::lane_base(lane);
  uint8_t* un = smem + L::off_union(p.n_sym);
  float* lut = reinterpret_cast<float*>(un);
  uint8_t* rstage = un + L::LUT_BYTES + grp * L::RSTAGE;
  uint32_t* hs = reinterpret_cast<uint32_t*>(un);  // aliases lut + rstage: written only after tile.build
  uint8_t* hstage = un + L::union_bytes(p.hs_cap);
  const float* __restrict__ mm = reinterpret_cast<const float*>(p.mm);

  for (int i = lane; i < 128; i += 32) lut[i] = reinterpret_cast<const float*>(p.ph2pr)[i];
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  TT tile;
  tile.cXX = p.c_xx_f; tile.cGM = p.c_gm_f; tile.cMM = p.c_mm_f; tile.cMX = p.c_mx_f;

  const bool active = grp < (int)task.n_reads;
  const uint32_t read = task.read0 + (active ? grp : 0);
  const ReadMeta rm = p.rmeta[read];
  const uint32_t rlen = active ? read_len_of(rm) : 0u;
  const uint32_t layout = read_layout(rm);
  const HapMeta h_first = p.hmeta[task.hap0];
  const HapMeta h_last = p.hmeta[task.hap0 + task.n_haps - 1];
  const uint32_t hap_bytes = (h_last.data_off16 - h_first.data_off16) * 16u + round_up16(h_last.len);
  const uint32_t my_bytes = ((active && lig == 0) ? read_blob_bytes(rlen, layout) : 0u) + (lane == 0 ? hap_bytes : 0u);
  const uint32_t tot = __reduce_add_sync(0xffffffffu, my_bytes);
  if (lane == 0) mbar_expect_tx(bar, tot);
  __syncwarp();
  if (active && lig == 0) bulk_g2s(rstage, p.reads + (size_t)rm.data_off16 * 16u, read_blob_bytes(rlen, layout), bar);
  if (lane == 0) bulk_g2s(hstage, p.haps + (size_t)h_first.data_off16 * 16u, hap_bytes, bar);
  mbar_wait(bar, 0u);
  tile.build(rstage, rlen, lig, lut, mm, tab_lane, p.n_sym > 5u, layout);
  if (p.n_sym > (uint32_t)kCodeOther) TT::build_other_rows(rstage, rlen, lig, lut, tab_lane, p.n_sym, p.extra_bytes, tile.off_last);
  __syncwarp();
  // ---- pair stream: [G-1 PAD] pair0 [G-1 PAD] pair1 ... [G-1 PAD]; a pair is max(len_a, len_b) entries long
  constexpr uint32_t kPad = (uint32_t)(kCodePad * L::HSCALE);
  constexpr uint32_t kPad2 = kPad | (kPad << 16);
  const uint32_t n_pairs = (task.n_haps + 1u) / 2u;
  uint32_t off = 0;
  for (uint32_t q = 0; q < n_pairs; ++q) {
    const HapMeta ha = p.hmeta[task.hap0 + 2u * q];
    const bool has_b = 2u * q + 1u < task.n_haps;
    const HapMeta hb = has_b ? p.hmeta[task.hap0 + 2u * q + 1u] : ha;
    const uint32_t len_b = has_b ? hb.len : 0u;
    const uint32_t lmax = max(ha.len, len_b);
    const uint8_t* sa = hstage + (ha.data_off16 - h_first.data_off16) * 16u;
    const uint8_t* sb = hstage + (hb.data_off16 - h_first.data_off16) * 16u;
    if (lane < G - 1) hs[off + lane] = kPad2;
    for (uint32_t x = lane; x < lmax; x += 32) {
      const uint32_t ca = x < ha.len ? (uint32_t)(hap_code(sa[x], p.n_sym, p.extra_bytes) * L::HSCALE) : kPad;
      const uint32_t cb = x < len_b ? (uint32_t)(hap_code(sb[x], p.n_sym, p.extra_bytes) * L::HSCALE) : kPad;
      hs[off + (G - 1) + x] = ca | (cb << 16);
    }
    off += (G - 1) + lmax;
  }
  if (lane < G - 1) hs[off + lane] = kPad2;
  __syncwarp();
  // ---- wavefront over every haplotype pair of the task
  off = 0;
  for (uint32_t q = 0; q < n_pairs; ++q) {
    const uint32_t la = p.hmeta[task.hap0 + 2u * q].len;
    const bool has_b = 2u * q + 1u < task.n_haps;
    const uint32_t lb = has_b ? p.hmeta[task.hap0 + 2u * q + 1u].len : 0u;
    const uint32_t lmax = max(la, lb);
    const float2 y_init = make_float2(__fdiv_rn(0x1p120f, (float)(int)la), has_b ? __fdiv_rn(0x1p120f, (float)(int)lb) : 0.f);
    const float2 acc = PT::run(tile, tab_lane, hs + off + (G - 1) - lig, (int)lmax + G - 1, y_init);
    off += (G - 1) + lmax;
    if (active && lig == G - 1) {
      emit_f32(p, rm, read, task.hap0 + 2u * q, acc.x);
      if (has_b) emit_f32(p, rm, read, task.hap0 + 2u * q + 1u, acc.y);
    }
  }
}

// Queue form (FP64 rerun): CTA `cta` of `nctas` grid-strides over queue `qid`, one pair per lane group.
template <typename T, int G, int R, int FORM>
__device__ __forceinline__ void run_queue(const KParams& p, uint32_t qid, uint32_t cta, uint32_t nctas, uint8_t* smem) {
  static_assert(FORM != 2, "the all-uniform form has no queue kernel");
  using L = Layout<T, G, R, true, FORM>;
  using A = Ar<T>;
  constexpr int NG = L::NG;
  const int lane = threadIdx.x;
  const int grp = lane / G, lig = lane % G;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint8_t* tab_lane = smem + L::OFF_TAB + Tile<T, G, R, FORM>::lane_base(lane);
  uint8_t* un = smem + L::off_union(p.n_sym);
  T* lut = reinterpret_cast<T*>(un);
  uint8_t* rstage = un + L::LUT_BYTES + grp * L::RSTAGE;
  const uint32_t hs_bytes = round_up16(p.hs_cap * 2u);
  const uint32_t hstage_bytes = round_up16(p.hap_stage_bytes);
  uint16_t* hs = reinterpret_cast<uint16_t*>(un) + grp * (hs_bytes / 2u);  // aliases lut + rstage
  uint8_t* hstage = un + L::union_bytes(p.hs_cap) + grp * hstage_bytes;
  const T* __restrict__ mm = reinterpret_cast<const T*>(p.mm);

  const uint32_t count = p.rerun_count[qid];
  if (cta * NG >= count) return;
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t parity = 0;
  Tile<T, G, R, FORM> tile;
  if constexpr (sizeof(T) == 4) { tile.cXX = p.c_xx_f; tile.cGM = p.c_gm_f; }
  else { tile.cXX = p.c_xx_d; tile.cGM = p.c_gm_d; }
  tile.cMM = T(0); tile.cMX = T(0);
  const RerunEntry* list = p.rerun + p.rerun_base[qid];
  for (uint32_t base = cta * NG; base < count; base += nctas * NG) {
    const bool active = base + grp < count;
    RerunEntry e;
    e.read = 0; e.hap = 0;
    if (active) e = list[base + grp];
    const ReadMeta rm = p.rmeta[e.read];
    const HapMeta hm = p.hmeta[e.hap];
    const uint32_t rlen = active ? read_len_of(rm) : 0u;
    const uint32_t layout = read_layout(rm);
    const uint32_t Lh = active ? hm.len : 0u;
    // the LUT shares its shared memory with the haplotype stream of the previous round: reload it
    for (int i = lane; i < 128; i += 32) lut[i] = reinterpret_cast<const T*>(p.ph2pr)[i];
    fence_proxy_async();  // staging / stream were touched through the generic proxy last round
    const uint32_t my_bytes = (active && lig == 0) ? read_blob_bytes(rlen, layout) + round_up16(Lh) : 0u;
    const uint32_t tot = __reduce_add_sync(0xffffffffu, my_bytes);
    if (lane == 0) mbar_expect_tx(bar, tot);
    __syncwarp();
    if (active && lig == 0) {
      bulk_g2s(rstage, p.reads + (size_t)rm.data_off16 * 16u, read_blob_bytes(rlen, layout), bar);
      bulk_g2s(hstage, p.haps + (size_t)hm.data_off16 * 16u, round_up16(Lh), bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    tile.build(rstage, rlen, lig, lut, mm, tab_lane, p.n_sym > 5u, layout);
    if (p.n_sym > (uint32_t)kCodeOther) Tile<T, G, R, FORM>::build_other_rows(rstage, rlen, lig, lut, tab_lane, p.n_sym, p.extra_bytes, tile.off_last);
    __syncwarp();  // done with the LUT and the read staging before the stream overwrites them
    const uint32_t Lmax = __reduce_max_sync(0xffffffffu, Lh);
    const uint32_t total = Lmax + 2u * (G - 1);
    for (uint32_t x = lig; x < total; x += G) {
      int c = kCodePad;
      if (x >= (uint32_t)(G - 1) && x < (uint32_t)(G - 1) + Lh) {
        c = hap_code(hstage[x - (G - 1)], p.n_sym, p.extra_bytes);
      }
      hs[x] = (uint16_t)(c * L::HSCALE);
    }
    __syncwarp();
    const T y_init = A::div(A::K(), (T)(int)(Lh ? Lh : 1u));
    const T acc = tile.run(tab_lane, hs + (G - 1) - lig, (int)Lmax + G - 1, y_init);
    if (active && lig == G - 1) {
      if constexpr (sizeof(T) == 4) emit_f32(p, rm, e.read, e.hap, acc);
      else emit_f64(p, rm, e.hap, acc);
    }
    __syncwarp();
  }
}

}  // namespace fcsphmm
