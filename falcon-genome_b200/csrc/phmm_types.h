// phmm_types.h — device-visible data layout shared by the packer (host) and the kernels.
//
// HBM layout of one packed chunk (every section 16-byte aligned so that one
// cp.async.bulk moves a read set or a haplotype set into shared memory):
//
//   reads   : per read  [bases | base_q | ins_q | del_q | gcp], each plane padded to
//             Lp = round_up(len, 16) bytes  -> 5*Lp bytes per read, reads of a task adjacent;
//             a read whose insertion, deletion and continuation qualities are each one constant
//             (what GATK passes without BAM BI/BD tags and without the PCR indel model) is packed
//             as TWO planes plus a 16-byte trailer [ins, del, gcp, 0...]: 2*Lp + 16 bytes
//   haps    : per hap   bases padded to round_up(len, 16); the haplotypes of a region are packed longest first (pairs of
//             neighbours then differ least in length, an odd one out is the shortest), HapMeta::col keeps the caller's index
//   rmeta[] : ReadMeta per read,  hmeta[] : HapMeta per hap,  tasks[] : Task per CTA
//   out[]   : double per pair, used[] : u8 per pair, raw[] : float per pair (optional)
//   rerun[] : (read, hap) pairs queued for the double-precision kernel, one segment per
//             FP64 kernel class, filled by the FP32 kernel with atomics
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHMM_HD __host__ __device__
#else
#define PHMM_HD
#endif

namespace fcsphmm {

constexpr int kTabRows = 6;  // haplotype symbol classes: A C G T PAD N  (N last: its table row is
constexpr int kCodePad = 4;  //   allocated only when some haplotype of the launch contains an N)
constexpr int kCodeN = 5;
// Haplotype bytes outside ACGTN (IUPAC codes in b37/hg19 derived haplotypes, lower case, ...).  GKL compares
// raw bytes, so such a byte matches a read base only if the read holds the very same byte or an N.  Bytes
// that no read of the chunk contains share the OTHER row (matches a read N only); the few that do occur in
// reads as well get a row of their own (kCodeExtra0 + e, byte values in KParams::extra_bytes).
constexpr int kCodeOther = 6;
constexpr int kCodeExtra0 = 7;
constexpr int kMaxExtraSyms = 8;
constexpr int kMaxSyms = kCodeExtra0 + kMaxExtraSyms;
constexpr int kMaxF64Classes = 40;
constexpr int kQueueGenericF64 = kMaxF64Classes - 1;  // FP64 rerun queue of the striped generic path

struct ReadMeta {
  uint32_t data_off16;  // offset of the read blob in `reads`, in 16-byte units
  uint32_t len_cls;     // len | layout flags (bits 21-23, see read_layout) | (fp64 class id << 24)
  uint32_t out_off;     // index in out[] of (this read, first hap of its region)
  uint32_t hap0;        // first hap (chunk-wide index) of its region
};

struct HapMeta {
  uint32_t data_off16;
  uint32_t len;
  uint32_t col;  // the haplotype's index in its region in the caller's order: the output column (haplotypes are packed longest first)
};

// One CTA (one warp) of the FP32 wavefront kernel: up to 32/G reads of one region
// against a run of that region's haplotypes.
struct Task {
  uint32_t read0;
  uint32_t hap0;
  uint16_t n_reads;
  uint16_t n_haps;
  uint32_t cls;  // index of the task's (G, R) class inside its register tier (phmm_tiers.h)
};

struct RerunEntry {
  uint32_t read;
  uint32_t hap;
};

struct KParams {
  const uint8_t* reads;
  const uint8_t* haps;
  const ReadMeta* rmeta;
  const HapMeta* hmeta;
  const Task* tasks;       // FP32 main kernels: tasks of this class
  uint32_t n_tasks;
  const void* ph2pr;       // T[128]
  const void* mm;          // T[8256], index ((max*(max+1))>>1)+min
  double* out;
  uint8_t* used_fp64;
  float* raw_f32;          // may be null
  RerunEntry* rerun;       // base of all segments
  uint32_t* rerun_count;   // [kMaxF64Classes]
  const uint32_t* rerun_base;  // [kMaxF64Classes] segment start (entries)
  // FP64 launches: the classes of one register tier share a launch; segment k drains queue seg_qid[k]
  // with tier-local class seg_cls[k] on CTAs [seg_cta0[k], seg_cta0[k+1])
  uint32_t n_seg;
  uint16_t seg_cls[32];
  uint16_t seg_qid[32];
  uint32_t seg_cta0[33];
  // a segment runs only if seg_min[k] <= queue length <= seg_max[k]: a short queue is drained by the
  // widest class (shortest serial chain per pair), a long one by the throughput class
  uint32_t seg_min[32];
  uint32_t seg_max[32];
  uint32_t hs_cap;         // u16 entries of haplotype stream in shared memory
  uint32_t hap_stage_bytes;  // bytes of raw haplotype staging in shared memory
  uint32_t n_sym;            // prior-table symbol rows in shared memory: 5 (no N in any haplotype), 6 (N), 7 + e (non-ACGTN bytes)
  uint64_t extra_bytes;      // byte e = value of the haplotype byte that owns symbol row kCodeExtra0 + e
  // generic (striped) path: host-built pair list for the FP32 pass, per-CTA boundary scratch rows
  const RerunEntry* gen_list;
  uint32_t gen_count;
  uint32_t scratch_cols;  // columns per scratch plane (3 planes of T per CTA)
  void* scratch;
  // uniform-GCP launches: ph2pr[gcp] and 1 - ph2pr[gcp], read from the constant bank
  float c_xx_f, c_gm_f;
  double c_xx_d, c_gm_d;
  // all-uniform launches (FP32 only; insertion quality == deletion quality): matchToMatch[q, q], ph2pr[q]
  float c_mm_f, c_mx_f;
};

PHMM_HD inline constexpr uint32_t round_up16(uint32_t x) { return (x + 15u) & ~15u; }
// Read blob layouts (flag bits of ReadMeta::len_cls; the packer decides per read, every kernel form reads all of them):
//   none of the bits : five planes [bases | base_q | ins_q | del_q | gcp]
//   kNoGcpPlaneBit   : the gap-continuation quality is one constant (what GATK always passes): no gcp plane,
//                      a 16-byte trailer {ins, del, gcp} follows the planes (only its gcp byte is used)
//   kSameIndelBit    : (with kNoGcpPlaneBit) the deletion qualities equal the insertion qualities position by position
//                      (GATK writes them from one PCR-model value): no del plane either -> [bases | base_q | ins_q | trailer]
//   kTwoPlaneBit     : insertion, deletion and continuation quality are each one constant -> [bases | base_q | trailer]
constexpr uint32_t kTwoPlaneBit = 1u << 23;
constexpr uint32_t kNoGcpPlaneBit = 1u << 22;
constexpr uint32_t kSameIndelBit = 1u << 21;
constexpr uint32_t kLayoutMask = kTwoPlaneBit | kNoGcpPlaneBit | kSameIndelBit;
PHMM_HD inline uint32_t read_len_of(const ReadMeta& m) { return m.len_cls & 0x1fffffu; }
PHMM_HD inline uint32_t read_layout(const ReadMeta& m) { return m.len_cls & kLayoutMask; }
PHMM_HD inline bool read_two_plane(const ReadMeta& m) { return (m.len_cls & kTwoPlaneBit) != 0u; }
// planes a layout stores (bases and base qualities always)
PHMM_HD inline constexpr uint32_t read_planes(uint32_t layout) {
  return (layout & kTwoPlaneBit) ? 2u : ((layout & kNoGcpPlaneBit) ? ((layout & kSameIndelBit) ? 3u : 4u) : 5u);
}
PHMM_HD inline constexpr uint32_t read_blob_bytes(uint32_t len, uint32_t layout) {
  return read_planes(layout) * round_up16(len) + ((layout & kLayoutMask) ? 16u : 0u);
}

// Table lane stride in bytes: smallest odd multiple of 16 that holds R values of size esz.
// Odd => the eight lanes of a quarter-warp LDS.128 phase hit eight distinct 16-byte bank groups.
PHMM_HD inline constexpr int tab_stride_bytes(int R, int esz) {
  int n16 = (R * esz + 15) / 16;
  if ((n16 & 1) == 0) n16 += 1;
  return n16 * 16;
}

// Kernel forms: 0 = general, 1 = uniform gap-continuation quality, 2 = all transition qualities uniform.
// The FP32 all-uniform form holds up to 38 rows per lane, so its prior table (5 symbols x 32 lanes x
// stride) decides how many CTAs share an SM; padding the stride to an odd multiple of 16 B would cost
// one CTA per SM for R = 37..40.  When the stride is 2 (mod 4) sixteen-byte chunks it stays unpadded
// and the table is ROTATED instead: lanes 4..7 of every quarter-warp keep their chunks one position
// further (the last chunk wraps to the front), which makes the eight lanes of an LDS.128 phase hit
// eight distinct bank groups again.
PHMM_HD inline constexpr bool tab_rotated(int R, int esz, int form) {
  return form == 2 && esz == 4 && (((R * esz + 15) / 16) % 4) == 2;
}
PHMM_HD inline constexpr int tab_stride_form(int R, int esz, int form) {
  return tab_rotated(R, esz, form) ? ((R * esz + 15) / 16) * 16 : tab_stride_bytes(R, esz);
}

}  // namespace fcsphmm
