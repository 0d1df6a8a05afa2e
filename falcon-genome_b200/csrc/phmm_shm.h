// Shared-memory transport between libfcs_pairhmm_client and the fcs-pairhmm-nam daemon (SURVEY.md §8(f) f3:
// one process owns the GPUs and serves the <= 32 JVMs of a stage "over shared memory").
//
// The client owns one memfd segment per connection and passes its descriptor to the daemon over the Unix
// socket (SCM_RIGHTS).  A request is the flat batch written straight into the segment in the layout below;
// the socket carries only a 16-byte doorbell each way, and the daemon writes the results into the segment.
// Compared with the byte-stream protocol (serialize -> socket write -> socket read -> parse into planes) the
// input is copied once on the client and not at all in the daemon before packing.
//
//   doorbells (little endian):
//     client -> daemon  u32 'PHSM', u64 segment bytes          + the memfd as ancillary data: (re)attach
//     client -> daemon  u32 'PHSQ', u64 0                      : run the batch described at offset 0
//     daemon -> client  u32 'PHRS', i32 rc, u64 n              : rc == 0: n pairs, results are in the segment;
//                                                                rc <  0: n bytes of error text follow
//
// The daemon never trusts the segment: every section is bounds-checked against the mapped size, and the index
// arrays (offsets, lengths, region tables) are copied to private memory before they are validated and used, so
// a client that rewrites them mid-call cannot steer the daemon outside the mapping.
#pragma once
#include <cstdint>

namespace fcsphmm {

constexpr uint32_t kShmAttach = 0x4D534850u;   // 'PHSM'
constexpr uint32_t kShmRequest = 0x51534850u;  // 'PHSQ'
constexpr uint32_t kShmMagic = 0x4D485346u;    // 'FSHM'
constexpr uint32_t kShmVersion = 1;

// At offset 0 of the segment.  All offsets are bytes from the start of the segment, 64-byte aligned.
struct ShmHeader {
  uint32_t magic, version;
  int64_t n_regions, n_reads, n_haps;
  uint64_t n_pairs;
  uint64_t read_bytes;  // bytes in each of the five read planes
  uint64_t hap_bytes;
  uint64_t off_read_bases, off_read_q, off_read_i, off_read_d, off_read_c;  // read_bytes each
  uint64_t off_rd_off, off_rd_len;                                          // int64[n_reads], int32[n_reads]
  uint64_t off_hap_bases;                                                   // hap_bytes
  uint64_t off_hp_off, off_hp_len;                                          // int64[n_haps], int32[n_haps]
  uint64_t off_reg_read0, off_reg_nreads, off_reg_hap0, off_reg_nhaps;      // int32[n_regions] each
  uint64_t off_out, off_used;                                               // double[n_pairs], uint8[n_pairs] (written by the daemon)
  uint64_t total_bytes;
};

static_assert(sizeof(ShmHeader) == 192, "ShmHeader is part of the client/daemon protocol");

inline uint64_t shm_align(uint64_t x) { return (x + 63u) & ~uint64_t(63); }

}  // namespace fcsphmm
