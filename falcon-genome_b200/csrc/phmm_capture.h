// phmm_capture.h — on-disk capture / replay of PairHMM testcases at the C-ABI boundary
// (SURVEY.md §8(f) row f4).  A handle with capture enabled appends every region it is asked to score,
// so a GATK run elsewhere (through the JNI shim) can dump its real testcases for replay here.
//
// Format (little endian):  "FCSPHMM1", then blocks:
//   u32 'RBLK', u32 n_regions, per region: u32 n_reads, u32 n_haps,
//     per read: u32 len, bases[len], base_q[len], ins_q[len], del_q[len], gcp[len]
//     per hap:  u32 len, bases[len]
#pragma once
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "phmm_engine.h"

namespace fcsphmm {

class CaptureWriter {
 public:
  ~CaptureWriter() { close(); }
  int open(const std::string& path);
  void close();
  bool active() const { return f_ != nullptr; }
  int append(const Input& in);

 private:
  std::FILE* f_ = nullptr;
  std::mutex mu_;
};

// A loaded capture: owns the planes a fcs_phmm_flat_batch points into.
struct LoadedCapture {
  std::vector<uint8_t> rb, rq, ri, rd, rc, hb;
  std::vector<int64_t> rd_off, hp_off, reg_out0;
  std::vector<int32_t> rd_len, hp_len, reg_read0, reg_nreads, reg_hap0, reg_nhaps;
  void view(fcs_phmm_flat_batch* out) const;
};
int load_capture(const std::string& path, LoadedCapture** out);

// One 'RBLK' block in memory (the broker protocol sends exactly this).
void serialize_block(const Input& in, std::vector<uint8_t>& out);
// Parses a sequence of blocks; appends to `c`.  Returns false on truncation / corruption.
bool parse_blocks(const uint8_t* p, size_t n, LoadedCapture& c);

}  // namespace fcsphmm
