// phmm_capture.cpp — see phmm_capture.h.
#include "phmm_capture.h"

#include <cstring>

namespace fcsphmm {

static const char kMagic[8] = {'F', 'C', 'S', 'P', 'H', 'M', 'M', '1'};
static const uint32_t kBlockTag = 0x4B4C4252u;  // "RBLK"

int CaptureWriter::open(const std::string& path) {
  std::lock_guard<std::mutex> lk(mu_);
  if (f_) std::fclose(f_);
  f_ = nullptr;
  std::FILE* probe = std::fopen(path.c_str(), "rb");
  bool fresh = true;
  if (probe) {
    char m[8];
    fresh = std::fread(m, 1, 8, probe) != 8;
    if (!fresh && std::memcmp(m, kMagic, 8) != 0) {
      std::fclose(probe);
      return set_error(FCS_PHMM_EINVAL, "capture file exists and is not a FCSPHMM1 file: " + path);
    }
    std::fclose(probe);
  }
  f_ = std::fopen(path.c_str(), fresh ? "wb" : "ab");
  if (!f_) return set_error(FCS_PHMM_EINVAL, "cannot open capture file " + path);
  if (fresh) std::fwrite(kMagic, 1, 8, f_);
  return FCS_PHMM_OK;
}

void CaptureWriter::close() {
  std::lock_guard<std::mutex> lk(mu_);
  if (f_) std::fclose(f_);
  f_ = nullptr;
}

void serialize_block(const Input& in, std::vector<uint8_t>& out) {
  auto w32 = [&](uint32_t v) {
    const uint8_t* q = reinterpret_cast<const uint8_t*>(&v);
    out.insert(out.end(), q, q + 4);
  };
  const int64_t n = in.n_regions();
  w32(kBlockTag);
  w32((uint32_t)n);
  for (int64_t g = 0; g < n; ++g) {
    int32_t nr = 0, nh = 0;
    in.shape(g, nr, nh);
    w32((uint32_t)std::max(0, nr));
    w32((uint32_t)std::max(0, nh));
    for (int32_t i = 0; i < nr; ++i) {
      const InRead r = in.read(g, i);
      const uint32_t len = (uint32_t)std::max(0, r.len);
      w32(len);
      const uint8_t* pl[5] = {r.b, r.q, r.i, r.d, r.c};
      for (int k = 0; k < 5; ++k) out.insert(out.end(), pl[k], pl[k] + len);
    }
    for (int32_t j = 0; j < nh; ++j) {
      const InHap h = in.hap(g, j);
      const uint32_t len = (uint32_t)std::max(0, h.len);
      w32(len);
      out.insert(out.end(), h.b, h.b + len);
    }
  }
}

int CaptureWriter::append(const Input& in) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!f_) return FCS_PHMM_OK;
  std::vector<uint8_t> buf;
  serialize_block(in, buf);
  std::fwrite(buf.data(), 1, buf.size(), f_);
  std::fflush(f_);
  return std::ferror(f_) ? set_error(FCS_PHMM_EINVAL, "write to capture file failed") : FCS_PHMM_OK;
}

void LoadedCapture::view(fcs_phmm_flat_batch* o) const {
  o->read_bases = rb.data(); o->read_q = rq.data(); o->read_i = ri.data(); o->read_d = rd.data(); o->read_c = rc.data();
  o->rd_off = rd_off.data(); o->rd_len = rd_len.data(); o->n_reads = (int64_t)rd_len.size();
  o->hap_bases = hb.data(); o->hp_off = hp_off.data(); o->hp_len = hp_len.data(); o->n_haps = (int64_t)hp_len.size();
  o->reg_read0 = reg_read0.data(); o->reg_nreads = reg_nreads.data(); o->reg_hap0 = reg_hap0.data();
  o->reg_nhaps = reg_nhaps.data(); o->reg_out0 = reg_out0.data(); o->n_regions = (int64_t)reg_read0.size();
}

bool parse_blocks(const uint8_t* p, size_t n, LoadedCapture& c) {
  size_t pos = 0;
  auto r32 = [&](uint32_t& v) {
    if (pos + 4 > n) return false;
    std::memcpy(&v, p + pos, 4);
    pos += 4;
    return true;
  };
  auto rbytes = [&](std::vector<uint8_t>& dst, uint32_t len) {
    if (pos + len > n) return false;
    dst.insert(dst.end(), p + pos, p + pos + len);
    pos += len;
    return true;
  };
  int64_t out0 = c.reg_out0.empty() ? 0 : c.reg_out0.back() + (int64_t)c.reg_nreads.back() * (int64_t)c.reg_nhaps.back();
  while (pos < n) {
    uint32_t tag = 0, nreg = 0;
    if (!r32(tag) || tag != kBlockTag || !r32(nreg)) return false;
    for (uint32_t g = 0; g < nreg; ++g) {
      uint32_t nr = 0, nh = 0;
      if (!r32(nr) || !r32(nh)) return false;
      c.reg_read0.push_back((int32_t)c.rd_len.size());
      c.reg_nreads.push_back((int32_t)nr);
      c.reg_hap0.push_back((int32_t)c.hp_len.size());
      c.reg_nhaps.push_back((int32_t)nh);
      c.reg_out0.push_back(out0);
      out0 += (int64_t)nr * (int64_t)nh;
      for (uint32_t i = 0; i < nr; ++i) {
        uint32_t len = 0;
        if (!r32(len) || len > (1u << 24)) return false;
        c.rd_off.push_back((int64_t)c.rb.size());
        c.rd_len.push_back((int32_t)len);
        if (!(rbytes(c.rb, len) && rbytes(c.rq, len) && rbytes(c.ri, len) && rbytes(c.rd, len) && rbytes(c.rc, len))) return false;
      }
      for (uint32_t j = 0; j < nh; ++j) {
        uint32_t len = 0;
        if (!r32(len) || len > (1u << 24)) return false;
        c.hp_off.push_back((int64_t)c.hb.size());
        c.hp_len.push_back((int32_t)len);
        if (!rbytes(c.hb, len)) return false;
      }
    }
  }
  return true;
}

int load_capture(const std::string& path, LoadedCapture** out) {
  *out = nullptr;
  std::FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return set_error(FCS_PHMM_EINVAL, "cannot open capture file " + path);
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t got;
  while ((got = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
  std::fclose(f);
  if (buf.size() < 8 || std::memcmp(buf.data(), kMagic, 8) != 0) return set_error(FCS_PHMM_EINVAL, "not a FCSPHMM1 capture file: " + path);
  std::unique_ptr<LoadedCapture> c(new LoadedCapture());
  if (!parse_blocks(buf.data() + 8, buf.size() - 8, *c)) return set_error(FCS_PHMM_EINVAL, "truncated or corrupt capture file: " + path);
  *out = c.release();
  return FCS_PHMM_OK;
}

}  // namespace fcsphmm
