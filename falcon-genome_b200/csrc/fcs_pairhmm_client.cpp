// libfcs_pairhmm_client — CUDA-free client of the fcs-pairhmm-nam daemon (protocol in fcs_pairhmm_nam.cpp).
// What a JVM-side shim links when the GPUs are owned by the daemon instead of by the JVM itself.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/fcs_pairhmm.h"

struct fcs_phmm_remote {
  int fd;
  std::string err;
};
static thread_local std::string g_cerr;

static bool rd(int fd, void* p, size_t n) {
  uint8_t* b = static_cast<uint8_t*>(p);
  while (n) {
    ssize_t r = ::read(fd, b, n);
    if (r <= 0) {
      if (r < 0 && errno == EINTR) continue;
      return false;
    }
    b += r;
    n -= (size_t)r;
  }
  return true;
}
static bool wr(int fd, const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  while (n) {
    ssize_t r = ::write(fd, b, n);
    if (r <= 0) {
      if (r < 0 && errno == EINTR) continue;
      return false;
    }
    b += r;
    n -= (size_t)r;
  }
  return true;
}

extern "C" {

FCS_PHMM_API int fcs_pairhmm_remote_open(const char* socket_path, fcs_phmm_remote** out) {
  if (!socket_path || !out) return FCS_PHMM_EINVAL;
  *out = nullptr;
  int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
  sockaddr_un addr;
  std::memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  std::strncpy(addr.sun_path, socket_path, sizeof(addr.sun_path) - 1);
  if (fd < 0 || ::connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
    g_cerr = std::string("cannot connect to the PairHMM daemon at ") + socket_path + ": " + std::strerror(errno) + " (no CPU fallback)";
    if (fd >= 0) ::close(fd);
    return FCS_PHMM_ENODEV;
  }
  *out = new fcs_phmm_remote{fd, ""};
  return FCS_PHMM_OK;
}

FCS_PHMM_API void fcs_pairhmm_remote_close(fcs_phmm_remote* r) {
  if (!r) return;
  ::close(r->fd);
  delete r;
}

FCS_PHMM_API const char* fcs_pairhmm_remote_last_error(const fcs_phmm_remote* r) { return r ? r->err.c_str() : g_cerr.c_str(); }

// Same contract as fcs_pairhmm_compute_flat, executed by the daemon.
FCS_PHMM_API int fcs_pairhmm_remote_compute_flat(fcs_phmm_remote* r, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64) {
  if (!r || !b || !out) return FCS_PHMM_EINVAL;
  std::vector<uint8_t> buf;
  auto w32 = [&](uint32_t v) {
    const uint8_t* q = reinterpret_cast<const uint8_t*>(&v);
    buf.insert(buf.end(), q, q + 4);
  };
  w32(0x4B4C4252u);  // 'RBLK'
  w32((uint32_t)b->n_regions);
  uint64_t pairs = 0;
  for (int64_t g = 0; g < b->n_regions; ++g) {
    const int32_t nr = b->reg_nreads[g], nh = b->reg_nhaps[g];
    if (b->reg_out0[g] != (int64_t)pairs) {
      r->err = "remote compute needs a dense output layout (reg_out0 = running sum of pairs)";
      return FCS_PHMM_EINVAL;
    }
    pairs += (uint64_t)nr * (uint64_t)nh;
    w32((uint32_t)nr);
    w32((uint32_t)nh);
    for (int32_t i = 0; i < nr; ++i) {
      const int64_t k = (int64_t)b->reg_read0[g] + i, o = b->rd_off[k];
      const uint32_t len = (uint32_t)b->rd_len[k];
      w32(len);
      const uint8_t* pl[5] = {b->read_bases + o, b->read_q + o, b->read_i + o, b->read_d + o, b->read_c + o};
      for (int p = 0; p < 5; ++p) buf.insert(buf.end(), pl[p], pl[p] + len);
    }
    for (int32_t j = 0; j < nh; ++j) {
      const int64_t k = (int64_t)b->reg_hap0[g] + j;
      const uint32_t len = (uint32_t)b->hp_len[k];
      w32(len);
      buf.insert(buf.end(), b->hap_bases + b->hp_off[k], b->hap_bases + b->hp_off[k] + len);
    }
  }
  const uint32_t rq = 0x51524850u;  // PHRQ
  const uint64_t len = buf.size();
  if (!wr(r->fd, &rq, 4) || !wr(r->fd, &len, 8) || !wr(r->fd, buf.data(), buf.size())) {
    r->err = "connection to the PairHMM daemon lost while sending";
    return FCS_PHMM_ENODEV;
  }
  uint32_t rs = 0;
  int32_t rc = 0;
  uint64_t n = 0;
  if (!rd(r->fd, &rs, 4) || rs != 0x53524850u || !rd(r->fd, &rc, 4) || !rd(r->fd, &n, 8)) {
    r->err = "connection to the PairHMM daemon lost while receiving";
    return FCS_PHMM_ENODEV;
  }
  if (rc != FCS_PHMM_OK) {
    std::string msg((size_t)n, '\0');
    rd(r->fd, &msg[0], (size_t)n);
    r->err = "daemon: " + msg;
    return rc;
  }
  if (n != pairs) {
    r->err = "daemon returned an unexpected number of pairs";
    return FCS_PHMM_EINVAL;
  }
  std::vector<uint8_t> flags((size_t)n);
  if (!rd(r->fd, out, (size_t)n * sizeof(double)) || !rd(r->fd, flags.data(), (size_t)n)) {
    r->err = "connection to the PairHMM daemon lost while receiving results";
    return FCS_PHMM_ENODEV;
  }
  if (used_fp64) std::memcpy(used_fp64, flags.data(), (size_t)n);
  return FCS_PHMM_OK;
}

}  // extern "C"
