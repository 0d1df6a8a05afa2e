// libfcs_pairhmm_client — CUDA-free client of the fcs-pairhmm-nam daemon (protocol in fcs_pairhmm_nam.cpp).
// What a JVM-side shim links when the GPUs are owned by the daemon instead of by the JVM itself.
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/syscall.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/fcs_pairhmm.h"
#include "phmm_shm.h"

#ifndef MFD_CLOEXEC
#define MFD_CLOEXEC 0x0001U
#endif
#ifndef MFD_ALLOW_SEALING
#define MFD_ALLOW_SEALING 0x0002U
#endif
#ifndef F_ADD_SEALS
#define F_ADD_SEALS 1033
#define F_SEAL_SHRINK 0x0002
#define F_SEAL_GROW 0x0004
#endif

struct fcs_phmm_remote {
  int fd;
  std::string err;
  bool shm = true;          // shared-memory transport (phmm_shm.h); false: byte-stream protocol
  uint8_t* seg = nullptr;   // this connection's segment, mapped here and in the daemon
  size_t seg_bytes = 0;
  uint64_t reserved_pairs = 0;  // fcs_pairhmm_remote_reserve: a batch is being built in the segment
  bool reserved = false;
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

// ---- shared-memory transport (layout and doorbells: phmm_shm.h) -----------------------------------
using fcsphmm::ShmHeader;
using fcsphmm::shm_align;

static bool read_reply(fcs_phmm_remote* r, int32_t& rc, uint64_t& n) {
  uint32_t rs = 0;
  return rd(r->fd, &rs, 4) && rs == 0x53524850u && rd(r->fd, &rc, 4) && rd(r->fd, &n, 8);
}

// A segment of at least `need` bytes, known to the daemon.  false: the shared path is unusable (r->shm is
// cleared, the caller falls back to the byte stream) or the connection is gone (r->err set).
static bool ensure_segment(fcs_phmm_remote* r, size_t need, bool& conn_lost) {
  conn_lost = false;
  if (r->seg && r->seg_bytes >= need) return true;
  const size_t bytes = (size_t)shm_align(std::max<size_t>(need + need / 2, (size_t)1 << 20));
  const int mfd = (int)::syscall(SYS_memfd_create, "fcs-pairhmm-client", MFD_CLOEXEC | MFD_ALLOW_SEALING);
  if (mfd < 0 || ::ftruncate(mfd, (off_t)bytes) != 0 || ::fcntl(mfd, F_ADD_SEALS, F_SEAL_SHRINK | F_SEAL_GROW) != 0) {
    if (mfd >= 0) ::close(mfd);
    r->shm = false;
    return false;
  }
  void* m = ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, mfd, 0);
  if (m == MAP_FAILED) {
    ::close(mfd);
    r->shm = false;
    return false;
  }
  // doorbell with the descriptor attached
  uint8_t msg[12];
  const uint32_t tag = fcsphmm::kShmAttach;
  const uint64_t b64 = bytes;
  std::memcpy(msg, &tag, 4);
  std::memcpy(msg + 4, &b64, 8);
  iovec iov{msg, sizeof(msg)};
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  std::memset(ctl, 0, sizeof(ctl));
  msghdr mh;
  std::memset(&mh, 0, sizeof(mh));
  mh.msg_iov = &iov;
  mh.msg_iovlen = 1;
  mh.msg_control = ctl;
  mh.msg_controllen = sizeof(ctl);
  cmsghdr* cm = CMSG_FIRSTHDR(&mh);
  cm->cmsg_level = SOL_SOCKET;
  cm->cmsg_type = SCM_RIGHTS;
  cm->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(cm), &mfd, sizeof(int));
  ssize_t sent;
  do sent = ::sendmsg(r->fd, &mh, MSG_NOSIGNAL); while (sent < 0 && errno == EINTR);
  ::close(mfd);  // the mappings keep the memory alive
  int32_t rc = 0;
  uint64_t n = 0;
  if (sent != (ssize_t)sizeof(msg) || !read_reply(r, rc, n)) {
    ::munmap(m, bytes);
    r->err = "connection to the PairHMM daemon lost while attaching the shared segment";
    conn_lost = true;
    return false;
  }
  if (rc != FCS_PHMM_OK) {  // the daemon declined (its text follows): use the byte stream from now on
    std::string txt((size_t)n, '\0');
    if (n) rd(r->fd, &txt[0], (size_t)n);
    ::munmap(m, bytes);
    r->shm = false;
    return false;
  }
  if (r->seg) ::munmap(r->seg, r->seg_bytes);
  r->seg = static_cast<uint8_t*>(m);
  r->seg_bytes = bytes;
  return true;
}

// Section offsets of a batch of the given shape (phmm_shm.h); returns the segment bytes it needs.
static uint64_t layout_segment(ShmHeader& h, int64_t n_regions, uint64_t n_reads, uint64_t n_haps, uint64_t read_bytes, uint64_t hap_bytes,
                               uint64_t pairs) {
  std::memset(&h, 0, sizeof(h));
  h.magic = fcsphmm::kShmMagic;
  h.version = fcsphmm::kShmVersion;
  h.n_regions = n_regions;
  h.n_reads = (int64_t)n_reads;
  h.n_haps = (int64_t)n_haps;
  h.n_pairs = pairs;
  h.read_bytes = read_bytes;
  h.hap_bytes = hap_bytes;
  uint64_t off = shm_align(sizeof(ShmHeader));
  auto section = [&](uint64_t bytes) { const uint64_t o = off; off = shm_align(off + bytes); return o; };
  h.off_read_bases = section(read_bytes);
  h.off_read_q = section(read_bytes);
  h.off_read_i = section(read_bytes);
  h.off_read_d = section(read_bytes);
  h.off_read_c = section(read_bytes);
  h.off_rd_off = section(n_reads * 8);
  h.off_rd_len = section(n_reads * 4);
  h.off_hap_bases = section(hap_bytes);
  h.off_hp_off = section(n_haps * 8);
  h.off_hp_len = section(n_haps * 4);
  h.off_reg_read0 = section((uint64_t)n_regions * 4);
  h.off_reg_nreads = section((uint64_t)n_regions * 4);
  h.off_reg_hap0 = section((uint64_t)n_regions * 4);
  h.off_reg_nhaps = section((uint64_t)n_regions * 4);
  h.off_out = section(pairs * 8);
  h.off_used = section(pairs);
  h.total_bytes = off;
  return off;
}

// Doorbell + reply for the batch described at offset 0 of the segment.  Returns the call's result code.
static int ring_segment(fcs_phmm_remote* r, uint64_t pairs) {
  uint8_t bell[12];
  const uint32_t tag = fcsphmm::kShmRequest;
  const uint64_t zero = 0;
  std::memcpy(bell, &tag, 4);
  std::memcpy(bell + 4, &zero, 8);
  int32_t rc = 0;
  uint64_t n = 0;
  if (!wr(r->fd, bell, sizeof(bell)) || !read_reply(r, rc, n)) {
    r->err = "connection to the PairHMM daemon lost";
    return FCS_PHMM_ENODEV;
  }
  if (rc != FCS_PHMM_OK) {
    std::string msg((size_t)n, '\0');
    if (n) rd(r->fd, &msg[0], (size_t)n);
    r->err = "daemon: " + msg;
    return rc;
  }
  if (n != pairs) {
    r->err = "daemon returned an unexpected number of pairs";
    return FCS_PHMM_EINVAL;
  }
  return FCS_PHMM_OK;
}

// 0 = done through the segment (rc in `result`), 1 = not possible, use the byte stream.
static int compute_via_shm(fcs_phmm_remote* r, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64, int& result) {
  uint64_t n_reads = 0, n_haps = 0, pairs = 0, read_bytes = 0, hap_bytes = 0;
  for (int64_t g = 0; g < b->n_regions; ++g) {
    const int32_t nr = b->reg_nreads[g], nh = b->reg_nhaps[g];
    if (nr < 0 || nh < 0) { r->err = "negative read or haplotype count"; result = FCS_PHMM_EINVAL; return 0; }
    if (b->reg_out0[g] != (int64_t)pairs) {
      r->err = "remote compute needs a dense output layout (reg_out0 = running sum of pairs)";
      result = FCS_PHMM_EINVAL;
      return 0;
    }
    pairs += (uint64_t)nr * (uint64_t)nh;
    for (int32_t i = 0; i < nr; ++i) {
      const int32_t len = b->rd_len[(int64_t)b->reg_read0[g] + i];
      if (len < 0) { r->err = "negative read length"; result = FCS_PHMM_EINVAL; return 0; }
      read_bytes += (uint64_t)len;
    }
    for (int32_t j = 0; j < nh; ++j) {
      const int32_t len = b->hp_len[(int64_t)b->reg_hap0[g] + j];
      if (len < 0) { r->err = "negative haplotype length"; result = FCS_PHMM_EINVAL; return 0; }
      hap_bytes += (uint64_t)len;
    }
    n_reads += (uint64_t)nr;
    n_haps += (uint64_t)nh;
  }
  if (n_reads > 0x7fffffffULL || n_haps > 0x7fffffffULL) return 1;
  ShmHeader h;
  const uint64_t off = layout_segment(h, b->n_regions, n_reads, n_haps, read_bytes, hap_bytes, pairs);
  bool lost = false;
  if (!ensure_segment(r, (size_t)off, lost)) {
    if (lost) { result = FCS_PHMM_ENODEV; return 0; }
    return 1;
  }
  uint8_t* s = r->seg;
  int64_t* rd_off = reinterpret_cast<int64_t*>(s + h.off_rd_off);
  int32_t* rd_len = reinterpret_cast<int32_t*>(s + h.off_rd_len);
  int64_t* hp_off = reinterpret_cast<int64_t*>(s + h.off_hp_off);
  int32_t* hp_len = reinterpret_cast<int32_t*>(s + h.off_hp_len);
  int32_t* g_r0 = reinterpret_cast<int32_t*>(s + h.off_reg_read0);
  int32_t* g_nr = reinterpret_cast<int32_t*>(s + h.off_reg_nreads);
  int32_t* g_h0 = reinterpret_cast<int32_t*>(s + h.off_reg_hap0);
  int32_t* g_nh = reinterpret_cast<int32_t*>(s + h.off_reg_nhaps);
  uint64_t rpos = 0, hpos = 0, ri = 0, hi = 0;
  for (int64_t g = 0; g < b->n_regions; ++g) {  // reads and haplotypes re-indexed densely, in region order
    const int32_t nr = b->reg_nreads[g], nh = b->reg_nhaps[g];
    g_r0[g] = (int32_t)ri;
    g_nr[g] = nr;
    g_h0[g] = (int32_t)hi;
    g_nh[g] = nh;
    for (int32_t i = 0; i < nr; ++i, ++ri) {
      const int64_t k = (int64_t)b->reg_read0[g] + i, o = b->rd_off[k];
      const size_t len = (size_t)b->rd_len[k];
      rd_off[ri] = (int64_t)rpos;
      rd_len[ri] = (int32_t)len;
      std::memcpy(s + h.off_read_bases + rpos, b->read_bases + o, len);
      std::memcpy(s + h.off_read_q + rpos, b->read_q + o, len);
      std::memcpy(s + h.off_read_i + rpos, b->read_i + o, len);
      std::memcpy(s + h.off_read_d + rpos, b->read_d + o, len);
      std::memcpy(s + h.off_read_c + rpos, b->read_c + o, len);
      rpos += len;
    }
    for (int32_t j = 0; j < nh; ++j, ++hi) {
      const int64_t k = (int64_t)b->reg_hap0[g] + j;
      const size_t len = (size_t)b->hp_len[k];
      hp_off[hi] = (int64_t)hpos;
      hp_len[hi] = (int32_t)len;
      std::memcpy(s + h.off_hap_bases + hpos, b->hap_bases + b->hp_off[k], len);
      hpos += len;
    }
  }
  std::memcpy(s, &h, sizeof(h));
  r->reserved = false;
  result = ring_segment(r, pairs);
  if (result != FCS_PHMM_OK) return 0;
  std::memcpy(out, s + h.off_out, (size_t)pairs * sizeof(double));
  if (used_fp64) std::memcpy(used_fp64, s + h.off_used, (size_t)pairs);
  result = FCS_PHMM_OK;
  return 0;
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
  fcs_phmm_remote* r = new fcs_phmm_remote();
  r->fd = fd;
  if (const char* e = std::getenv("FCS_PHMM_REMOTE_SHM")) r->shm = std::atoi(e) != 0;
  *out = r;
  return FCS_PHMM_OK;
}

FCS_PHMM_API void fcs_pairhmm_remote_close(fcs_phmm_remote* r) {
  if (!r) return;
  if (r->seg) ::munmap(r->seg, r->seg_bytes);
  ::close(r->fd);
  delete r;
}

/* 1 while requests travel through the shared segment, 0 on the byte-stream protocol. */
FCS_PHMM_API int fcs_pairhmm_remote_uses_shm(const fcs_phmm_remote* r) { return r && r->shm ? 1 : 0; }

FCS_PHMM_API const char* fcs_pairhmm_remote_last_error(const fcs_phmm_remote* r) { return r ? r->err.c_str() : g_cerr.c_str(); }

// Build the batch IN the segment: reserve() lays out a batch of the announced shape in this connection's segment
// and hands out writable views; the caller fills them (a JNI shim: GetByteArrayRegion straight into the planes --
// the one copy that leaves the JVM) and compute_reserved() rings the daemon; the results stay in the segment
// (views.out_log10 / out_used_fp64) until the next reserve or compute call on this connection.
FCS_PHMM_API int fcs_pairhmm_remote_reserve(fcs_phmm_remote* r, int64_t n_regions, int64_t n_reads, int64_t n_haps, uint64_t read_bytes,
                                            uint64_t hap_bytes, uint64_t n_pairs, fcs_phmm_remote_views* v) {
  if (!r || !v || n_regions < 0 || n_reads < 0 || n_haps < 0) return FCS_PHMM_EINVAL;
  r->reserved = false;
  if (!r->shm) {
    r->err = "in-segment batches need the shared-memory transport (FCS_PHMM_REMOTE_SHM=0 or the daemon declined the segment)";
    return FCS_PHMM_EUNSUPPORTED;
  }
  if ((uint64_t)n_reads > 0x7fffffffULL || (uint64_t)n_haps > 0x7fffffffULL || n_pairs > 0x7fffffffULL) {
    r->err = "batch too large for one call";
    return FCS_PHMM_EUNSUPPORTED;
  }
  ShmHeader h;
  const uint64_t need = layout_segment(h, n_regions, (uint64_t)n_reads, (uint64_t)n_haps, read_bytes, hap_bytes, n_pairs);
  bool lost = false;
  if (!ensure_segment(r, (size_t)need, lost)) {
    if (!lost) r->err = "the shared segment could not be set up";
    return lost ? FCS_PHMM_ENODEV : FCS_PHMM_EUNSUPPORTED;
  }
  uint8_t* s = r->seg;
  std::memcpy(s, &h, sizeof(h));
  v->read_bases = s + h.off_read_bases; v->read_q = s + h.off_read_q; v->read_i = s + h.off_read_i;
  v->read_d = s + h.off_read_d; v->read_c = s + h.off_read_c;
  v->rd_off = reinterpret_cast<int64_t*>(s + h.off_rd_off); v->rd_len = reinterpret_cast<int32_t*>(s + h.off_rd_len);
  v->hap_bases = s + h.off_hap_bases;
  v->hp_off = reinterpret_cast<int64_t*>(s + h.off_hp_off); v->hp_len = reinterpret_cast<int32_t*>(s + h.off_hp_len);
  v->reg_read0 = reinterpret_cast<int32_t*>(s + h.off_reg_read0); v->reg_nreads = reinterpret_cast<int32_t*>(s + h.off_reg_nreads);
  v->reg_hap0 = reinterpret_cast<int32_t*>(s + h.off_reg_hap0); v->reg_nhaps = reinterpret_cast<int32_t*>(s + h.off_reg_nhaps);
  v->out_log10 = reinterpret_cast<const double*>(s + h.off_out);
  v->out_used_fp64 = s + h.off_used;
  r->reserved = true;
  r->reserved_pairs = n_pairs;
  return FCS_PHMM_OK;
}

FCS_PHMM_API int fcs_pairhmm_remote_compute_reserved(fcs_phmm_remote* r) {
  if (!r) return FCS_PHMM_EINVAL;
  if (!r->reserved) {
    r->err = "no batch reserved on this connection";
    return FCS_PHMM_EINVAL;
  }
  return ring_segment(r, r->reserved_pairs);  // (the reservation stays valid: the same batch may be filled again)
}

// Same contract as fcs_pairhmm_compute_flat, executed by the daemon.
FCS_PHMM_API int fcs_pairhmm_remote_compute_flat(fcs_phmm_remote* r, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64) {
  if (!r || !b || !out) return FCS_PHMM_EINVAL;
  if (r->shm) {
    int result = FCS_PHMM_OK;
    if (compute_via_shm(r, b, out, used_fp64, result) == 0) return result;
  }
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
