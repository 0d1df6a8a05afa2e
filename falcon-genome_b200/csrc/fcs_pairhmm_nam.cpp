// fcs-pairhmm-nam — accelerator-manager daemon: ONE process owns the box's GPUs (a libfcs_pairhmm
// handle over all devices) and serves PairHMM requests from many client processes over a Unix socket.
//
// It takes the place of the Blaze NAM daemon in the reference's lifecycle:
// fcs-genome starts `<blaze.nam_path> <blaze.conf_path>` in the background before the HaplotypeCaller /
// Mutect2 fan-out and kills it with SIGALRM at scope exit
//   /root/reference/src/worker-htc.cpp:99-112, src/workers/BlazeWorker.cpp:22-26,
//   src/BackgroundExecutor.cpp:13-84 (kill(child, SIGALRM) at :79)
// so up to gatk.htc.nprocs (<= 32) JVMs share the accelerators through it.  Written against the
// public header only.
//
//   fcs-pairhmm-nam <socket path | conf file holding the socket path> [--devices N] [--double]
//
// Protocol (little endian), one request at a time per connection:
//   request : u32 'PHRQ', u64 payload bytes, payload = one capture block ('RBLK' ..., phmm_capture.h)
//   response: u32 'PHRS', i32 rc, u64 n;  rc == 0: n pairs -> n doubles then n flag bytes;
//                                          rc <  0: n bytes of error text
// Shared-memory transport (the default of libfcs_pairhmm_client; layout, doorbells and trust rules in
// phmm_shm.h): the client attaches a sealed memfd segment ('PHSM' + descriptor), writes each batch into it and
// rings 'PHSQ'; the daemon scores the batch in place and writes the results into the segment.
#include <atomic>
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/fcs_pairhmm.h"
#include "phmm_shm.h"

#ifndef F_GET_SEALS
#define F_GET_SEALS 1034
#define F_SEAL_SHRINK 0x0002
#endif

static std::atomic<bool> g_stop{false};
static int g_listen_fd = -1;
static void on_signal(int) {
  g_stop = true;
  if (g_listen_fd >= 0) ::shutdown(g_listen_fd, SHUT_RDWR);
}

static bool read_all(int fd, void* p, size_t n) {
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
static bool write_all(int fd, const void* p, size_t n) {
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

// Reads n bytes; a descriptor that arrives as ancillary data of those bytes is returned in *got_fd
// (a plain read() would silently drop it).
static bool read_with_fd(int fd, void* p, size_t n, int* got_fd) {
  uint8_t* b = static_cast<uint8_t*>(p);
  while (n) {
    iovec iov{b, n};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int) * 4)];
    msghdr mh;
    std::memset(&mh, 0, sizeof(mh));
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    mh.msg_control = ctl;
    mh.msg_controllen = sizeof(ctl);
    ssize_t r = ::recvmsg(fd, &mh, MSG_CMSG_CLOEXEC);
    if (r <= 0) {
      if (r < 0 && errno == EINTR) continue;
      return false;
    }
    for (cmsghdr* cm = CMSG_FIRSTHDR(&mh); cm; cm = CMSG_NXTHDR(&mh, cm)) {
      if (cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) continue;
      const size_t nfd = (cm->cmsg_len - CMSG_LEN(0)) / sizeof(int);
      for (size_t i = 0; i < nfd; ++i) {
        int f;
        std::memcpy(&f, CMSG_DATA(cm) + i * sizeof(int), sizeof(int));
        if (*got_fd < 0) *got_fd = f;
        else ::close(f);
      }
    }
    b += r;
    n -= (size_t)r;
  }
  return true;
}

static bool reply(int fd, int32_t rc, uint64_t n, const char* text) {
  const uint32_t rs = 0x53524850u;  // PHRS
  bool ok = write_all(fd, &rs, 4) && write_all(fd, &rc, 4);
  if (rc == FCS_PHMM_OK) return ok && write_all(fd, &n, 8);
  const uint64_t m = std::strlen(text);
  return ok && write_all(fd, &m, 8) && write_all(fd, text, m);
}

// A client's segment as mapped here, and the private copies of its index arrays.
struct Segment {
  uint8_t* p = nullptr;
  size_t bytes = 0;
  std::vector<int64_t> rd_off, hp_off, reg_out0;
  std::vector<int32_t> rd_len, hp_len, reg_read0, reg_nreads, reg_hap0, reg_nhaps;
  void unmap() {
    if (p) ::munmap(p, bytes);
    p = nullptr;
    bytes = 0;
  }
  ~Segment() { unmap(); }
};

// Maps the descriptor a client sent.  The segment must be sealed against shrinking: a client that truncates
// the file under a live mapping would otherwise turn the daemon's next access into SIGBUS.
static const char* attach_segment(Segment& sg, int mfd, uint64_t claimed) {
  struct stat st;
  if (mfd < 0) return "attach request without a descriptor";
  if (::fstat(mfd, &st) != 0 || (uint64_t)st.st_size < claimed || claimed < sizeof(fcsphmm::ShmHeader)) return "segment smaller than announced";
  const int seals = ::fcntl(mfd, F_GET_SEALS);
  if (seals < 0 || !(seals & F_SEAL_SHRINK)) return "segment is not sealed against shrinking (memfd with F_SEAL_SHRINK required)";
  void* m = ::mmap(nullptr, (size_t)claimed, PROT_READ | PROT_WRITE, MAP_SHARED, mfd, 0);
  if (m == MAP_FAILED) return "cannot map the segment";
  sg.unmap();
  sg.p = static_cast<uint8_t*>(m);
  sg.bytes = (size_t)claimed;
  return nullptr;
}

// Validates the batch in the segment and builds the flat view: bulk bytes stay in the segment, every index
// goes through a private copy.  Returns an error text or nullptr.
static const char* view_segment(Segment& sg, fcs_phmm_flat_batch& b, fcsphmm::ShmHeader& h) {
  if (!sg.p) return "no segment attached";
  std::memcpy(&h, sg.p, sizeof(h));
  if (h.magic != fcsphmm::kShmMagic || h.version != fcsphmm::kShmVersion) return "bad segment header";
  const uint64_t lim = 0x7fffffffULL;
  if (h.n_regions < 0 || h.n_reads < 0 || h.n_haps < 0 || (uint64_t)h.n_regions > lim || (uint64_t)h.n_reads > lim ||
      (uint64_t)h.n_haps > lim || h.read_bytes > sg.bytes || h.hap_bytes > sg.bytes || h.n_pairs > sg.bytes)
    return "segment header counts out of range";
  auto inside = [&](uint64_t off, uint64_t bytes, uint64_t align) {
    return off % align == 0 && off <= sg.bytes && bytes <= sg.bytes - off;
  };
  const uint64_t nr = (uint64_t)h.n_reads, nh = (uint64_t)h.n_haps, ng = (uint64_t)h.n_regions;
  if (!inside(h.off_read_bases, h.read_bytes, 1) || !inside(h.off_read_q, h.read_bytes, 1) || !inside(h.off_read_i, h.read_bytes, 1) ||
      !inside(h.off_read_d, h.read_bytes, 1) || !inside(h.off_read_c, h.read_bytes, 1) || !inside(h.off_hap_bases, h.hap_bytes, 1) ||
      !inside(h.off_rd_off, nr * 8, 8) || !inside(h.off_rd_len, nr * 4, 4) || !inside(h.off_hp_off, nh * 8, 8) ||
      !inside(h.off_hp_len, nh * 4, 4) || !inside(h.off_reg_read0, ng * 4, 4) || !inside(h.off_reg_nreads, ng * 4, 4) ||
      !inside(h.off_reg_hap0, ng * 4, 4) || !inside(h.off_reg_nhaps, ng * 4, 4) || !inside(h.off_out, h.n_pairs * 8, 8) ||
      !inside(h.off_used, h.n_pairs, 1))
    return "segment section outside the mapping";
  auto copy = [&](auto& v, uint64_t off, uint64_t n) {
    v.resize((size_t)n);
    if (n) std::memcpy(v.data(), sg.p + off, (size_t)n * sizeof(v[0]));
  };
  copy(sg.rd_off, h.off_rd_off, nr);
  copy(sg.rd_len, h.off_rd_len, nr);
  copy(sg.hp_off, h.off_hp_off, nh);
  copy(sg.hp_len, h.off_hp_len, nh);
  copy(sg.reg_read0, h.off_reg_read0, ng);
  copy(sg.reg_nreads, h.off_reg_nreads, ng);
  copy(sg.reg_hap0, h.off_reg_hap0, ng);
  copy(sg.reg_nhaps, h.off_reg_nhaps, ng);
  for (uint64_t k = 0; k < nr; ++k)
    if (sg.rd_len[k] < 0 || sg.rd_off[k] < 0 || (uint64_t)sg.rd_off[k] > h.read_bytes || (uint64_t)sg.rd_len[k] > h.read_bytes - (uint64_t)sg.rd_off[k])
      return "read outside its plane";
  for (uint64_t k = 0; k < nh; ++k)
    if (sg.hp_len[k] < 0 || sg.hp_off[k] < 0 || (uint64_t)sg.hp_off[k] > h.hap_bytes || (uint64_t)sg.hp_len[k] > h.hap_bytes - (uint64_t)sg.hp_off[k])
      return "haplotype outside its plane";
  sg.reg_out0.resize((size_t)ng);
  uint64_t pairs = 0;
  for (uint64_t g = 0; g < ng; ++g) {
    const int64_t r0 = sg.reg_read0[g], n_r = sg.reg_nreads[g], h0 = sg.reg_hap0[g], n_h = sg.reg_nhaps[g];
    if (r0 < 0 || n_r < 0 || h0 < 0 || n_h < 0 || (uint64_t)(r0 + n_r) > nr || (uint64_t)(h0 + n_h) > nh) return "region outside the read or haplotype tables";
    sg.reg_out0[g] = (int64_t)pairs;
    pairs += (uint64_t)n_r * (uint64_t)n_h;
    if (pairs > h.n_pairs) return "regions hold more pairs than the header announces";
  }
  if (pairs != h.n_pairs) return "pair count of the regions differs from the header";
  std::memset(&b, 0, sizeof(b));
  b.read_bases = sg.p + h.off_read_bases;
  b.read_q = sg.p + h.off_read_q;
  b.read_i = sg.p + h.off_read_i;
  b.read_d = sg.p + h.off_read_d;
  b.read_c = sg.p + h.off_read_c;
  b.rd_off = sg.rd_off.data();
  b.rd_len = sg.rd_len.data();
  b.n_reads = h.n_reads;
  b.hap_bases = sg.p + h.off_hap_bases;
  b.hp_off = sg.hp_off.data();
  b.hp_len = sg.hp_len.data();
  b.n_haps = h.n_haps;
  b.reg_read0 = sg.reg_read0.data();
  b.reg_nreads = sg.reg_nreads.data();
  b.reg_hap0 = sg.reg_hap0.data();
  b.reg_nhaps = sg.reg_nhaps.data();
  b.reg_out0 = sg.reg_out0.data();
  b.n_regions = h.n_regions;
  return nullptr;
}

// Byte-stream requests are sized by the client: bound what one request may make the daemon allocate.
static const uint64_t kMaxStreamPayload = 1ull << 31;  // 2 GiB of capture block (the engine's own per-chunk input limit)
static const uint64_t kMaxPairsPerRequest = 0x7fffffffULL;  // the engine's pair limit per chunk

static bool reply_stream_error(int fd, int32_t rc, const char* msg) {
  const uint32_t rs = 0x53524850u;  // PHRS
  const uint64_t m = std::strlen(msg);
  return write_all(fd, &rs, 4) && write_all(fd, &rc, 4) && write_all(fd, &m, 8) && write_all(fd, msg, m);
}

// One connection.  Returns when the client closes, the protocol is violated, or a reply cannot be written.
static void serve_loop(int fd, fcs_phmm_handle* h) {
  std::vector<uint8_t> payload;
  std::vector<double> out;
  std::vector<uint8_t> used;
  Segment sg;
  for (;;) {
    uint32_t magic = 0;
    uint64_t len = 0;
    int mfd = -1;
    if (!read_with_fd(fd, &magic, 4, &mfd) || !read_with_fd(fd, &len, 8, &mfd)) {
      if (mfd >= 0) ::close(mfd);
      break;
    }
    if (magic == fcsphmm::kShmAttach) {
      const char* err = attach_segment(sg, mfd, len);
      if (mfd >= 0) ::close(mfd);
      if (!reply(fd, err ? FCS_PHMM_EINVAL : FCS_PHMM_OK, 0, err ? err : "")) break;
      continue;
    }
    if (mfd >= 0) ::close(mfd);
    if (magic == fcsphmm::kShmRequest) {
      fcs_phmm_flat_batch b;
      fcsphmm::ShmHeader sh;
      const char* err = nullptr;
      int32_t rc = FCS_PHMM_OK;
      try {
        err = view_segment(sg, b, sh);
        rc = err ? FCS_PHMM_EINVAL : FCS_PHMM_OK;
        if (!err) {
          rc = fcs_pairhmm_compute_flat(h, &b, reinterpret_cast<double*>(sg.p + sh.off_out), sg.p + sh.off_used, nullptr);
          if (rc != FCS_PHMM_OK) err = fcs_pairhmm_last_error(h);
        }
      } catch (const std::bad_alloc&) {  // the private index copies of view_segment
        rc = FCS_PHMM_ENOMEM;
        err = "daemon out of host memory for this request";
      }
      if (!reply(fd, rc, err ? 0 : sh.n_pairs, err ? err : "")) break;
      continue;
    }
    if (magic != 0x51524850u /* PHRQ */) break;
    if (len > kMaxStreamPayload) {  // cannot skip that much and stay in sync: answer, then drop the connection
      reply_stream_error(fd, FCS_PHMM_EINVAL, "request payload larger than the daemon accepts");
      break;
    }
    fcs_phmm_flat_batch b;
    void* owner = nullptr;
    int32_t rc = FCS_PHMM_OK;
    uint64_t n = 0;
    const char* err = nullptr;
    try {
      payload.resize((size_t)len);
      if (!read_all(fd, payload.data(), payload.size())) break;
      rc = fcs_pairhmm_capture_parse(payload.data(), payload.size(), &b, &owner);
      if (rc != FCS_PHMM_OK) err = fcs_pairhmm_last_error(h);
      if (rc == FCS_PHMM_OK) {
        // zero-length reads / haplotypes cost four payload bytes each, so the pair count is NOT bounded by the
        // payload size: check it before sizing the result arrays
        for (int64_t g = 0; g < b.n_regions && n <= kMaxPairsPerRequest; ++g) n += (uint64_t)b.reg_nreads[g] * (uint64_t)b.reg_nhaps[g];
        if (n > kMaxPairsPerRequest) {
          rc = FCS_PHMM_EINVAL;
          err = "request holds more pairs than one call may carry (2^31 - 1)";
        }
      }
      if (rc == FCS_PHMM_OK) {
        out.resize((size_t)n);
        used.resize((size_t)n);
        rc = fcs_pairhmm_compute_flat(h, &b, out.data(), used.data(), nullptr);
        if (rc != FCS_PHMM_OK) err = fcs_pairhmm_last_error(h);
      }
    } catch (const std::bad_alloc&) {
      rc = FCS_PHMM_ENOMEM;
      err = "daemon out of host memory for this request";
    } catch (const std::length_error&) {
      rc = FCS_PHMM_EINVAL;
      err = "request too large";
    }
    bool ok;
    if (rc == FCS_PHMM_OK) {
      const uint32_t rs = 0x53524850u;  // PHRS
      ok = write_all(fd, &rs, 4) && write_all(fd, &rc, 4) && write_all(fd, &n, 8) && write_all(fd, out.data(), n * sizeof(double)) &&
           write_all(fd, used.data(), n);
    } else {
      ok = reply_stream_error(fd, rc, err ? err : "error");
    }
    fcs_pairhmm_capture_free(owner);
    if (!ok) break;
    if (out.capacity() > (64u << 20)) {  // do not let one large request pin its buffers for the connection's lifetime
      std::vector<double>().swap(out);
      std::vector<uint8_t>().swap(used);
      std::vector<uint8_t>().swap(payload);
    }
  }
}

// Thread body of a connection: nothing a client sends may take the daemon down (it owns the GPUs of every JVM
// of the stage, /root/reference/src/worker-htc.cpp:99-112); an exception ends this connection only.
static void serve(int fd, fcs_phmm_handle* h) {
  try {
    serve_loop(fd, h);
  } catch (...) {
    std::fprintf(stderr, "fcs-pairhmm-nam: connection %d dropped after an internal error\n", fd);
  }
  ::close(fd);
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <socket path | conf file> [--devices N] [--double]\n", argv[0]);
    return 2;
  }
  std::string path = argv[1];
  struct stat st;
  if (::stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode)) {  // conf file (reference: blaze.conf_path): first line = socket path
    std::FILE* f = std::fopen(path.c_str(), "r");
    char line[4096] = {0};
    if (f && std::fgets(line, sizeof(line), f)) {
      path = line;
      while (!path.empty() && (path.back() == '\n' || path.back() == '\r' || path.back() == ' ')) path.pop_back();
    }
    if (f) std::fclose(f);
  }
  int ndev = 0, use_double = 0;
  for (int i = 2; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--devices") && i + 1 < argc) ndev = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--double")) use_double = 1;
  }
  fcs_phmm_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = sizeof(cfg);
  cfg.n_devices = ndev;
  cfg.use_double = use_double;
  fcs_phmm_handle* h = nullptr;
  if (fcs_pairhmm_create(&cfg, &h) != FCS_PHMM_OK) {  // no CPU fallback: the daemon refuses to start
    std::fprintf(stderr, "fcs-pairhmm-nam: cannot start: %s\n", fcs_pairhmm_last_error(nullptr));
    return 3;
  }
  struct sigaction sa;
  std::memset(&sa, 0, sizeof(sa));
  sa.sa_handler = on_signal;
  sigaction(SIGALRM, &sa, nullptr);  // how the reference's BackgroundExecutor stops NAM
  sigaction(SIGTERM, &sa, nullptr);
  sigaction(SIGINT, &sa, nullptr);
  signal(SIGPIPE, SIG_IGN);
  g_listen_fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
  sockaddr_un addr;
  std::memset(&addr, 0, sizeof(addr));
  addr.sun_family = AF_UNIX;
  if (path.size() >= sizeof(addr.sun_path)) {
    std::fprintf(stderr, "fcs-pairhmm-nam: socket path too long\n");
    return 2;
  }
  std::strcpy(addr.sun_path, path.c_str());
  ::unlink(path.c_str());
  if (g_listen_fd < 0 || ::bind(g_listen_fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || ::listen(g_listen_fd, 64) != 0) {
    std::fprintf(stderr, "fcs-pairhmm-nam: cannot listen on %s: %s\n", path.c_str(), std::strerror(errno));
    fcs_pairhmm_destroy(h);
    return 1;
  }
  std::printf("fcs-pairhmm-nam ready on %s with %d device(s)\n", path.c_str(), fcs_pairhmm_device_count(h));
  std::fflush(stdout);
  while (!g_stop) {
    int fd = ::accept(g_listen_fd, nullptr, nullptr);
    if (fd < 0) {
      if (errno == EINTR) continue;
      break;
    }
    // detached at creation: a connection ends when its client closes, and a long stage with many reconnects
    // must not accumulate thread handles; the process exits with _exit() below
    try {
      std::thread(serve, fd, h).detach();
    } catch (...) {  // thread creation failed (resource limit): refuse this client, keep serving the others
      ::close(fd);
    }
  }
  ::close(g_listen_fd);
  ::unlink(path.c_str());
  fcs_phmm_stats s;
  if (fcs_pairhmm_get_stats(h, &s) == FCS_PHMM_OK)
    std::printf("fcs-pairhmm-nam stopping: %llu pairs, %llu cells, %llu chunks served\n", (unsigned long long)s.pairs, (unsigned long long)s.cells,
                (unsigned long long)s.chunks);
  std::fflush(stdout);
  _exit(0);
}
