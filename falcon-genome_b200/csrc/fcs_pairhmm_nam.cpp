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
#include <atomic>
#include <cerrno>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/fcs_pairhmm.h"

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

static void serve(int fd, fcs_phmm_handle* h) {
  std::vector<uint8_t> payload;
  std::vector<double> out;
  std::vector<uint8_t> used;
  for (;;) {
    uint32_t magic = 0;
    uint64_t len = 0;
    if (!read_all(fd, &magic, 4) || magic != 0x51524850u /* PHRQ */ || !read_all(fd, &len, 8) || len > (1ull << 32)) break;
    payload.resize((size_t)len);
    if (!read_all(fd, payload.data(), payload.size())) break;
    fcs_phmm_flat_batch b;
    void* owner = nullptr;
    int32_t rc = fcs_pairhmm_capture_parse(payload.data(), payload.size(), &b, &owner);
    uint64_t n = 0;
    if (rc == FCS_PHMM_OK) {
      for (int64_t g = 0; g < b.n_regions; ++g) n += (uint64_t)b.reg_nreads[g] * (uint64_t)b.reg_nhaps[g];
      out.resize((size_t)n);
      used.resize((size_t)n);
      rc = fcs_pairhmm_compute_flat(h, &b, out.data(), used.data(), nullptr);
    }
    const uint32_t rs = 0x53524850u;  // PHRS
    bool ok = write_all(fd, &rs, 4) && write_all(fd, &rc, 4);
    if (rc == FCS_PHMM_OK) {
      ok = ok && write_all(fd, &n, 8) && write_all(fd, out.data(), n * sizeof(double)) && write_all(fd, used.data(), n);
    } else {
      const char* msg = fcs_pairhmm_last_error(h);
      const uint64_t m = std::strlen(msg);
      ok = ok && write_all(fd, &m, 8) && write_all(fd, msg, m);
    }
    fcs_pairhmm_capture_free(owner);
    if (!ok) break;
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
  std::vector<std::thread> conns;
  while (!g_stop) {
    int fd = ::accept(g_listen_fd, nullptr, nullptr);
    if (fd < 0) {
      if (errno == EINTR) continue;
      break;
    }
    conns.emplace_back(serve, fd, h);
  }
  ::close(g_listen_fd);
  ::unlink(path.c_str());
  for (auto& t : conns)
    if (t.joinable()) t.detach();  // connections end when their clients close; the process exits now
  fcs_phmm_stats s;
  if (fcs_pairhmm_get_stats(h, &s) == FCS_PHMM_OK)
    std::printf("fcs-pairhmm-nam stopping: %llu pairs, %llu cells, %llu chunks served\n", (unsigned long long)s.pairs, (unsigned long long)s.cells,
                (unsigned long long)s.chunks);
  std::fflush(stdout);
  _exit(0);
}
