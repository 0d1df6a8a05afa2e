// phmm_engine.cu — packer (batcher), per-device chunk pipeline, multi-GPU dispatcher.
//
// The reference fans HaplotypeCaller / Mutect2 out as independent processes per genome
// partition (/root/reference/src/worker-htc.cpp:113-145, src/Executor.cpp:50-108); every
// active region, read and haplotype pair is independent, so here regions are partitioned
// over the devices by cell count with no device-to-device exchange.
//
// Per device, `slots` chunk pipelines run concurrently on their own streams:
//   host pack (pinned) -> H2D -> FP32 wavefront kernels (one launch per kernel class)
//   -> FP64 rerun kernels (drain the per-class queues the FP32 kernels filled) -> D2H
//   -> host scatter into the caller's out_log10.
// Packing chunk k+1 overlaps the GPU work of chunk k.
#include "phmm_engine.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include <sched.h>

#include "phmm_capture.h"
#include "phmm_luts.h"

namespace fcsphmm {

// ---------------------------------------------------------------------------------------
// error text (thread local, as the ABI promises)
static thread_local std::string g_err;
int set_error(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
const char* last_error() { return g_err.c_str(); }

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return set_error(FCS_PHMM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int64_t env_i64(const char* name, int64_t dflt) {
  const char* v = std::getenv(name);
  if (!v || !*v) return dflt;
  return std::strtoll(v, nullptr, 10);
}

// class lookup by read length (hot in the planner: one lookup per read)
static std::vector<const ClassRef*> g_f32_by_len[4];  // [form]
static std::vector<const ClassRef*> g_f32_coarse_by_len[4];
static std::vector<int16_t> g_qid_by_len;             // FP64 queue id
static std::once_flag g_cls_once;
static void build_len_tables() {
  for (int form = 0; form < 4; ++form) {
    g_f32_by_len[form].assign(1025, nullptr);
    for (int len = 1; len <= 1024; ++len) g_f32_by_len[form][len] = select_class(false, form, len);
    g_f32_coarse_by_len[form].assign(1025, nullptr);
    for (int len = 1; len <= 1024; ++len) g_f32_coarse_by_len[form][len] = select_class(false, form, len, true);
  }
  g_qid_by_len.assign(1025, -1);
  for (int len = 1; len <= 1024; ++len)
    if (const ClassRef* k = select_class(true, false, len)) g_qid_by_len[len] = (int16_t)f64_queue_id(k->G, k->R);
}
static inline const ClassRef* f32_class_of_len(int form, int len) { return (len >= 1 && len <= 1024) ? g_f32_by_len[form][len] : nullptr; }
static inline const ClassRef* f32_coarse_class_of_len(int form, int len) { return (len >= 1 && len <= 1024) ? g_f32_coarse_by_len[form][len] : nullptr; }
static inline int qid_of_len(int len) { return (len >= 1 && len <= 1024) ? g_qid_by_len[len] : -1; }

// The regions of a chunk are visited in the chunk's order, not in memory order: every region starts with a cache miss on each
// plane it touches (a full-size config-3 stream is 9 GB of distinct input).  Pull the first lines of the NEXT region's planes
// while the current one is being worked on.  planes: bit 0 bases, 1 base qualities, 2 insertion, 3 deletion, 4 continuation.
static const bool g_prefetch = env_i64("FCS_PHMM_NO_PREFETCH", 0) == 0;  // developer knob
static inline void prefetch_region(const Input& in, int64_t g, unsigned planes, bool haps) {
  if (!g_prefetch) return;
  int32_t nr = 0, nh = 0;
  in.shape(g, nr, nh);
  if (nr > 0) {
    const InRead r = in.read(g, 0);
    const uint8_t* pl[5] = {r.b, r.q, r.i, r.d, r.c};
    for (int k = 0; k < 5; ++k)
      if ((planes >> k & 1u) && pl[k])
        for (int off = 0; off < 256; off += 64) __builtin_prefetch(pl[k] + off, 0, 1);
  }
  if (haps && nh > 0) {
    const InHap h = in.hap(g, 0);
    if (h.b)
      for (int off = 0; off < 256; off += 64) __builtin_prefetch(h.b + off, 0, 1);
  }
}

// One pass over the three transition-quality planes of a read (runs once per read of every call):
//   gcp / ins / del = the value all bytes of the plane share (masked & 127 like the kernels), or -1;
//   same_indel      = the deletion plane equals the insertion plane byte for byte.
struct QualScan { int gcp, ins, del; bool same_indel; };
static inline QualScan scan_quals(const uint8_t* c, const uint8_t* i, const uint8_t* d, int32_t len) {
  QualScan r;
#if defined(__SSE2__)
  if (len >= 16) {
    const __m128i c0 = _mm_set1_epi8((char)c[0]), i0 = _mm_set1_epi8((char)i[0]), d0 = _mm_set1_epi8((char)d[0]);
    __m128i ac = _mm_setzero_si128(), ai = ac, ad = ac, aid = ac;
    auto block = [&](int32_t x) {
      const __m128i vc = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c + x));
      const __m128i vi = _mm_loadu_si128(reinterpret_cast<const __m128i*>(i + x));
      const __m128i vd = _mm_loadu_si128(reinterpret_cast<const __m128i*>(d + x));
      ac = _mm_or_si128(ac, _mm_xor_si128(vc, c0));
      ai = _mm_or_si128(ai, _mm_xor_si128(vi, i0));
      ad = _mm_or_si128(ad, _mm_xor_si128(vd, d0));
      aid = _mm_or_si128(aid, _mm_xor_si128(vi, vd));
    };
    int32_t x = 0;
    for (; x + 16 <= len; x += 16) block(x);
    if (x < len) block(len - 16);  // the last, partly filled block overlaps the previous one
    const __m128i m7 = _mm_set1_epi8(0x7f), z = _mm_setzero_si128();
    r.gcp = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_and_si128(ac, m7), z)) == 0xffff ? (int)(c[0] & 127u) : -1;
    r.ins = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_and_si128(ai, m7), z)) == 0xffff ? (int)(i[0] & 127u) : -1;
    r.del = _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_and_si128(ad, m7), z)) == 0xffff ? (int)(d[0] & 127u) : -1;
    r.same_indel = _mm_movemask_epi8(_mm_cmpeq_epi8(aid, z)) == 0xffff;
    return r;
  }
#endif
  uint8_t ac = 0, ai = 0, ad = 0, aid = 0;
  for (int32_t x = 0; x < len; ++x) {
    ac |= (uint8_t)(c[x] ^ c[0]); ai |= (uint8_t)(i[x] ^ i[0]); ad |= (uint8_t)(d[x] ^ d[0]); aid |= (uint8_t)(i[x] ^ d[x]);
  }
  r.gcp = (ac & 127u) ? -1 : (int)(c[0] & 127u);
  r.ins = (ai & 127u) ? -1 : (int)(i[0] & 127u);
  r.del = (ad & 127u) ? -1 : (int)(d[0] & 127u);
  r.same_indel = aid == 0;
  return r;
}

// dst[0, round_up16(len)) = src[0, len) followed by `pad` bytes.  The packer runs this ten times per read on
// 100-250 byte pieces: 16-byte loads/stores inline instead of a memcpy + memset call pair each.  Never reads
// past src + len (the caller's arrays may end there).
static inline void copy_padded16(uint8_t* dst, const uint8_t* src, uint32_t len, uint8_t pad) {
#if defined(__SSE2__)
  uint32_t i = 0;
  // (non-temporal stores were tried for the staging buffer -- it is only ever read by the DMA engine -- and lost by
  // 10x: planes are 16-byte, not 64-byte aligned, so regular tail stores and the next plane share cache lines with
  // the streamed ones and every write-combining buffer is flushed partially filled)
  for (; i + 16 <= len; i += 16) _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
  if (i < len) {
    // last, partly filled block: the padding first, then the final 16 source bytes on top of it (they overlap the
    // previous block with identical data) -- no byte loop, no read or write outside [0, len) / [0, round_up16(len))
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_set1_epi8((char)pad));
    if (len >= 16) _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + len - 16), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + len - 16)));
    else for (uint32_t k = 0; k < len; ++k) dst[k] = src[k];
  }
#else
  std::memcpy(dst, src, len);
  std::memset(dst + len, pad, round_up16(len) - len);
#endif
}

// Which byte classes a haplotype holds: bit 0 = an N, bit 1 = a byte outside ACGTN.
static inline uint32_t hap_classes(const uint8_t* b, int32_t len) {
  uint32_t seen = 0;
  int32_t i = 0;
#if defined(__SSE2__)
  const __m128i vA = _mm_set1_epi8('A'), vC = _mm_set1_epi8('C'), vG = _mm_set1_epi8('G'), vT = _mm_set1_epi8('T'), vN = _mm_set1_epi8('N');
  for (; i + 16 <= len; i += 16) {
    const __m128i w = _mm_loadu_si128(reinterpret_cast<const __m128i*>(b + i));
    const __m128i acgt = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(w, vA), _mm_cmpeq_epi8(w, vC)), _mm_or_si128(_mm_cmpeq_epi8(w, vG), _mm_cmpeq_epi8(w, vT)));
    const int mn = _mm_movemask_epi8(_mm_cmpeq_epi8(w, vN));
    const int ma = _mm_movemask_epi8(acgt);
    if (mn) seen |= 1u;
    if ((ma | mn) != 0xffff) seen |= 2u;
  }
#endif
  for (; i < len; ++i) {
    const uint8_t c = b[i];
    if (c == 'N') seen |= 1u;
    else if (c != 'A' && c != 'C' && c != 'G' && c != 'T') seen |= 2u;
  }
  return seen;
}

static cudaError_t init_slot(Slot& s);

// Cores of the NUMA node a device is attached to (sysfs), restricted to the process's affinity mask.
static bool device_node_cpus(int ordinal, cpu_set_t* out) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), ordinal) != cudaSuccess) return false;
  for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
  int node = -1;
  {
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    std::FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
  }
  if (node < 0) return false;
  char list[4096] = {0};
  {
    const std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    std::FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    const bool ok = std::fgets(list, sizeof(list), f) != nullptr;
    std::fclose(f);
    if (!ok) return false;
  }
  cpu_set_t mine, node_set;
  CPU_ZERO(&mine);
  CPU_ZERO(&node_set);
  if (sched_getaffinity(0, sizeof(mine), &mine) != 0) return false;
  for (const char* p = list; *p;) {  // "0-15,32-47"
    char* end = nullptr;
    long a = std::strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    if (*end == '-') { p = end + 1; b = std::strtol(p, &end, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
      if (c >= 0 && CPU_ISSET((int)c, &mine)) CPU_SET((int)c, &node_set);
    p = (*end == ',') ? end + 1 : end;
    if (*end != ',') break;
  }
  if (CPU_COUNT(&node_set) == 0) return false;
  *out = node_set;
  return true;
}

static thread_local bool t_pool_thread = false;  // set by WorkerPool::loop: only the library's own threads are ever re-bound

int ChunkPlan::launches() const {
  int n = 0;
  if (!force_double) {
    for (const auto& r : f32) n += r.n_tasks ? 1 : 0;
    n += n_gen ? 1 : 0;
  }
  n += (int)f64.size() + (gen64_cap ? 1 : 0);
  return n;
}

// ---------------------------------------------------------------------------------------
Engine::Engine() {}

int Engine::set_capture(const char* path) {
  if (!path || !*path) {
    capture_.reset();
    return FCS_PHMM_OK;
  }
  std::unique_ptr<CaptureWriter> w(new CaptureWriter());
  int rc = w->open(path);
  if (rc != FCS_PHMM_OK) return rc;
  capture_ = std::move(w);
  return FCS_PHMM_OK;
}

int Engine::create(const fcs_phmm_config* cfg, Engine** out) {
  *out = nullptr;
  std::unique_ptr<Engine> e(new Engine());
  int rc = e->init(cfg);
  if (rc != FCS_PHMM_OK) return rc;
  *out = e.release();
  return FCS_PHMM_OK;
}

int Engine::init(const fcs_phmm_config* cfg) {
  std::call_once(g_cls_once, build_len_tables);
  if (f64_queue_count() > kMaxF64Classes) return set_error(FCS_PHMM_EINVAL, "too many FP64 kernel classes compiled in");
  fcs_phmm_config c;
  std::memset(&c, 0, sizeof(c));
  if (cfg) std::memcpy(&c, cfg, std::min<size_t>(sizeof(c), cfg->struct_size ? cfg->struct_size : sizeof(c)));
  use_double_ = c.use_double != 0;
  keep_raw_ = c.keep_raw_f32 != 0;
  pack_threads_ = c.max_threads > 0 ? c.max_threads : (int)env_i64("FCS_PHMM_PACK_THREADS", 0);
  max_chunk_cells_ = c.max_chunk_cells > 0 ? c.max_chunk_cells : env_i64("FCS_PHMM_CHUNK_CELLS", 0);  // 0 = adaptive
  if (const char* cap = std::getenv("FCS_PHMM_CAPTURE")) {
    int rc = set_capture(cap);
    if (rc != FCS_PHMM_OK) return rc;
  }

  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return set_error(FCS_PHMM_ENODEV, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                                          " (libfcs_pairhmm has no CPU fallback)");
  std::vector<int> ords;
  if (c.n_devices > 0) {
    for (int i = 0; i < c.n_devices; ++i) ords.push_back(c.devices ? c.devices[i] : i);
  } else {
    for (int i = 0; i < ndev; ++i) ords.push_back(i);
  }
  if (pack_threads_ <= 0) {
    // Default: GATK's --native-pair-hmm-threads default of 4 per device, but never more threads than this
    // process has cores for.  Oversubscribed packing threads and their event waits fight for the same cores
    // (measured with 2 cores per GPU: 4 threads 3.7, 2 threads 5.3 TCUPS end to end on two GPUs).
    cpu_set_t set;
    CPU_ZERO(&set);
    int cores = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    pack_threads_ = std::max(1, std::min(4, cores / std::max<int>(1, (int)ords.size())));
  }
  // two slots per packing thread: a thread packs into one while its previous chunk is on the device
  const int nslots = std::max(c.slots_per_device > 0 ? c.slots_per_device : 0, 2 * pack_threads_);
  const Luts& L = luts();
  for (int ord : ords) {
    if (ord < 0 || ord >= ndev) return set_error(FCS_PHMM_EINVAL, "device ordinal out of range");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ord));
    if (prop.major != 10)
      return set_error(FCS_PHMM_ENODEV, std::string("device ") + std::to_string(ord) + " (" + prop.name + ") is sm_" +
                                            std::to_string(prop.major * 10 + prop.minor) + "; the kernels are built for sm_100a only");
    std::unique_ptr<Device> d(new Device());
    d->ordinal = ord;
    d->sm_count = prop.multiProcessorCount;
    CK(cudaSetDevice(ord));
    CK(cudaMalloc(&d->d_ph2pr_f, sizeof(L.ph2pr_f)));
    CK(cudaMalloc(&d->d_mm_f, sizeof(L.mm_f)));
    CK(cudaMalloc(&d->d_ph2pr_d, sizeof(L.ph2pr_d)));
    CK(cudaMalloc(&d->d_mm_d, sizeof(L.mm_d)));
    CK(cudaMemcpy(d->d_ph2pr_f, L.ph2pr_f, sizeof(L.ph2pr_f), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->d_mm_f, L.mm_f, sizeof(L.mm_f), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->d_ph2pr_d, L.ph2pr_d, sizeof(L.ph2pr_d), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->d_mm_d, L.mm_d, sizeof(L.mm_d), cudaMemcpyHostToDevice));
    {
      int nk = 0;
      const TierKernel* const* tks = tier_kernels(&nk);
      for (int i = 0; i < nk; ++i) CK(tks[i]->set_max_smem(prop.sharedMemPerBlockOptin));
    }
    d->slots = std::vector<Slot>((size_t)nslots);
    d->use.assign((size_t)pack_threads_, 0);
    static const bool numa_bind = env_i64("FCS_PHMM_NUMA_BIND", 0) != 0;  // opt-in, see Device::node_cpus
    cpu_set_t before;
    bool rebound = false;
    if (numa_bind && device_node_cpus(ord, &d->node_cpus)) {
      d->has_node_cpus = true;
      // allocate the pinned staging from the device's node (first touch happens when the pages are pinned)
      CPU_ZERO(&before);
      if (sched_getaffinity(0, sizeof(before), &before) == 0 && sched_setaffinity(0, sizeof(cpu_set_t), &d->node_cpus) == 0) rebound = true;
      if (env_i64("FCS_PHMM_DEBUG", 0)) fprintf(stderr, "[fcs_phmm] device %d: packing threads bound to %d cores of its NUMA node\n", ord, CPU_COUNT(&d->node_cpus));
    }
    cudaError_t slot_err = cudaSuccess;
    for (Slot& s : d->slots)
      if ((slot_err = init_slot(s)) != cudaSuccess) break;
    if (rebound) sched_setaffinity(0, sizeof(before), &before);
    CK(slot_err);
    devs_.push_back(std::move(d));
  }
  pool_.reset(new WorkerPool(std::max(0, (int)devs_.size() * pack_threads_ - 1)));
  return FCS_PHMM_OK;
}

static cudaError_t init_slot(Slot& s) {
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&s.ev_k0)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&s.ev_k1)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&s.ev_k2)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&s.ev_done)) != cudaSuccess) return e;
  if ((e = cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
  {
    int prio_lo = 0, prio_hi = 0;
    if ((e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)) != cudaSuccess) return e;
    for (int i = 0; i < Slot::kHp; ++i) {
      if ((e = cudaStreamCreateWithPriority(&s.hp[i], cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&s.ev_hp[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
  }
  for (int i = 0; i < Slot::kSide; ++i) {
    if ((e = cudaStreamCreateWithFlags(&s.side[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&s.ev_side[i], cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

static void free_slot(Slot& s) {
  if (s.stream) cudaStreamSynchronize(s.stream);
  for (int i = 0; i < Slot::kHp; ++i) {
    if (s.hp[i]) { cudaStreamSynchronize(s.hp[i]); cudaStreamDestroy(s.hp[i]); }
    if (s.ev_hp[i]) cudaEventDestroy(s.ev_hp[i]);
  }
  for (int i = 0; i < Slot::kSide; ++i) {
    if (s.side[i]) { cudaStreamSynchronize(s.side[i]); cudaStreamDestroy(s.side[i]); }
    if (s.ev_side[i]) cudaEventDestroy(s.ev_side[i]);
  }
  if (s.ev_fork) cudaEventDestroy(s.ev_fork);
  if (s.h_in) cudaFreeHost(s.h_in);
  if (s.h_out) cudaFreeHost(s.h_out);
  if (s.d_buf) cudaFree(s.d_buf);
  if (s.ev_k0) cudaEventDestroy(s.ev_k0);
  if (s.ev_k1) cudaEventDestroy(s.ev_k1);
  if (s.ev_k2) cudaEventDestroy(s.ev_k2);
  if (s.ev_done) cudaEventDestroy(s.ev_done);
  if (s.stream) cudaStreamDestroy(s.stream);
  // back to the state of a fresh slot (not `s = Slot()`: the slot's mutex is neither copyable nor movable)
  s.stream = nullptr;
  s.ev_k0 = s.ev_k1 = s.ev_k2 = s.ev_done = s.ev_fork = nullptr;
  for (int i = 0; i < Slot::kSide; ++i) { s.side[i] = nullptr; s.ev_side[i] = nullptr; }
  for (int i = 0; i < Slot::kHp; ++i) { s.hp[i] = nullptr; s.ev_hp[i] = nullptr; }
  s.h_in = s.h_out = s.d_buf = nullptr;
  s.h_in_cap = s.h_out_cap = s.d_cap = 0;
  s.busy = s.timed = false;
  s.owner = nullptr;
  s.input = nullptr;
  s.plan = ChunkPlan();
}

Engine::~Engine() {
  {
    std::lock_guard<std::mutex> lk(tickets_mu_);
    for (auto& kv : tickets_)
      if (kv.second->th.joinable()) kv.second->th.join();
    tickets_.clear();
  }
  pool_.reset();
  for (auto& d : devs_) {
    cudaSetDevice(d->ordinal);
    for (Slot& s : d->slots) free_slot(s);
    cudaFree(d->d_ph2pr_f);
    cudaFree(d->d_mm_f);
    cudaFree(d->d_ph2pr_d);
    cudaFree(d->d_mm_d);
  }
}

int Engine::ensure_buffers(Slot& s, size_t in_bytes, size_t out_bytes, size_t dev_bytes) {
  // Growing a slot means cudaFreeHost + cudaHostAlloc (+ cudaFree + cudaMalloc): milliseconds each, and device-wide
  // synchronisation.  So the first growth already jumps to a floor that covers a regular chunk of any BASELINE shape
  // (3 Gcells: ~10 MB in, ~2 MB out), and later growth adds half again: a stream of calls of varying size settles
  // after one or two calls instead of re-allocating whenever a slightly larger chunk shows up.
  auto grow = [](size_t need, size_t floor_bytes) { return align_up(std::max(need + need / 2 + 4096, floor_bytes), 4096); };
  static const size_t floor_in = (size_t)env_i64("FCS_PHMM_SLOT_FLOOR_MB", 16) << 20;
  const size_t floor_out = floor_in / 4, floor_dev = floor_in * 2;
  static const bool dbg_grow = env_i64("FCS_PHMM_DEBUG", 0) != 0;
  if (dbg_grow && (in_bytes > s.h_in_cap || out_bytes > s.h_out_cap || dev_bytes > s.d_cap))
    fprintf(stderr, "[fcs_phmm] slot grows: in %zu -> %zu, out %zu -> %zu, dev %zu -> %zu bytes\n", s.h_in_cap, in_bytes, s.h_out_cap, out_bytes, s.d_cap, dev_bytes);
  if (in_bytes > s.h_in_cap) {
    if (s.h_in) CK(cudaFreeHost(s.h_in));
    s.h_in = nullptr;
    s.h_in_cap = grow(in_bytes, floor_in);
    if (cudaHostAlloc((void**)&s.h_in, s.h_in_cap, cudaHostAllocDefault) != cudaSuccess) {
      s.h_in_cap = 0;
      return set_error(FCS_PHMM_ENOMEM, "pinned host allocation failed");
    }
  }
  if (out_bytes > s.h_out_cap) {
    if (s.h_out) CK(cudaFreeHost(s.h_out));
    s.h_out = nullptr;
    s.h_out_cap = grow(out_bytes, floor_out);
    if (cudaHostAlloc((void**)&s.h_out, s.h_out_cap, cudaHostAllocDefault) != cudaSuccess) {
      s.h_out_cap = 0;
      return set_error(FCS_PHMM_ENOMEM, "pinned host allocation failed");
    }
  }
  if (dev_bytes > s.d_cap) {
    if (s.d_buf) CK(cudaFree(s.d_buf));
    s.d_buf = nullptr;
    s.d_cap = grow(dev_bytes, floor_dev);
    if (cudaMalloc((void**)&s.d_buf, s.d_cap) != cudaSuccess) {
      s.d_cap = 0;
      return set_error(FCS_PHMM_ENOMEM, "device allocation failed");
    }
  }
  return FCS_PHMM_OK;
}

// developer knob: FCS_PHMM_NO_TWO_PLANE=1 packs all five planes of every read
static bool two_plane_enabled() {
  static const bool on = env_i64("FCS_PHMM_NO_TWO_PLANE", 0) == 0;
  return on;
}

// haplotype byte classes: 0 = ACGT, bit 0 = N, bit 1 = any other byte
static const uint8_t* hap_byte_class() {
  static uint8_t lut[256];
  static std::once_flag once;
  std::call_once(once, [] {
    std::memset(lut, 2, sizeof(lut));
    lut[(int)'A'] = lut[(int)'C'] = lut[(int)'G'] = lut[(int)'T'] = 0;
    lut[(int)'N'] = 1;
  });
  return lut;
}

// ---------------------------------------------------------------------------------------
// Pass 1: choose the regions of the chunk, order each region's reads by length, cut the
// (read group x haplotype run) tasks per kernel class and size every section.
namespace {

// developer probe (FCS_PHMM_PLAN_PROF=1): where the planner's time goes, summed per thread and printed by plan_check
struct PlanProf { double scan = 0, region = 0, sort = 0, tasks = 0, f64 = 0, post = 0; };
static thread_local PlanProf t_prof;
static const bool g_plan_prof = env_i64("FCS_PHMM_PLAN_PROF", 0) != 0;
struct ProfScope {
  double* acc; double t0;
  explicit ProfScope(double* a) : acc(g_plan_prof ? a : nullptr), t0(acc ? now_ms() : 0) {}
  ~ProfScope() { if (acc) *acc += now_ms() - t0; }
};

struct Planner {
  const Input& in;
  Slot& s;
  bool force_double;
  bool keep_raw;
  int64_t max_cells;
  uint32_t hs_cols;  // haplotype columns per task (bounds the shared-memory stream)
  int sm_count = 148;
  bool shape_tail = true;  // false for every chunk of a call but the last one on its device: their tails overlap the next chunk
  bool coarse_classes = true;  // unpopular read lengths round up to the coarse class grid (off for resident batches: one launch set, no chunk overlap)
  bool finalize = false;       // reserve the sections of the per-read cap / poorly-modelled epilogue
  std::vector<uint8_t> fine_len;  // per read length: bit `form` set = keep the exact class (bit 0 also covers the uniform-GCP form)

  int run(const std::vector<int64_t>& regions, size_t first, size_t& next) {
    ChunkPlan& P = s.plan;
    P = ChunkPlan();
    P.force_double = force_double;
    for (auto& b : s.buckets) { b.tasks.clear(); b.hs = 0; b.stage = 0; b.cls_mask = 0; }
    s.order.clear();
    s.ukeys.clear();
    s.hap_order.clear();
    s.rlayout.clear();
    s.genlist.clear();
    s.gen_flags.clear();
    uint32_t gen64_cap = 0, gen_maxlh = 0;
    bool any_n = false;  // some haplotype contains an N: the prior table needs its sixth symbol row
    bool hap_other[256] = {false};  // haplotype bytes outside ACGTN seen in this chunk (GKL compares raw bytes: they are legal input)
    bool any_other = false;
    // Latency policy for under-filled chunks: if even with one task per (4 reads x 1 haplotype) the chunk
    // cannot fill the one-warp CTA slots of the device, the call is latency bound: the time is the serial
    // chain of one task.  Then give every read the widest lane group (shortest chain per column) that still
    // yields enough tasks, and one haplotype per task.
    int min_G = 0;
    {
      uint64_t reads_x_haps = 0;
      for (size_t kk = first; kk < regions.size(); ++kk) {
        int32_t nr = 0, nh = 0;
        in.shape(regions[kk], nr, nh);
        reads_x_haps += (uint64_t)std::max(0, nr) * (uint64_t)std::max(0, nh);
        if (reads_x_haps > (uint64_t)sm_count * 64) break;
      }
      // tasks = reads_x_haps / (32 / G) must fit one wave of CTAs, else the throughput policy is better
      const uint64_t slots = (uint64_t)sm_count * 10;
      if (reads_x_haps <= slots) min_G = 32;
      else if (reads_x_haps / 2 <= slots) min_G = 16;
      static const int force_min_g = (int)env_i64("FCS_PHMM_FORCE_MIN_G", -1);  // developer knob
      if (force_min_g >= 0) min_G = force_min_g;
    }
    // Tail shaping works on the pairs that are still to come after a region (suffix sums over this
    // chunk's regions): tasks are launched longest first, so what is cut finer here ends up in the last
    // wave of CTAs.  One "wave" is about sm_count x 64 pairs (8 CTAs x 8 reads x 1 haplotype).
    std::vector<uint64_t> pairs_after(regions.size() + 1, 0);
    std::vector<size_t> qual_off(regions.size() + 1, 0);  // first read of each region in all_gcp / all_ukey
    for (size_t kk = regions.size(); kk-- > first;) {
      int32_t nr = 0, nh = 0;
      in.shape(regions[kk], nr, nh);
      pairs_after[kk] = pairs_after[kk + 1] + (uint64_t)std::max(0, nr) * (uint64_t)std::max(0, nh);
    }
    // Which reads have constant qualities: gap continuation only (uniform-GCP form) or insertion and
    // deletion as well (all-uniform form).  The all-uniform kernels are used only if they carry the bulk
    // of the chunk: a thin extra launch spread over many (G, R) classes next to the main one costs more
    // (instruction cache, 8-CTA/SM footprint) than its faster loop gains (measured on config 3: -7 %).
    std::vector<int> all_gcp, all_ukey;
    std::vector<uint32_t> all_layout;  // blob layout flags per read (phmm_types.h read_layout)
    bool ua_chunk = false;
    // reads per fine class (keyed by the general-form class of the length: the forms share the (G, R) grid
    // closely enough for a popularity test)
    std::vector<uint32_t> len_hist(1025, 0);
    {
      ProfScope ps(&t_prof.scan);
      uint64_t n_elig = 0, n_tot = 0;
      for (size_t kk = first; kk < regions.size(); ++kk) {
        int32_t nr = 0, nh = 0;
        in.shape(regions[kk], nr, nh);
        qual_off[kk] = all_gcp.size();
        if (kk + 1 < regions.size()) prefetch_region(in, regions[kk + 1], 0x1cu, false);  // ins, del, gcp
        if (nr <= 0 || nh <= 0) continue;
        for (int32_t i = 0; i < nr; ++i) {
          const InRead r = in.read(regions[kk], i);
          int gq = -1, uk = -1;
          uint32_t lay = 0;
          if (r.len > 0 && r.i && r.d && r.c) {
            const QualScan qs = scan_quals(r.c, r.i, r.d, r.len);
            gq = qs.gcp;
            // (equal insertion and deletion quality, as GATK writes them: the kernels share M * pMX = M * pMY)
            if (gq >= 0 && qs.ins >= 0 && qs.del == qs.ins) uk = gq | (qs.ins << 8) | (qs.del << 16);
            // compact blob layouts: constant qualities travel in a 16-byte trailer, a deletion plane equal to the insertion
            // plane (GATK writes both from one value) is not copied at all
            if (two_plane_enabled() && gq >= 0) lay = uk >= 0 ? kTwoPlaneBit : (kNoGcpPlaneBit | (qs.same_indel ? kSameIndelBit : 0u));
          }
          all_gcp.push_back(gq);
          all_ukey.push_back(uk);
          all_layout.push_back(lay);
          n_elig += uk >= 0;
          if (r.len >= 1 && r.len <= 1024) ++len_hist[(size_t)r.len];
        }
        n_tot += (uint64_t)nr;
      }
      static const bool ua_enabled = env_i64("FCS_PHMM_NO_UA", 0) == 0;  // developer knob: disable the all-uniform kernels
      ua_chunk = ua_enabled && n_elig * 2 >= n_tot && n_tot > 0;
      // A read length keeps its exact ("fine") class only if that class is popular in this chunk; the
      // other lengths round up to the coarse grid of rows per lane (phmm_registry.cpp: fewer distinct loop
      // bodies in flight).  Popular = at least 1/8 of the chunk's reads fall into the class.
      static const bool coarse_enabled = env_i64("FCS_PHMM_NO_COARSE", 0) == 0;  // developer knob
      static const uint64_t pop_div = (uint64_t)env_i64("FCS_PHMM_POP_DIV", 8);  // developer knob: popular = 1/pop_div of the reads
      fine_len.assign(1025, (coarse_classes && coarse_enabled) ? 0 : 0xff);
      if (coarse_classes && coarse_enabled) {
        for (int form = 0; form < 3; form += 2) {  // general/uniform-GCP grid, all-uniform grid
          std::vector<std::pair<const ClassRef*, uint64_t>> pop;
          for (int len = 1; len <= 1024; ++len) {
            if (!len_hist[(size_t)len]) continue;
            const ClassRef* kf = f32_class_of_len(form, len);
            if (!kf) continue;
            bool found = false;
            for (auto& pr : pop)
              if (pr.first == kf) { pr.second += len_hist[(size_t)len]; found = true; break; }
            if (!found) pop.emplace_back(kf, len_hist[(size_t)len]);
          }
          for (int len = 1; len <= 1024; ++len) {
            if (!len_hist[(size_t)len]) continue;
            const ClassRef* kf = f32_class_of_len(form, len);
            for (const auto& pr : pop)
              if (pr.first == kf && pr.second * pop_div >= n_tot) fine_len[(size_t)len] |= (uint8_t)(1u << form);
          }
        }
      }
    }
    static const double tail1_x = (double)env_i64("FCS_PHMM_TAIL1_X100", 150) / 100.0;  // in waves: half-length tasks
    static const double tail2_x = (double)env_i64("FCS_PHMM_TAIL2_X100", 25) / 100.0;   // in waves: wide lane groups, one haplotype (0.5 before the pair kernels: their tasks are twice as long, the scalar wide tasks cost more against them; config 1 +4 %)
    static const int tail2_g = (int)env_i64("FCS_PHMM_TAIL2_G", 16);
    const uint64_t wave_pairs = (uint64_t)sm_count * 64u;
    std::vector<uint32_t> f64_cap((size_t)f64_queue_count(), 0), f64_maxlh((size_t)f64_queue_count(), 0);
    std::vector<int> gcps, ukeys;
    std::vector<uint32_t> layouts;
    size_t last_bucket = 0;  // the bucket the previous task went to (nearly always the next one's too)
    std::vector<uint8_t> paired;  // per read of the region (sorted order): its even haplotypes ran on the haplotype-pair kernels
    std::vector<uint32_t> hap_len_chunk;  // by chunk-wide haplotype index
    int chunk_gcp = -2;  // -2: nothing seen yet, -1: mixed, >= 0: the one value every read shares
    size_t reads_bytes = 0, haps_bytes = 0;
    std::vector<uint32_t> lens, hlens, hord, hl_tmp;
    std::vector<uint64_t> sort_keys;
    size_t k = first;
    for (; k < regions.size(); ++k) {
      const int64_t g = regions[k];
      int32_t nr = 0, nh = 0;
      in.shape(g, nr, nh);
      if (nr < 0 || nh < 0) return set_error(FCS_PHMM_EINVAL, "negative read or haplotype count");
      if (nr == 0 || nh == 0) {  // nothing to compute; keep the slot so scatter stays aligned
        P.regions.push_back(g);
        P.reg_out0.push_back(P.n_pairs);
        continue;
      }
      if (!in.out(g)) return set_error(FCS_PHMM_EINVAL, "out_log10 is null");
      lens.resize(nr);
      gcps.resize(nr);
      ukeys.resize(nr);
      layouts.resize(nr);
      hlens.resize(nh);
      uint64_t sum_r = 0, sum_h = 0;
      size_t rb = 0, hb = 0;
      const double tp0 = g_plan_prof ? now_ms() : 0;
      for (int32_t i = 0; i < nr; ++i) {
        const InRead r = in.read(g, i);
        if (r.len <= 0 || !r.b || !r.q || !r.i || !r.d || !r.c)
          return set_error(FCS_PHMM_EINVAL, "read with non-positive length or null array");
        if (r.len > FCS_PHMM_MAX_READ_LEN) return set_error(FCS_PHMM_EUNSUPPORTED, "read longer than FCS_PHMM_MAX_READ_LEN");
        lens[i] = (uint32_t)r.len;
        gcps[i] = all_gcp[qual_off[k] + (size_t)i];
        ukeys[i] = all_ukey[qual_off[k] + (size_t)i];
        layouts[i] = all_layout[qual_off[k] + (size_t)i];
        sum_r += (uint64_t)r.len;
        rb += read_blob_bytes((uint32_t)r.len, layouts[i]);
      }
      for (int32_t j = 0; j < nh; ++j) {
        const InHap h = in.hap(g, j);
        if (h.len <= 0 || !h.b) return set_error(FCS_PHMM_EINVAL, "haplotype with non-positive length or null array");
        if (h.len > FCS_PHMM_MAX_HAP_LEN) return set_error(FCS_PHMM_EUNSUPPORTED, "haplotype longer than FCS_PHMM_MAX_HAP_LEN");
        hlens[j] = (uint32_t)h.len;
        sum_h += (uint64_t)h.len;
        {
          const uint32_t seen = hap_classes(h.b, h.len);
          any_n = any_n || (seen & 1u);
          if (seen & 2u) {  // rare: collect the foreign byte values
            const uint8_t* cls = hap_byte_class();
            any_other = true;
            for (int32_t x = 0; x < h.len; ++x)
              if (cls[h.b[x]] & 2u) hap_other[h.b[x]] = true;
          }
        }
        hb += round_up16((uint32_t)h.len);
      }
      // haplotypes longest first (stable): neighbours -- the pairs of the haplotype-pair kernels -- then differ least in
      // length, and an odd one out is the shortest.  HapMeta::col maps a packed haplotype back to the caller's column.
      hord.resize((size_t)nh);
      std::iota(hord.begin(), hord.end(), 0u);
      {
        bool sorted = true;
        for (int32_t j = 1; j < nh && sorted; ++j) sorted = hlens[j] <= hlens[j - 1];
        if (!sorted) {
          std::stable_sort(hord.begin(), hord.end(), [&](uint32_t a, uint32_t b) { return hlens[a] > hlens[b]; });
          hl_tmp.assign(hlens.begin(), hlens.begin() + nh);
          for (int32_t j = 0; j < nh; ++j) hlens[j] = hl_tmp[hord[(size_t)j]];
        }
      }
      const uint64_t cells = sum_r * sum_h;
      const uint64_t pairs = (uint64_t)nr * (uint64_t)nh;
      if (!P.regions.empty() && P.cells > 0 &&
          ((int64_t)(P.cells + cells) > max_cells || P.n_pairs + pairs > 0x7fffffffULL ||
           reads_bytes + rb + haps_bytes + hb > (size_t)1 << 31))
        break;
      if (pairs > 0x7fffffffULL) return set_error(FCS_PHMM_EUNSUPPORTED, "region with more than 2^31 pairs");
      const double tp1 = g_plan_prof ? now_ms() : 0;
      // ---- reads sorted by length (descending, stable) so a lane group shares a class
      const size_t ord0 = s.order.size();
      s.order.resize(ord0 + nr);
      uint32_t* ord = s.order.data() + ord0;
      // Descending length, ties in the caller's order.  Nearly every read of a region has the region's (maximum) length:
      // those keep their order up front, only the few shorter (clipped) ones are sorted -- one packed key per read, plain
      // sort (std::stable_sort allocates a scratch buffer per call, and this runs once per region).
      uint32_t maxlen = 0;
      for (int32_t i = 0; i < nr; ++i) maxlen = std::max(maxlen, lens[i]);
      sort_keys.clear();
      int32_t n_top = 0;
      for (int32_t i = 0; i < nr; ++i) {
        if (lens[i] == maxlen) ord[n_top++] = (uint32_t)i;
        else sort_keys.push_back(((uint64_t)(0xffffffffu - lens[i]) << 32) | (uint32_t)i);
      }
      if (!sort_keys.empty()) {
        std::sort(sort_keys.begin(), sort_keys.end());
        for (size_t x = 0; x < sort_keys.size(); ++x) ord[n_top + (int32_t)x] = (uint32_t)(sort_keys[x] & 0xffffffffu);
      }
      const double tp2 = g_plan_prof ? now_ms() : 0;
      // ---- tasks
      // Tail shaping (equal-size tasks finish in lock step otherwise): the reads within the last ~1.5
      // waves of the chunk are cut into tasks of half the haplotype columns, those within the last ~0.5
      // wave into one-haplotype tasks on wide lane groups (a quarter of the run time of a regular task).
      const bool use_ua = ua_chunk;
      const uint32_t read_base = (uint32_t)P.n_reads, hap_base = (uint32_t)P.n_haps;
      uint32_t maxlh = 0;
      for (int32_t j = 0; j < nh; ++j) maxlh = std::max(maxlh, hlens[j]);
      const bool long_hap = maxlh >= (uint32_t)kGenericMinHapLen;
      s.gen_flags.resize(read_base + (size_t)nr, 0);
      paired.assign((size_t)nr, 0);
      for (int32_t i = 0; i < nr;) {
        const double pairs_left = (double)(pairs_after[k + 1] + (uint64_t)(nr - i) * (uint64_t)nh);
        const bool tail1 = shape_tail && pairs_left <= tail1_x * (double)wave_pairs;
        const int wide_G = min_G ? min_G : ((shape_tail && pairs_left <= tail2_x * (double)wave_pairs) ? tail2_g : 0);
        const uint32_t cols_limit = wide_G ? 1u : (tail1 ? std::max(hs_cols / 2, 1u) : hs_cols);
        if (long_hap || lens[ord[i]] > (uint32_t)kGenericMaxSinglePassRead) {
          // striped generic path: one (read, hap) pair per list entry
          s.gen_flags[read_base + (size_t)i] = 1;
          for (int32_t j = 0; j < nh; ++j) s.genlist.push_back(RerunEntry{read_base + (uint32_t)i, hap_base + (uint32_t)j});
          gen64_cap += (uint32_t)nh;
          gen_maxlh = std::max(gen_maxlh, maxlh);
          ++i;
          continue;
        }
        // full groups of the longest remaining read use the table; the last, partly filled group of a
        // region asks for the class that is cheapest per read actually served
        const int len0 = (int)lens[ord[i]];
        const bool fine0 = len0 > 1024 || (fine_len[(size_t)len0] & 1u), fine2 = len0 > 1024 || (fine_len[(size_t)len0] & 4u);
        const ClassRef* k0 = fine0 ? f32_class_of_len(0, len0) : f32_coarse_class_of_len(0, len0);
        if (wide_G) k0 = select_class_wide(false, false, (int)lens[ord[i]], wide_G);
        else if (nr - i < 32 / k0->G) k0 = select_class_for(false, false, len0, nr - i, (int)(sum_h / (uint64_t)nh), !fine0);
        // All-uniform form: a full warp of reads that share one (continuation, insertion, deletion)
        // quality triple.  Throughput policy only; leftover groups and latency-bound calls keep the
        // wide-group classes of the other forms.
        const ClassRef* ku = (!wide_G && use_ua && ukeys[ord[i]] >= 0) ? (fine2 ? f32_class_of_len(2, len0) : f32_coarse_class_of_len(2, len0)) : nullptr;
        if (ku) {
          // fewer reads left than the class has lane groups: the all-uniform class that is cheapest per read
          // served, if the leftover fills every group of it (e.g. 4 reads -> G=8); else the other forms' wide classes
          if (nr - i < 32 / ku->G) {
            const TierKernel* main_tk = ku->tk;
            ku = select_class_for(false, 2, len0, nr - i, (int)(sum_h / (uint64_t)nh), !fine2);
            if (ku && nr - i < 32 / ku->G) ku = nullptr;
            // a slightly taller class that lives in the kernel of the region's main class saves a launch (and
            // its fork/join: ~25 us of driver calls per chunk) for a row or two of padding
            for (int up = 0; ku && ku->tk != main_tk && up <= 3; ++up)
              if (const ClassRef* alt = find_class(false, 2, ku->G, ku->R + up))
                if (alt->tk == main_tk) ku = alt;
          }
          const int ngu = ku ? 32 / ku->G : 0;
          for (int32_t x = 1; ku && x < ngu; ++x)
            if (ukeys[ord[i + x]] != ukeys[ord[i]]) ku = nullptr;
        }
        if (ku) k0 = ku;
        // Haplotype-pair form: reads with one gap-continuation quality against two haplotypes at a time in packed
        // f32x2 arithmetic.  Throughput policy only (pair classes on wide lane groups for the tail window were measured:
        // config 3 -1.4 %, config 2 -1 %); an odd haplotype is left to the scalar uniform-GCP class.
        static const bool pairs_enabled = env_i64("FCS_PHMM_NO_PAIRS", 0) == 0;  // developer knob
        const ClassRef* kp = nullptr;
        if (!ku && !wide_G && pairs_enabled && nh >= 2 && gcps[ord[i]] >= 0) {
          kp = fine0 ? f32_class_of_len(3, len0) : f32_coarse_class_of_len(3, len0);
          if (kp && nr - i < 32 / kp->G) kp = select_class_for(false, 3, len0, nr - i, (int)(sum_h / (uint64_t)nh), !fine0);
          const int cp = kp ? std::min<int32_t>(32 / kp->G, nr - i) : 0;
          for (int32_t x = 1; kp && x < cp; ++x)
            if (gcps[ord[i + x]] != gcps[ord[i]]) kp = nullptr;
        }
        const int NG = 32 / (kp ? kp->G : k0->G);
        const int cnt = std::min<int32_t>(NG, nr - i);
        // tasks of reads [i + r0, i + r0 + rc) x haplotypes [j0, j1) for class kc (pairs: two haplotypes per wavefront step)
        auto emit = [&](const ClassRef* kc, int tg, int32_t r0, int32_t rc, int32_t j0, int32_t j1, bool pairs) {
          TaskBucket* bk = nullptr;
          if (last_bucket < s.buckets.size() && s.buckets[last_bucket].tk == kc->tk && s.buckets[last_bucket].gcp == tg) bk = &s.buckets[last_bucket];
          for (size_t bi = 0; !bk && bi < s.buckets.size(); ++bi)
            if (s.buckets[bi].tk == kc->tk && s.buckets[bi].gcp == tg) { bk = &s.buckets[bi]; last_bucket = bi; }
          if (!bk) {
            s.buckets.emplace_back();
            bk = &s.buckets.back();
            bk->tk = kc->tk;
            bk->gcp = tg;
            last_bucket = s.buckets.size() - 1;
          }
          const int32_t hstep = pairs ? 2 : 1;
          for (int32_t j = j0; j < j1;) {
            uint32_t cols = 0, stage = 0;
            int32_t jn = j;
            while (jn < j1 && (jn - j) < 0xfffe) {
              const uint32_t w = (pairs && jn + 1 < j1) ? std::max(hlens[jn], hlens[jn + 1]) : hlens[jn];
              if (jn != j && cols + w + (uint32_t)(kc->G - 1) > cols_limit) break;
              cols += w + (uint32_t)(kc->G - 1);
              for (int32_t q = jn; q < std::min(jn + hstep, j1); ++q) stage += round_up16(hlens[q]);
              jn = std::min(jn + hstep, j1);
            }
            cols += (uint32_t)(kc->G - 1);
            Task t;
            t.read0 = read_base + (uint32_t)(i + r0);
            t.hap0 = hap_base + (uint32_t)j;
            t.n_reads = (uint16_t)rc;
            t.n_haps = (uint16_t)(jn - j);
            t.cls = (uint32_t)kc->cls;
            bk->tasks.push_back(t);
            bk->hs = std::max(bk->hs, pairs ? 2u * cols : cols);  // in 16-bit units: a pair entry holds two table offsets
            bk->stage = std::max(bk->stage, stage);
            bk->cls_mask |= 1ull << kc->cls;
            j = jn;
          }
        };
        if (kp) {
          emit(kp, gcps[ord[i]], 0, cnt, 0, nh & ~1, true);
          for (int32_t x = 0; x < cnt; ++x) paired[(size_t)(i + x)] = 1;  // an odd last haplotype is handled below
        } else {
          int tg = gcps[ord[i]];  // uniform-GCP form only if every read of the task shares the value
          for (int32_t x = 1; x < cnt; ++x)
            if (gcps[ord[i + x]] != tg) tg = -1;
          const ClassRef* kc = ku ? ku : ((tg >= 0 && k0->twin) ? k0->twin : k0);
          if (ku) tg = ukeys[ord[i]];
          emit(kc, tg, 0, cnt, 0, nh, false);
        }
        i += cnt;
      }
      if (nh & 1) {
        // The odd last haplotype of the reads that ran on the pair kernels: scalar uniform-GCP classes, grouped on
        // their own (a scalar class may seat more reads per warp than the pair class did).
        for (int32_t a = 0; a < nr;) {
          if (!paired[(size_t)a]) { ++a; continue; }
          int32_t b = a;
          while (b < nr && paired[(size_t)b]) ++b;
          for (int32_t i = a; i < b;) {
            const int len0 = (int)lens[ord[i]];
            const bool fine0 = len0 > 1024 || (fine_len[(size_t)len0] & 1u);
            const ClassRef* k0 = fine0 ? f32_class_of_len(0, len0) : f32_coarse_class_of_len(0, len0);
            if (b - i < 32 / k0->G) k0 = select_class_for(false, false, len0, b - i, (int)hlens[nh - 1], !fine0);
            const int cnt = std::min<int32_t>(32 / k0->G, b - i);
            int tg = gcps[ord[i]];
            for (int32_t x = 1; x < cnt; ++x)
              if (gcps[ord[i + x]] != tg) tg = -1;
            const ClassRef* kc = (tg >= 0 && k0->twin) ? k0->twin : k0;
            TaskBucket* bk = nullptr;
            for (auto& bb : s.buckets)
              if (bb.tk == kc->tk && bb.gcp == tg) { bk = &bb; break; }
            if (!bk) {
              s.buckets.emplace_back();
              bk = &s.buckets.back();
              bk->tk = kc->tk;
              bk->gcp = tg;
            }
            Task t;
            t.read0 = read_base + (uint32_t)i;
            t.hap0 = hap_base + (uint32_t)(nh - 1);
            t.n_reads = (uint16_t)cnt;
            t.n_haps = 1;
            t.cls = (uint32_t)kc->cls;
            bk->tasks.push_back(t);
            bk->hs = std::max(bk->hs, hlens[nh - 1] + 2u * (uint32_t)(kc->G - 1));
            bk->stage = std::max(bk->stage, round_up16(hlens[nh - 1]));
            bk->cls_mask |= 1ull << kc->cls;
            i += cnt;
          }
          a = b;
        }
      }
      const double tp3 = g_plan_prof ? now_ms() : 0;
      // ---- FP64 queue capacity per class (worst case: every pair of the read falls back)
      for (int32_t i = 0; i < nr; ++i) {
        if (long_hap || lens[i] > (uint32_t)kGenericMaxSinglePassRead) continue;  // counted in gen64_cap
        int c64 = qid_of_len((int)lens[i]);
        if (min_G) {
          const ClassRef* kw = select_class_wide(true, false, (int)lens[i], 32);
          c64 = f64_queue_id(kw->G, kw->R);
        }
        f64_cap[c64] += (uint32_t)nh;
        f64_maxlh[c64] = std::max(f64_maxlh[c64], maxlh);
        chunk_gcp = (chunk_gcp == -2) ? gcps[i] : (chunk_gcp == gcps[i] ? chunk_gcp : -1);
      }
      hap_len_chunk.insert(hap_len_chunk.end(), hlens.begin(), hlens.end());
      s.ukeys.insert(s.ukeys.end(), ukeys.begin(), ukeys.begin() + nr);
      s.hap_order.insert(s.hap_order.end(), hord.begin(), hord.begin() + nh);
      s.rlayout.insert(s.rlayout.end(), layouts.begin(), layouts.begin() + nr);
      P.regions.push_back(g);
      P.reg_out0.push_back(P.n_pairs);
      P.n_reads += nr;
      P.n_haps += nh;
      P.n_pairs += pairs;
      P.cells += cells;
      reads_bytes += rb;
      haps_bytes += hb;
      if (g_plan_prof) {
        const double tp4 = now_ms();
        t_prof.region += tp1 - tp0; t_prof.sort += tp2 - tp1; t_prof.tasks += tp3 - tp2; t_prof.f64 += tp4 - tp3;
      }
    }
    next = k;
    ProfScope ps_post(&t_prof.post);
    // ---- layout
    size_t off = 0;
    P.latency_mode = min_G != 0;
    P.n_sym = any_n ? 6u : 5u;
    P.extra_bytes = 0;
    if (any_other) {
      // Haplotype bytes outside ACGTN match a read base only if the read holds the same byte (or an N).  The ones
      // no read of the chunk contains share one symbol row; each of the others needs its own.
      bool shared[256] = {false};
      uint32_t n_extra = 0;
      for (int64_t g : P.regions) {
        int32_t nr = 0, nh = 0;
        in.shape(g, nr, nh);
        if (nr <= 0 || nh <= 0) continue;
        for (int32_t i = 0; i < nr; ++i) {
          const InRead r = in.read(g, i);
          for (int32_t x = 0; x < r.len; ++x)
            if (hap_other[r.b[x]] && !shared[r.b[x]]) {
              shared[r.b[x]] = true;
              if (n_extra >= (uint32_t)kMaxExtraSyms)
                return set_error(FCS_PHMM_EUNSUPPORTED, "more than 8 distinct byte values outside ACGTN occur in both reads and haplotypes of one chunk");
              P.extra_bytes |= (uint64_t)r.b[x] << (8u * n_extra);
              ++n_extra;
            }
        }
      }
      P.n_sym = (uint32_t)kCodeExtra0 + n_extra;
    }
    P.off_reads = off; off = align_up(off + reads_bytes, 256);
    P.off_haps = off; off = align_up(off + haps_bytes, 256);
    P.off_rmeta = off; off = align_up(off + P.n_reads * sizeof(ReadMeta), 256);
    P.off_hmeta = off; off = align_up(off + P.n_haps * sizeof(HapMeta), 256);
    P.finalize = finalize;
    P.off_rnh = off;
    if (finalize) off = align_up(off + P.n_reads * sizeof(uint32_t), 256);
    P.off_tasks = off;
    P.n_tasks = 0;
    for (size_t bi = 0; bi < s.buckets.size(); ++bi) {
      TaskBucket& b = s.buckets[bi];
      if (b.tasks.empty()) continue;
      // Order: class by class (rows per lane descending), longest tasks first inside a class.  CTAs
      // that are resident together then run the same class, i.e. the same unrolled loop body; mixing
      // classes freely (pure longest-first) thrashes the instruction cache (C3: 2.5 -> 0.9 TCUPS).
      {
        const uint32_t n = (uint32_t)b.tasks.size();
        // one integer key per task: class rank (rows per lane desc, lanes desc) in the high half, inverted cost below
        std::vector<uint32_t> rank((size_t)b.tk->n_classes);
        {
          std::vector<int> byrank((size_t)b.tk->n_classes);
          std::iota(byrank.begin(), byrank.end(), 0);
          std::sort(byrank.begin(), byrank.end(), [&](int x, int y) {
            const ClassDesc &cx = b.tk->classes[x], &cy = b.tk->classes[y];
            return cx.R != cy.R ? cx.R > cy.R : cx.G > cy.G; });
          for (size_t r2 = 0; r2 < byrank.size(); ++r2) rank[(size_t)byrank[r2]] = (uint32_t)r2;
        }
        std::vector<uint32_t> cost(n);
        b.max_task_cost = 0;
        for (uint32_t t = 0; t < n; ++t) {
          const Task& x = b.tasks[t];
          const ClassDesc& cd = b.tk->classes[x.cls];
          uint32_t cols = 0;
          if (b.tk->form == 3) {
            for (uint32_t j = 0; j < x.n_haps; j += 2)
              cols += std::max(hap_len_chunk[x.hap0 + j], j + 1 < x.n_haps ? hap_len_chunk[x.hap0 + j + 1] : 0u) + (uint32_t)cd.G - 1u;
          } else {
            for (uint32_t j = 0; j < x.n_haps; ++j) cols += hap_len_chunk[x.hap0 + j] + (uint32_t)cd.G - 1u;
          }
          cost[t] = (uint32_t)cd.R * cols;
          b.max_task_cost = std::max(b.max_task_cost, cost[t]);
        }
        // stable counting sort on (class rank, cost quantised to 256 levels, descending): O(n); the order
        // inside a level is planning order -- longest-first is a heuristic, exact ties do not matter
        const uint32_t nb = (uint32_t)b.tk->n_classes * 256u;
        std::vector<uint32_t> start(nb + 1, 0), slot(n);
        const uint64_t mc = std::max<uint32_t>(b.max_task_cost, 1u);
        bool ordered = true;
        uint32_t prev = 0;
        for (uint32_t t = 0; t < n; ++t) {
          const uint32_t q = 255u - (uint32_t)((uint64_t)cost[t] * 255u / mc);
          slot[t] = rank[b.tasks[t].cls] * 256u + q;
          ordered = ordered && slot[t] >= prev;
          prev = slot[t];
          ++start[slot[t] + 1];
        }
        if (!ordered) {
          for (uint32_t i = 0; i < nb; ++i) start[i + 1] += start[i];
          std::vector<Task> sorted(n);
          for (uint32_t t = 0; t < n; ++t) sorted[start[slot[t]]++] = b.tasks[t];
          b.tasks.swap(sorted);
        }
      }
      F32Range r;
      r.tk = b.tk;
      r.gcp = b.tk->form ? b.gcp : -1;
      r.bucket = (uint32_t)bi;
      r.task0 = (uint32_t)P.n_tasks;
      r.n_tasks = (uint32_t)b.tasks.size();
      r.hs_cap = b.hs;
      r.hap_stage = b.stage;
      r.max_task_cost = b.max_task_cost;
      r.smem = 0;
      for (int c = 0; c < b.tk->n_classes; ++c)
        if (b.cls_mask >> c & 1ull) r.smem = std::max(r.smem, b.tk->classes[c].smem_bytes(b.hs, b.stage, P.n_sym));
      P.f32.push_back(r);
      P.n_tasks += r.n_tasks;
    }
    off = align_up(off + P.n_tasks * sizeof(Task), 256);
    P.off_genlist = off;
    P.n_gen = (uint32_t)s.genlist.size();
    P.gen64_cap = gen64_cap;
    off = align_up(off + (size_t)P.n_gen * sizeof(RerunEntry), 256);
    P.off_rbase = off; off += kMaxF64Classes * sizeof(uint32_t);
    P.off_rcount = off; off += kMaxF64Classes * sizeof(uint32_t);
    off = align_up(off, 256);
    P.off_rerun = off;
    P.f64_gcp = chunk_gcp >= 0 ? chunk_gcp : -1;
    P.f64.clear();
    // FP64 queues (one per general-form FP64 class) and the launches that drain them: queues whose
    // class lives in the same tier kernel share a launch.
    P.queues.clear();
    for (size_t q = 0; q < f64_cap.size(); ++q) {
      if (!f64_cap[q]) continue;
      F64Queue qu;
      qu.qid = (uint32_t)q;
      qu.cap = f64_cap[q];
      qu.maxlh = f64_maxlh[q];
      P.queues.push_back(qu);
      // two candidate drains per queue: the throughput class of the queue and, for a short queue, the
      // widest class covering the same reads (the FP64 phase of a mostly-FP32 batch is a handful of
      // pairs whose serial chain is the whole cost)
      const ClassRef* kt = f64_queue_class((int)q, P.f64_gcp >= 0);
      const ClassRef* kw = select_class_wide(true, P.f64_gcp >= 0, kt->G * kt->R - 1, 32);
      const uint32_t short_q = (uint32_t)sm_count * 2u;
      const bool two = !force_double && kw && !(kw->G == kt->G && kw->R == kt->R);
      for (int pass = 0; pass < (two ? 2 : 1); ++pass) {
        const ClassRef* k = pass == 0 ? kt : kw;
        F64Range* rr = nullptr;
        for (auto& r : P.f64)
          if (r.tk == k->tk) { rr = &r; break; }
        if (!rr) {
          P.f64.emplace_back();
          rr = &P.f64.back();
          rr->tk = k->tk;
          rr->n_seg = 0;
          rr->hs_cap = 0;
          rr->hap_stage = 0;
          rr->smem = 0;
        }
        if (rr->n_seg >= 32) return set_error(FCS_PHMM_EINVAL, "internal: more than 32 FP64 segments in one launch");
        rr->seg_cls[rr->n_seg] = (uint16_t)k->cls;
        rr->seg_qid[rr->n_seg] = (uint16_t)q;
        rr->seg_G[rr->n_seg] = (uint16_t)k->G;
        if (!two) { rr->seg_min[rr->n_seg] = 0; rr->seg_max[rr->n_seg] = 0xffffffffu; rr->seg_cap[rr->n_seg] = f64_cap[q]; }
        else if (pass == 0) { rr->seg_min[rr->n_seg] = short_q + 1; rr->seg_max[rr->n_seg] = 0xffffffffu; rr->seg_cap[rr->n_seg] = f64_cap[q]; }
        else { rr->seg_min[rr->n_seg] = 0; rr->seg_max[rr->n_seg] = short_q; rr->seg_cap[rr->n_seg] = std::min(f64_cap[q], short_q); }
        rr->n_seg++;
        rr->hap_stage = std::max(rr->hap_stage, round_up16(f64_maxlh[q]));
      }
    }
    for (auto& r : P.f64) {
      // hs_cap / hap_stage are per lane group; use each class's own G for the stream padding bound
      uint32_t hs = 0;
      for (uint32_t k = 0; k < r.n_seg; ++k) {
        uint32_t mlh = 0;
        for (const auto& qu : P.queues)
          if (qu.qid == r.seg_qid[k]) mlh = qu.maxlh;
        hs = std::max(hs, mlh + 2u * (uint32_t)(r.seg_G[k] - 1));
      }
      r.hs_cap = hs;
      for (uint32_t k = 0; k < r.n_seg; ++k) r.smem = std::max(r.smem, r.tk->classes[r.seg_cls[k]].smem_bytes(r.hs_cap, r.hap_stage, P.n_sym));
    }
    const size_t rerun_bytes = align_up(P.n_pairs * sizeof(RerunEntry), 256);
    if (force_double) { off += rerun_bytes; P.in_bytes = off; }
    else { P.in_bytes = off; off += rerun_bytes; }
    // per-CTA boundary rows of the striped kernels (3 planes of up to 8-byte values per CTA)
    P.gen_ctas = 0;
    P.scratch_cols = 0;
    P.off_scratch = off;
    if (P.n_gen) {
      P.scratch_cols = (gen_maxlh + 31u) / 32u * 32u + 32u;
      const size_t per_cta = (size_t)3 * P.scratch_cols * sizeof(double);
      size_t ctas = std::min<size_t>((size_t)sm_count * 4, std::max<size_t>(1, ((size_t)256 << 20) / per_cta));
      ctas = std::min<size_t>(ctas, std::max<uint32_t>(P.n_gen, 1u));
      P.gen_ctas = (uint32_t)ctas;
      off = align_up(off + ctas * per_cta, 256);
    }
    P.off_out = off; off += P.n_pairs * sizeof(double);
    P.off_raw = off; if (keep_raw) off += P.n_pairs * sizeof(float);
    P.off_used = off; off += P.n_pairs;
    P.off_poor = off; if (finalize) off += P.n_reads;
    P.total_bytes = align_up(off, 256);
    return FCS_PHMM_OK;
  }
};

}  // namespace


// Pass 2: copy reads / quals / haplotypes into the pinned staging buffer in device layout.
int Engine::pack_chunk(Slot& s, const Input& in) { return pack_chunk_static(s, in); }

int Engine::pack_chunk_static(Slot& s, const Input& in) {
  ChunkPlan& P = s.plan;
  uint8_t* base = s.h_in;
  ReadMeta* rmeta = reinterpret_cast<ReadMeta*>(base + P.off_rmeta);
  HapMeta* hmeta = reinterpret_cast<HapMeta*>(base + P.off_hmeta);
  uint32_t* rbase = reinterpret_cast<uint32_t*>(base + P.off_rbase);
  uint32_t* rcount = reinterpret_cast<uint32_t*>(base + P.off_rcount);
  RerunEntry* rerun = P.force_double ? reinterpret_cast<RerunEntry*>(base + P.off_rerun) : nullptr;
  std::memset(rbase, 0, kMaxF64Classes * sizeof(uint32_t));
  std::memset(rcount, 0, kMaxF64Classes * sizeof(uint32_t));
  {
    uint32_t acc = 0;
    for (const F64Queue& q : P.queues) { rbase[q.qid] = acc; acc += q.cap; }
    rbase[kQueueGenericF64] = acc;
  }
  std::vector<uint32_t> fill(kMaxF64Classes, 0);
  size_t rpos = 0, hpos = 0, ridx = 0, hidx = 0, opos = 0;
  for (size_t k = 0; k < P.regions.size(); ++k) {
    const int64_t g = P.regions[k];
    int32_t nr = 0, nh = 0;
    in.shape(g, nr, nh);
    if (k + 1 < P.regions.size()) prefetch_region(in, P.regions[k + 1], 0x1fu, true);  // every plane that may be copied + haplotypes
    if (nr == 0 || nh == 0) continue;
    const uint32_t hap0 = (uint32_t)hidx;
    const uint32_t* hord = s.hap_order.data() + hidx;  // packed position -> caller's haplotype index (longest first)
    for (int32_t j = 0; j < nh; ++j) {
      const uint32_t col = hord[j];
      const InHap h = in.hap(g, (int32_t)col);
      uint8_t* dst = base + P.off_haps + hpos;
      const uint32_t lp = round_up16((uint32_t)h.len);
      copy_padded16(dst, h.b, (uint32_t)h.len, (uint8_t)'N');
      hmeta[hap0 + (uint32_t)j].data_off16 = (uint32_t)(hpos / 16);
      hmeta[hap0 + (uint32_t)j].len = (uint32_t)h.len;
      hmeta[hap0 + (uint32_t)j].col = col;
      hpos += lp;
    }
    hidx += (size_t)nh;
    const uint32_t* ord = s.order.data() + opos;
    for (int32_t i = 0; i < nr; ++i) {
      const uint32_t oi = ord[i];
      const InRead r = in.read(g, (int32_t)oi);
      const uint32_t lp = round_up16((uint32_t)r.len);
      uint8_t* dst = base + P.off_reads + rpos;
      const uint8_t* src[5] = {r.b, r.q, r.i, r.d, r.c};
      // blob layout from the planner's scan: constant qualities go to a 16-byte trailer, a deletion plane that equals the
      // insertion plane is not copied
      const uint32_t layout = s.rlayout[opos + oi];
      const uint32_t npl = read_planes(layout);
      for (uint32_t pl = 0; pl < npl; ++pl) copy_padded16(dst + (size_t)pl * lp, src[pl], (uint32_t)r.len, 0);
      if (layout) {
        uint8_t* tr = dst + (size_t)npl * lp;
        std::memset(tr, 0, 16);
        tr[0] = (uint8_t)(r.i[0] & 127);  // insertion  (read by the kernels only when the plane is absent, i.e. constant)
        tr[1] = (uint8_t)(r.d[0] & 127);  // deletion
        tr[2] = (uint8_t)(r.c[0] & 127);  // gap continuation
      }
      int c64 = s.gen_flags[ridx] ? kQueueGenericF64 : qid_of_len(r.len);
      if (P.latency_mode && !s.gen_flags[ridx]) {
        const ClassRef* kw = select_class_wide(true, false, r.len, 32);
        c64 = f64_queue_id(kw->G, kw->R);
      }
      ReadMeta& m = rmeta[ridx];
      m.data_off16 = (uint32_t)(rpos / 16);
      m.len_cls = (uint32_t)r.len | layout | ((uint32_t)c64 << 24);
      m.out_off = (uint32_t)(P.reg_out0[k] + (uint64_t)oi * (uint64_t)nh);
      m.hap0 = hap0;
      if (P.finalize) reinterpret_cast<uint32_t*>(base + P.off_rnh)[ridx] = (uint32_t)nh;
      if (rerun) {
        RerunEntry* e = rerun + rbase[c64] + fill[c64];
        for (int32_t j = 0; j < nh; ++j) { e[j].read = (uint32_t)ridx; e[j].hap = hap0 + (uint32_t)j; }
        fill[c64] += (uint32_t)nh;
      }
      rpos += read_blob_bytes((uint32_t)r.len, layout);
      ++ridx;
    }
    opos += nr;
  }
  if (rerun) {
    for (const F64Queue& q : P.queues) rcount[q.qid] = fill[q.qid];
    rcount[kQueueGenericF64] = fill[kQueueGenericF64];
  }
  if (P.n_gen) std::memcpy(base + P.off_genlist, s.genlist.data(), (size_t)P.n_gen * sizeof(RerunEntry));
  Task* tasks = reinterpret_cast<Task*>(base + P.off_tasks);
  for (const F32Range& r : P.f32) {
    std::memcpy(tasks + r.task0, s.buckets[r.bucket].tasks.data(), (size_t)r.n_tasks * sizeof(Task));
  }
  return FCS_PHMM_OK;
}

void Engine::fill_kparams(const Device& d, const Slot& s, KParams& p, bool f64) const {
  const ChunkPlan& P = s.plan;
  uint8_t* b = s.d_buf;
  p.reads = b + P.off_reads;
  p.haps = b + P.off_haps;
  p.rmeta = reinterpret_cast<const ReadMeta*>(b + P.off_rmeta);
  p.hmeta = reinterpret_cast<const HapMeta*>(b + P.off_hmeta);
  p.tasks = reinterpret_cast<const Task*>(b + P.off_tasks);
  p.n_tasks = 0;
  p.ph2pr = f64 ? d.d_ph2pr_d : d.d_ph2pr_f;
  p.mm = f64 ? d.d_mm_d : d.d_mm_f;
  p.out = reinterpret_cast<double*>(b + P.off_out);
  p.used_fp64 = b + P.off_used;
  p.raw_f32 = keep_raw_ ? reinterpret_cast<float*>(b + P.off_raw) : nullptr;
  p.rerun = reinterpret_cast<RerunEntry*>(b + P.off_rerun);
  p.rerun_count = reinterpret_cast<uint32_t*>(b + P.off_rcount);
  p.rerun_base = reinterpret_cast<const uint32_t*>(b + P.off_rbase);
  p.n_seg = 0;
  p.gen_list = reinterpret_cast<const RerunEntry*>(b + P.off_genlist);
  p.gen_count = P.n_gen;
  p.scratch_cols = P.scratch_cols;
  p.scratch = b + P.off_scratch;
  p.hs_cap = 0;
  p.hap_stage_bytes = 0;
  p.n_sym = P.n_sym;
  p.extra_bytes = P.extra_bytes;
  p.c_xx_f = 0.f; p.c_gm_f = 0.f; p.c_xx_d = 0.0; p.c_gm_d = 0.0;
  p.c_mm_f = 0.f; p.c_mx_f = 0.f;
}

// launch constants of a uniform-GCP / all-uniform launch, from the same tables the kernels index.
// key = gcp | ins << 8 | del << 16 (the insertion / deletion bytes matter to the all-uniform form only)
static void set_gcp_constants(KParams& p, int key) {
  if (key < 0) return;
  const Luts& L = luts();
  const int gcp = key & 127;
  const uint32_t iq = (uint32_t)(key >> 8) & 127u, dq = (uint32_t)(key >> 16) & 127u;
  const uint32_t mn = std::min(iq, dq), mx = std::max(iq, dq);
  p.c_mm_f = L.mm_f[((mx * (mx + 1u)) >> 1) + mn];
  p.c_mx_f = L.ph2pr_f[iq];  // == ph2pr[dq]: the all-uniform key requires iq == dq
  p.c_xx_f = L.ph2pr_f[gcp & 127];
  p.c_gm_f = 1.0f - p.c_xx_f;
  p.c_xx_d = L.ph2pr_d[gcp & 127];
  p.c_gm_d = 1.0 - p.c_xx_d;
}

// Enqueue one chunk on the slot's stream.  upload/download = include the H2D / D2H copies.
int Engine::launch_chunk(Device& d, Slot& s, bool upload, bool download, bool timing) {
  const ChunkPlan& P = s.plan;
  static const bool lc_probe = env_i64("FCS_PHMM_TIMELINE", 0) >= 2;  // developer probe: host cost of the driver calls
  const double lc0 = lc_probe ? now_ms() : 0;
  double lc1 = 0, lc2 = 0, lc3 = 0;
  if (upload && P.in_bytes) {
    CK(cudaMemcpyAsync(s.d_buf, s.h_in, P.in_bytes, cudaMemcpyHostToDevice, s.stream));
    stats_.h2d += P.in_bytes;
  }
  if (!upload && !P.force_double)  // resident batch: the queues must start empty on every run
    CK(cudaMemsetAsync(s.d_buf + P.off_rcount, 0, kMaxF64Classes * sizeof(uint32_t), s.stream));
  // (the three timing events are driver calls on the host's critical path -- four packing threads share one driver
  // lock -- so streamed chunks record them only on request; resident runs always do)
  s.timed = timing;
  if (timing) CK(cudaEventRecord(s.ev_k0, s.stream));
  if (lc_probe) lc1 = now_ms();
  // Fork / join.  Launches that carry a sizeable share of the chunk get a stream each (main stream first), so
  // that their tails overlap; the small ones (leftover groups, tail-shaped tasks) share ONE stream.  Only the
  // side streams that are used are forked and joined: every call here is a driver round trip on the host's
  // critical path, and the packing threads of a device contend for one driver lock.
  int n_side_used = 0;
  auto fork = [&](int n_side) -> cudaError_t {
    n_side_used = n_side;
    if (n_side <= 0) return cudaSuccess;
    cudaError_t e = cudaEventRecord(s.ev_fork, s.stream);
    for (int i = 0; i < n_side && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(s.side[i], s.ev_fork, 0);
    return e;
  };
  auto join = [&]() -> cudaError_t {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n_side_used && e == cudaSuccess; ++i) {
      e = cudaEventRecord(s.ev_side[i], s.side[i]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s.stream, s.ev_side[i], 0);
    }
    n_side_used = 0;
    return e;
  };
  // stream of slot index k: 0 = the chunk's own stream, 1.. = side streams
  auto stream_of = [&](int k) { return k <= 0 ? s.stream : s.side[std::min(k, (int)Slot::kSide) - 1]; };
  if (!P.force_double) {
    // launches whose individual tasks run longest go first, so that they overlap the bulk instead of
    // starting in its tail
    std::vector<const F32Range*> ord;
    for (const F32Range& r : P.f32)
      if (r.n_tasks) ord.push_back(&r);
    std::stable_sort(ord.begin(), ord.end(), [](const F32Range* a, const F32Range* b) { return a->max_task_cost > b->max_task_cost; });
    // stream slots: the striped generic launch and every "big" tier launch (>= 1/8 of the chunk's tasks) own one,
    // all small launches share the last one
    size_t total_tasks = 0;
    for (const F32Range* r : ord) total_tasks += r->n_tasks;
    int n_big = P.n_gen ? 1 : 0, n_small = 0;
    for (const F32Range* r : ord) ((size_t)r->n_tasks * 8 >= total_tasks ? n_big : n_small)++;
    const int n_slots = std::min(n_big + (n_small ? 1 : 0), 1 + (int)Slot::kSide);
    CK(fork(std::max(0, n_slots - 1)));
    int next_big = 0;
    if (P.n_gen) {  // striped generic kernel first: its pairs are the longest-running work items
      KParams p;
      fill_kparams(d, s, p, false);
      CK(launch_generic_f32(p, P.gen_ctas, stream_of(next_big++ % n_slots)));
      stats_.launches += 1;
    }
    for (const F32Range* rp : ord) {
      const F32Range& r = *rp;
      const bool big = (size_t)r.n_tasks * 8 >= total_tasks;
      cudaStream_t st = big ? stream_of(next_big++ % n_slots) : stream_of(n_slots - 1);
      KParams p;
      fill_kparams(d, s, p, false);
      p.tasks += r.task0;
      p.n_tasks = r.n_tasks;
      p.hs_cap = r.hs_cap;
      p.hap_stage_bytes = r.hap_stage;
      set_gcp_constants(p, r.gcp);
      if (r.smem > 227 * 1024)
        return set_error(FCS_PHMM_EUNSUPPORTED, "haplotype run does not fit in shared memory (" + std::to_string(r.smem) + " bytes)");
      static const size_t smem_pad = (size_t)env_i64("FCS_PHMM_SMEM_PAD", 0);  // developer knob: lowers residency
      static const bool dbg = env_i64("FCS_PHMM_DEBUG", 0) != 0;
      if (dbg) {
        std::string cl;
        const TaskBucket& bk = s.buckets[r.bucket];
        for (int c = 0; c < r.tk->n_classes; ++c)
          if (bk.cls_mask >> c & 1ull) {
            size_t n = 0;
            for (const Task& t : bk.tasks) n += t.cls == (uint32_t)c;
            cl += " G" + std::to_string(r.tk->classes[c].G) + "R" + std::to_string(r.tk->classes[c].R) + ":" + std::to_string(n);
          }
        // geometric efficiency: useful cells / cells the tiles sweep (32 lanes x R rows x (Lh + G - 1) steps)
        double swept = 0, useful = 0;
        const ReadMeta* rm = reinterpret_cast<const ReadMeta*>(s.h_in + P.off_rmeta);
        const HapMeta* hm = reinterpret_cast<const HapMeta*>(s.h_in + P.off_hmeta);
        for (const Task& t : bk.tasks) {
          const ClassDesc& cd = r.tk->classes[t.cls];
          double cols = 0, hl = 0, rl = 0;
          const bool pr = r.tk->form == 3;  // two haplotype columns per step
          for (uint32_t j = 0; j < t.n_haps; ++j) hl += hm[t.hap0 + j].len;
          for (uint32_t j = 0; j < t.n_haps; j += pr ? 2 : 1)
            cols += std::max(hm[t.hap0 + j].len, (pr && j + 1 < t.n_haps) ? hm[t.hap0 + j + 1].len : 0u) + cd.G - 1;
          for (uint32_t i = 0; i < t.n_reads; ++i) rl += read_len_of(rm[t.read0 + i]);
          swept += (pr ? 64.0 : 32.0) * cd.R * cols;
          useful += rl * hl;
        }
        fprintf(stderr, "[fcs_phmm] f32 launch tier %d form %d key 0x%x tasks %u (%s) smem %zu hs_cap %u stage %u geom_eff %.3f classes%s\n", r.tk->tier,
                r.tk->form, (unsigned)r.gcp, r.n_tasks, big ? "own stream" : "shared stream", r.smem, r.hs_cap, r.hap_stage, useful / swept, cl.c_str());
      }
      CK(r.tk->launch(p, r.n_tasks, r.smem + smem_pad, st));
      stats_.launches += 1;
    }
    CK(join());
  }
  if (timing) CK(cudaEventRecord(s.ev_k1, s.stream));
  if (lc_probe) lc2 = now_ms();
  {
    const int nl = (int)P.f64.size() + (P.gen64_cap ? 1 : 0);
    int li = 0;
    static const bool f64_prio = env_i64("FCS_PHMM_F64_PRIO", 1) != 0;  // developer knob
    // (resident batches run alone: nothing to overtake; one or two launches -- the usual wide + throughput
    // drain of one tier pair, of which only one has work -- go to the chunk's own stream back to back: the
    // fork/join around them costs more host time than their overlap saves)
    static const int f64_serial = (int)env_i64("FCS_PHMM_F64_SERIAL", 2);  // developer knob: 0 off, 1 own stream, 2 one high-priority stream
    const bool serial64 = f64_serial == 1 && upload && nl <= 2;
    const bool one_hp = f64_serial == 2 && upload && nl <= 2;
    const bool hp = f64_prio && upload && nl > 0 && !serial64;
    const int n_hp = one_hp ? std::min(nl, 1) : std::min(nl, (int)Slot::kHp);
    auto pick64 = [&](int i) { return serial64 ? s.stream : (hp ? s.hp[one_hp ? 0 : i % Slot::kHp] : stream_of(i % (1 + (int)Slot::kSide))); };
    if (hp) {
      CK(cudaEventRecord(s.ev_fork, s.stream));
      for (int i = 0; i < n_hp; ++i) CK(cudaStreamWaitEvent(s.hp[i], s.ev_fork, 0));
    } else if (!serial64) {
      CK(fork(std::min(nl - 1, (int)Slot::kSide)));
    }
    if (P.gen64_cap) {
      KParams p;
      fill_kparams(d, s, p, true);
      CK(launch_generic_f64(p, std::min(P.gen_ctas, P.gen64_cap), pick64(li++)));
      stats_.launches += 1;
    }
    for (const F64Range& r : P.f64) {
      KParams p;
      fill_kparams(d, s, p, true);
      p.hs_cap = r.hs_cap;
      p.hap_stage_bytes = r.hap_stage;
      set_gcp_constants(p, r.tk->form ? P.f64_gcp : -1);
      if (r.smem > 227 * 1024)
        return set_error(FCS_PHMM_EUNSUPPORTED, "haplotype too long for the FP64 kernel's shared memory (" + std::to_string(r.smem) + " bytes)");
      const unsigned resident = (unsigned)d.sm_count * (unsigned)std::max(1, r.tk->min_blocks);
      p.n_seg = r.n_seg;
      unsigned grid = 0;
      for (uint32_t k = 0; k < r.n_seg; ++k) {
        p.seg_cls[k] = r.seg_cls[k];
        p.seg_qid[k] = r.seg_qid[k];
        p.seg_min[k] = r.seg_min[k];
        p.seg_max[k] = r.seg_max[k];
        p.seg_cta0[k] = grid;
        const unsigned ng = 32u / r.seg_G[k];
        static const unsigned grid_cap = (unsigned)env_i64("FCS_PHMM_F64_GRID_CAP", 0);  // developer knob
        grid += std::min(std::min((r.seg_cap[k] + ng - 1) / ng, resident), grid_cap ? grid_cap : ~0u);
      }
      p.seg_cta0[r.n_seg] = grid;
      CK(r.tk->launch(p, grid, r.smem, pick64(li++)));
      stats_.launches += 1;
    }
    if (hp) {
      for (int i = 0; i < n_hp; ++i) {
        CK(cudaEventRecord(s.ev_hp[i], s.hp[i]));
        CK(cudaStreamWaitEvent(s.stream, s.ev_hp[i], 0));
      }
    } else if (!serial64) {
      CK(join());
    }
  }
  if (P.finalize) {
    CK(launch_finalize(reinterpret_cast<double*>(s.d_buf + P.off_out), reinterpret_cast<const ReadMeta*>(s.d_buf + P.off_rmeta),
                       reinterpret_cast<const uint32_t*>(s.d_buf + P.off_rnh), s.d_buf + P.off_poor, (uint32_t)P.n_reads, fin_mismap_, fin_err_, s.stream));
    stats_.launches += 1;
  }
  if (timing) CK(cudaEventRecord(s.ev_k2, s.stream));
  if (lc_probe) lc3 = now_ms();
  if (download && P.n_pairs) {
    const size_t nbytes = P.total_bytes - P.off_out;
    CK(cudaMemcpyAsync(s.h_out, s.d_buf + P.off_out, nbytes, cudaMemcpyDeviceToHost, s.stream));
    stats_.d2h += nbytes;
  }
  CK(cudaEventRecord(s.ev_done, s.stream));
  if (lc_probe)
    fprintf(stderr, "[fcs_phmm launch_chunk, us] h2d+ev %.1f  fp32 phase (%zu launches) %.1f  fp64 phase (%zu launches) %.1f  d2h+ev %.1f\n", (lc1 - lc0) * 1e3,
            P.f32.size(), (lc2 - lc1) * 1e3, P.f64.size(), (lc3 - lc2) * 1e3, (now_ms() - lc3) * 1e3);
  return FCS_PHMM_OK;
}

// Wait for the slot's chunk and scatter its results into the caller's arrays.
// developer probe (FCS_PHMM_TIMELINE=1): device-side times of every chunk relative to the start of the call
static cudaEvent_t g_tl_ref = nullptr;
static std::mutex g_tl_mu;
static std::string g_tl_gpu;

int Engine::retire_slot(Device& d, Slot& s, BatchCtx* only_owner) {
  (void)d;
  std::lock_guard<std::mutex> slot_lock(s.mu);
  if (!s.busy || (only_owner && s.owner != only_owner)) return FCS_PHMM_OK;
  const int rc = retire_locked(s);
  // the chunk is gone whatever happened: tell its batch (which may not be the caller's: a worker of the next batch
  // retires what it finds in the slot it is about to reuse)
  BatchCtx* o = s.owner;
  s.busy = false;
  s.owner = nullptr;
  if (o) {
    std::lock_guard<std::mutex> l(o->mu);
    if (rc != FCS_PHMM_OK && o->rc == FCS_PHMM_OK) { o->rc = rc; o->err = last_error(); }
    if (o->pending.fetch_sub(1) == 1) o->cv.notify_all();
  }
  return rc;
}

int Engine::retire_locked(Slot& s) {
  const double tw0 = now_ms();
  CK(cudaEventSynchronize(s.ev_done));
  const double tw1 = now_ms();
  const ChunkPlan& P = s.plan;
  float ms_all = 0.f, ms_main = 0.f;
  if (s.timed) {
    CK(cudaEventElapsedTime(&ms_all, s.ev_k0, s.ev_k2));
    CK(cudaEventElapsedTime(&ms_main, s.ev_k0, s.ev_k1));
  }
  if (g_tl_ref && s.timed) {
    float a = 0, b = 0, c = 0, e = 0;
    cudaEventElapsedTime(&a, g_tl_ref, s.ev_k0);
    cudaEventElapsedTime(&b, g_tl_ref, s.ev_k1);
    cudaEventElapsedTime(&c, g_tl_ref, s.ev_k2);
    cudaEventElapsedTime(&e, g_tl_ref, s.ev_done);
    char buf[160];
    snprintf(buf, sizeof buf, "\n  gpu chunk %5.2f Gcells: h2d done/kernels start %.3f  fp32 done %.3f  fp64 done %.3f  d2h done %.3f", s.plan.cells / 1e9, a, b, c, e);
    std::lock_guard<std::mutex> l(g_tl_mu);
    g_tl_gpu += buf;
  }
  const double* out = reinterpret_cast<const double*>(s.h_out);
  const uint8_t* used = s.h_out + (P.off_used - P.off_out);
  const float* raw = reinterpret_cast<const float*>(s.h_out + (P.off_raw - P.off_out));
  uint64_t n64 = 0;
  const Input& in = *s.input;
  const uint8_t* poor = s.h_out + (P.off_poor - P.off_out);
  size_t rbase = 0;  // packed read index of the region's first read (reads are packed sorted by length: s.order)
  for (size_t k = 0; k < P.regions.size(); ++k) {
    const int64_t g = P.regions[k];
    int32_t nr = 0, nh = 0;
    in.shape(g, nr, nh);
    const size_t n = (size_t)nr * (size_t)nh;
    if (!n) continue;
    if (P.finalize) {
      if (uint8_t* pf = in.poorly(g)) {
        const uint32_t* ord = s.order.data() + rbase;
        for (int32_t i = 0; i < nr; ++i) pf[ord[i]] = poor[rbase + (size_t)i];
      }
    }
    rbase += (size_t)nr;
    const uint64_t o = P.reg_out0[k];
    std::memcpy(in.out(g), out + o, n * sizeof(double));
    if (uint8_t* u = in.used(g)) std::memcpy(u, used + o, n);
    if (keep_raw_)
      if (float* r = in.raw(g)) std::memcpy(r, raw + o, n * sizeof(float));
  }
  for (uint64_t i = 0; i < P.n_pairs; ++i) n64 += used[i];
  stats_.pairs += P.n_pairs;
  stats_.cells += P.cells;
  stats_.fp64_pairs += n64;
  stats_.chunks += 1;
  {
    std::lock_guard<std::mutex> lk(stats_.mu);
    stats_.kernel_ms += ms_all;
    stats_.main_ms += ms_main;
    stats_.wait_ms += tw1 - tw0;
    stats_.scatter_ms += now_ms() - tw1;
  }
  return FCS_PHMM_OK;
}

// ---------------------------------------------------------------------------------------
WorkerPool::WorkerPool(int n_threads) {
  for (int i = 0; i < n_threads; ++i) th_.emplace_back([this] { loop(); });
}
WorkerPool::~WorkerPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto& t : th_) t.join();
}
void WorkerPool::drain(const std::shared_ptr<Run>& r) {
  for (;;) {
    const int j = r->next.fetch_add(1);
    if (j >= r->n_jobs) break;  // r->fn is dereferenced only for a job of this very run, which run() is still waiting for
    (*r->fn)(j);
    std::lock_guard<std::mutex> lk(mu_);
    if (--r->pending == 0) done_cv_.notify_all();
  }
}
void WorkerPool::loop() {
  t_pool_thread = true;
  uint64_t seen = 0;
  for (;;) {
    // Calls arrive back to back (one per active region or per batch): spin briefly for the next one
    // before sleeping, a condition-variable wake-up costs tens of microseconds of a ~2 ms call.
    const double t0 = now_ms();
    while (gen_.load(std::memory_order_acquire) == seen && now_ms() - t0 < 0.2) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::shared_ptr<Run> r;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return stop_ || gen_.load() != seen; });
      if (stop_) return;
      seen = gen_.load();
      r = cur_;
    }
    if (r) drain(r);
  }
}
void WorkerPool::run(int n_jobs, const std::function<void(int)>& fn) {
  if (n_jobs <= 0) return;
  auto r = std::make_shared<Run>();
  r->fn = &fn;
  r->n_jobs = n_jobs;
  r->pending = n_jobs;
  {
    std::lock_guard<std::mutex> lk(mu_);
    cur_ = r;
    gen_.fetch_add(1, std::memory_order_release);
  }
  cv_.notify_all();
  drain(r);  // the caller works too
  std::unique_lock<std::mutex> lk(mu_);
  done_cv_.wait(lk, [&] { return r->pending == 0; });  // every job of THIS run has returned: fn and its captures may go
  if (cur_ == r) cur_.reset();
}

// One call: a single pass sizes every region; regions are split over the devices by DP cells
// (longest-processing-time-first: independent regions, no exchange -- the reference fans out one
// process per genome partition the same way, /root/reference/src/worker-htc.cpp:113-145); each device's
// share is cut into chunks; `pack_threads_` persistent host threads per device (GATK's
// --native-pair-hmm-threads, src/workers/HTCWorker.cpp:85) each plan + pack + launch whole chunks on
// their own pair of slots, so packing chunk k+1 overlaps the device work of chunk k and chunks of
// different threads overlap on the device.
namespace {
// Several callers' inputs seen as one: region g of the batch is region (g - first[c]) of call c.
class CombinedInput : public Input {
 public:
  explicit CombinedInput(const std::vector<const Input*>& parts) : parts_(parts) {
    first_.push_back(0);
    for (const Input* p : parts_) first_.push_back(first_.back() + p->n_regions());
  }
  int64_t n_regions() const override { return first_.back(); }
  void shape(int64_t g, int32_t& nr, int32_t& nh) const override { const auto w = where(g); parts_[w.first]->shape(w.second, nr, nh); }
  InRead read(int64_t g, int32_t i) const override { const auto w = where(g); return parts_[w.first]->read(w.second, i); }
  InHap hap(int64_t g, int32_t j) const override { const auto w = where(g); return parts_[w.first]->hap(w.second, j); }
  double* out(int64_t g) const override { const auto w = where(g); return parts_[w.first]->out(w.second); }
  uint8_t* used(int64_t g) const override { const auto w = where(g); return parts_[w.first]->used(w.second); }
  float* raw(int64_t g) const override { const auto w = where(g); return parts_[w.first]->raw(w.second); }
  uint8_t* poorly(int64_t g) const override { const auto w = where(g); return parts_[w.first]->poorly(w.second); }
  void sum_lens(int64_t g, uint64_t& sr, uint64_t& sh, uint32_t& max_rl) const override { const auto w = where(g); parts_[w.first]->sum_lens(w.second, sr, sh, max_rl); }

 private:
  std::pair<size_t, int64_t> where(int64_t g) const {
    const size_t c = (size_t)(std::upper_bound(first_.begin(), first_.end(), g) - first_.begin()) - 1;
    return {c, g - first_[c]};
  }
  std::vector<const Input*> parts_;
  std::vector<int64_t> first_;
};
}  // namespace

// GATK calls the native PairHMM once per active region from several threads (--native-pair-hmm-threads,
// /root/reference/src/workers/HTCWorker.cpp:85); one region is far too little work for a B200.  Callers that
// arrive while a batch is on the device are queued; whoever finds no leader becomes the leader and runs
// everything queued so far as ONE batch (flat combining).  Results are independent of the batching, so the
// merge is invisible to the callers; if a merged batch fails, its calls are re-run one by one so that each
// caller gets its own error.
// Cheap structural check of a call (pointers and lengths only): what the planner would reject with EINVAL /
// EUNSUPPORTED, found before anything dereferences the arrays (capture) or a merged batch is formed.
static int validate_input(const Input& in) {
  const int64_t n = in.n_regions();
  if (n < 0) return set_error(FCS_PHMM_EINVAL, "negative region count");
  for (int64_t g = 0; g < n; ++g) {
    int32_t nr = 0, nh = 0;
    in.shape(g, nr, nh);
    if (nr < 0 || nh < 0) return set_error(FCS_PHMM_EINVAL, "negative read or haplotype count");
    if (nr == 0 || nh == 0) continue;
    if (!in.out(g)) return set_error(FCS_PHMM_EINVAL, "out_log10 is null");
    for (int32_t i = 0; i < nr; ++i) {
      const InRead r = in.read(g, i);
      if (r.len <= 0 || !r.b || !r.q || !r.i || !r.d || !r.c) return set_error(FCS_PHMM_EINVAL, "read with non-positive length or null array");
      if (r.len > FCS_PHMM_MAX_READ_LEN) return set_error(FCS_PHMM_EUNSUPPORTED, "read longer than FCS_PHMM_MAX_READ_LEN");
    }
    for (int32_t j = 0; j < nh; ++j) {
      const InHap h = in.hap(g, j);
      if (h.len <= 0 || !h.b) return set_error(FCS_PHMM_EINVAL, "haplotype with non-positive length or null array");
      if (h.len > FCS_PHMM_MAX_HAP_LEN) return set_error(FCS_PHMM_EUNSUPPORTED, "haplotype longer than FCS_PHMM_MAX_HAP_LEN");
    }
  }
  return FCS_PHMM_OK;
}

// compute_front() behind a firewall: nothing thrown on the host side of a batch (std::bad_alloc from the planner's
// vectors is the realistic case) may leave the flat-combining leader section.
int Engine::compute_front_noexcept(const Input& in, BatchCtx& ctx) {
  try {
    return compute_front(in, ctx);
  } catch (const std::bad_alloc&) {
    return set_error(FCS_PHMM_ENOMEM, "host allocation failed");
  } catch (const std::exception& ex) {
    return set_error(FCS_PHMM_EINVAL, std::string("internal: ") + ex.what());
  } catch (...) {
    return set_error(FCS_PHMM_EINVAL, "internal: unknown exception");
  }
}

// Drains a batch on scope exit unless finish() already did: a BatchCtx must outlive every chunk whose slot points at it.
struct Engine::BackGuard {
  Engine* e;
  BatchCtx* ctx;
  bool done = false;
  int finish() { done = true; return e->compute_back(*ctx); }
  ~BackGuard() {
    if (done) return;
    try { e->compute_back(*ctx); } catch (...) {}
    // (compute_back throws only while formatting an error; the chunks are retired before that.  Should even the wait be
    // cut short, block until the count is zero: returning earlier would leave dangling owners.)
    while (ctx->pending.load() != 0) std::this_thread::yield();
  }
};

int Engine::compute(const Input& in) {
  // Per call, before it can join a batch: the capture hook -- once per original call (a failing merged batch is
  // re-run call by call below and must not be captured twice), and only for calls that pass the structural checks,
  // so the writer never dereferences a null plane.  (Without capture the planner's own per-read checks reject a
  // malformed call; if it was merged with others, the batch is re-run call by call and only its owner sees the error.)
  if (in.n_regions() < 0) return set_error(FCS_PHMM_EINVAL, "negative region count");
  if (in.n_regions() == 0) return FCS_PHMM_OK;
  if (capture_ && capture_->active()) {
    int rc = validate_input(in);  // an extra pass over the reads: only paid while capturing
    if (rc == FCS_PHMM_OK) rc = capture_->append(in);
    if (rc != FCS_PHMM_OK) return rc;
  }
  PendingCall me;
  me.in = &in;
  std::unique_lock<std::mutex> lk(comb_mu_);
  comb_queue_.push_back(&me);
  comb_waiting_.fetch_add(1);
  while (!me.done) {
    // sleep while another leader is in its front phase -- or while there is nothing to lead: this caller's own call may
    // already be part of a batch whose leader has handed the role on and is waiting for the devices (the queue is then
    // empty until new calls arrive, and their callers take the role themselves)
    if (comb_leader_ || comb_queue_.empty()) {
      comb_cv_.wait(lk);
      continue;
    }
    comb_leader_ = true;
    std::vector<PendingCall*> batch;
    batch.swap(comb_queue_);  // (no allocation: swap)
    comb_waiting_.fetch_sub((int)batch.size());
    lk.unlock();
    // Leader section.  Whatever happens here, every call of the batch is marked done with a result and the
    // leadership is given up: a caller left waiting would hang its JVM thread for good.
    // The leader role covers the FRONT phase only (plan + pack + launch of every chunk).  It is handed on before the
    // batch's last chunks have left the devices, so the next batch's sizing, planning and packing overlap this one's tail
    // (a call on eight devices spends as long in its serial prefix and its first chunk as on the devices).
    bool front_released = false;
    auto release_front = [&] {
      if (front_released) return;
      front_released = true;
      std::lock_guard<std::mutex> l(comb_mu_);
      comb_leader_ = false;
      comb_cv_.notify_all();
    };
    try {
      int rc;
      if (batch.size() == 1) {
        BatchCtx ctx;
        BackGuard guard{this, &ctx};  // whatever is thrown below, ctx is not destroyed while a chunk in flight points at it
        rc = compute_front_noexcept(*batch[0]->in, ctx);
        std::string err = rc != FCS_PHMM_OK ? last_error() : std::string();
        release_front();
        const int rc2 = guard.finish();
        if (rc == FCS_PHMM_OK && rc2 != FCS_PHMM_OK) { rc = rc2; err = last_error(); }
        batch[0]->rc = rc;
        if (rc != FCS_PHMM_OK) batch[0]->err = err;
      } else {
        std::vector<const Input*> parts;
        for (PendingCall* c : batch) parts.push_back(c->in);
        CombinedInput all(parts);
        BatchCtx ctx;
        BackGuard guard{this, &ctx};  // (declared after `all`: the merged input outlives every chunk that points at it)
        rc = compute_front_noexcept(all, ctx);
        release_front();
        const int rc2 = guard.finish();
        if (rc == FCS_PHMM_OK) rc = rc2;
        if (rc != FCS_PHMM_OK) {
          // a call of the merged batch is malformed (or a device failed): re-run call by call, so that only its owner
          // sees the error; serial, each call front + back under the leader role
          {
            std::unique_lock<std::mutex> l(comb_mu_);
            comb_cv_.wait(l, [&] { return !comb_leader_; });
            comb_leader_ = true;
            front_released = false;
          }
          for (PendingCall* c : batch) {
            c->rc = compute_whole(*c->in);
            if (c->rc != FCS_PHMM_OK) c->err = last_error();
          }
          release_front();
        }
      }
    } catch (...) {  // allocation of `parts` / CombinedInput / an error string
      for (PendingCall* c : batch)
        if (c->rc == FCS_PHMM_OK) c->rc = FCS_PHMM_ENOMEM;  // err text stays empty: assigning it could throw again
    }
    release_front();
    lk.lock();
    for (PendingCall* c : batch) c->done = true;
    comb_cv_.notify_all();
  }
  lk.unlock();
  if (me.rc != FCS_PHMM_OK) return set_error(me.rc, me.err.empty() ? std::string("host allocation failed") : me.err);
  return FCS_PHMM_OK;
}

int Engine::compute_whole(const Input& in) {
  BatchCtx ctx;
  BackGuard guard{this, &ctx};
  int rc = compute_front_noexcept(in, ctx);
  const std::string err = rc != FCS_PHMM_OK ? last_error() : std::string();
  const int rc2 = guard.finish();
  if (rc != FCS_PHMM_OK) return set_error(rc, err);
  return rc2;
}

// Back phase of a batch: retire whatever is still in flight for it, wait for the chunks other threads are retiring.
int Engine::compute_back(BatchCtx& ctx) {
  int rc = FCS_PHMM_OK;
  std::string err;
  for (auto& dp : devs_) {
    if (ctx.pending.load() == 0) break;
    bool set = false;
    for (Slot& s : dp->slots) {
      if (ctx.pending.load() == 0) break;
      if (!set) { cudaSetDevice(dp->ordinal); set = true; }
      const int r = retire_slot(*dp, s, &ctx);
      if (r != FCS_PHMM_OK && rc == FCS_PHMM_OK) { rc = r; err = last_error(); }
    }
  }
  std::unique_lock<std::mutex> l(ctx.mu);
  ctx.cv.wait(l, [&] { return ctx.pending.load() == 0; });
  if (rc == FCS_PHMM_OK && ctx.rc != FCS_PHMM_OK) { rc = ctx.rc; err = ctx.err; }
  l.unlock();
  if (rc != FCS_PHMM_OK) return set_error(rc, err);
  return FCS_PHMM_OK;
}

int Engine::compute_front(const Input& in, BatchCtx& ctx) {
  const int64_t n = in.n_regions();
  if (n < 0) return set_error(FCS_PHMM_EINVAL, "negative region count");
  if (n == 0) return FCS_PHMM_OK;
  std::lock_guard<std::mutex> call_lock(front_mu_);
  const size_t D = devs_.size();
  const uint32_t hs_cols = (uint32_t)env_i64("FCS_PHMM_HS_COLS", 640);
  static const bool timeline = env_i64("FCS_PHMM_TIMELINE", 0) != 0;  // developer knob: host timeline of the call on stderr
  // per-chunk device times (fcs_phmm_stats kernel_ms / main_kernel_ms of streamed calls): off unless asked for
  static const bool chunk_timing = timeline || env_i64("FCS_PHMM_CHUNK_TIMING", 0) != 0;
  const double tl0 = now_ms();
  if (timeline && D == 1) {
    if (!g_tl_ref) cudaEventCreate(&g_tl_ref);
    cudaSetDevice(devs_[0]->ordinal);
    cudaEventRecord(g_tl_ref, devs_[0]->slots[0].stream);
    std::lock_guard<std::mutex> l(g_tl_mu);
    g_tl_gpu.clear();
  }
  std::mutex tl_mu;
  std::string tl_text;
  auto tl_mark = [&](const char* what, int w, size_t c) {
    if (!timeline) return;
    char buf[96];
    snprintf(buf, sizeof buf, " %s[w%d c%zu]@%.3f", what, w, c, now_ms() - tl0);
    std::lock_guard<std::mutex> l(tl_mu);
    tl_text += buf;
  };
  // ---- size every region once
  std::vector<uint64_t> rc_cells((size_t)n), rc_pairs((size_t)n), rc_bytes((size_t)n), rc_ub_in((size_t)n);
  std::vector<uint32_t> rc_maxrl((size_t)n);
  auto size_range = [&](int64_t g0, int64_t g1) {
    for (int64_t g = g0; g < g1; ++g) {
      int32_t nr = 0, nh = 0;
      in.shape(g, nr, nh);
      uint64_t sr = 0, sh = 0;
      in.sum_lens(g, sr, sh, rc_maxrl[(size_t)g]);
      rc_cells[(size_t)g] = sr * sh;
      rc_pairs[(size_t)g] = (uint64_t)std::max(0, nr) * (uint64_t)std::max(0, nh);
      rc_bytes[(size_t)g] = 5 * sr + 80ull * (uint64_t)std::max(0, nr) + sh + 20ull * (uint64_t)std::max(0, nh);
      // upper bound of the region's share of a chunk's input section: padded planes + metadata + one task and
      // one striped-path entry per pair at worst
      rc_ub_in[(size_t)g] = 5 * sr + 91ull * (uint64_t)std::max(0, nr) + sh + 27ull * (uint64_t)std::max(0, nh) + 24ull * rc_pairs[(size_t)g];
    }
  };
  // the pass reads one length per read from the caller's (cold) arrays: for a large call it is split over the pool
  // (config 3, 2000 regions / 80k reads: 0.34 ms alone, before any device work can start)
  {
    const int parts = (int)std::min<int64_t>(std::max<int64_t>(1, n / 256), (int64_t)std::max(1, (int)devs_.size() * pack_threads_));
    if (parts <= 1) size_range(0, n);
    else {
      const std::function<void(int)> sizer = [&](int p) { size_range(n * p / parts, n * (p + 1) / parts); };
      pool_->run(parts, sizer);
    }
  }
  // ---- regions -> devices
  std::vector<std::vector<int64_t>> part(D);
  if (D == 1) {
    part[0].resize((size_t)n);
    std::iota(part[0].begin(), part[0].end(), (int64_t)0);
  } else {
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      return rc_cells[(size_t)a] > rc_cells[(size_t)b] || (rc_cells[(size_t)a] == rc_cells[(size_t)b] && a < b); });
    std::vector<uint64_t> load(D, 0);
    for (int64_t g : order) {
      size_t best = 0;
      for (size_t d = 1; d < D; ++d)
        if (load[d] < load[best]) best = d;
      part[best].push_back(g);
      load[best] += rc_cells[(size_t)g];
    }
    for (auto& p : part) std::sort(p.begin(), p.end());
  }
  // ---- per device: chunk boundaries (cells, pairs and byte limits).  Unless the caller fixed a chunk
  // size, aim at about two chunks per packing thread: enough to pipeline host packing against the
  // device, few enough that each launch still has a device-filling number of tasks.  The first chunk
  // of every packing thread is a quarter of the regular size so that the device starts early, the
  // second a half: the host packs ~2.3x faster than the device computes, so sizes may double per round
  // without the device running dry (a 4x jump left it idle: config 2 x8 e2e 3343 -> 3646 GCUPS).
  struct DevWork {
    std::vector<std::vector<int64_t>> chunks;
    std::atomic<size_t> next{0};
    int threads = 0;
  };
  std::vector<DevWork> work(D);
  int n_jobs = 0;
  std::vector<std::pair<int, int>> jobs;  // (device, worker)
  for (size_t d = 0; d < D; ++d) {
    uint64_t total = 0;
    for (int64_t g : part[d]) total += rc_cells[(size_t)g];
    // The chunk schedule (sizes, ramp) is laid out for at least four packing threads even when the handle has fewer:
    // how finely a call is pipelined against the device should not depend on how many cores the process was given
    // (measured on a 4-core mask, config 2: 1 / 2 / 3 / 4 threads with their own schedules 2579 / 2971 / 3262 / 3411 GCUPS
    // end to end, the host work itself being ~0.9 ms of one core per 1.3 ms call).
    static const int min_sched = (int)std::max<int64_t>(1, env_i64("FCS_PHMM_SCHED_THREADS", 4));  // developer knob
    const int sched_threads = std::max(pack_threads_, min_sched);
    int64_t limit = max_chunk_cells_;
    const bool ramp = limit <= 0;
    if (limit <= 0) {
      static const int64_t cpt_x10 = env_i64("FCS_PHMM_CHUNKS_PER_THREAD_X10", 20);  // developer knob
      const int64_t want = (int64_t)(total * 10 / (uint64_t)(cpt_x10 * sched_threads));
      // at most ~3 Gcells (about 1 ms of device work, 8+ waves of CTAs): larger chunks gain nothing on the device, and
      // a batch that merges several callers would otherwise grow every slot's pinned staging to a multiple of what
      // a single call needs (re-allocating pinned memory costs milliseconds per slot)
      static const int64_t cap_cells = env_i64("FCS_PHMM_CHUNK_CAP_CELLS", 3000000000LL);  // developer knob
      limit = std::min<int64_t>(cap_cells, std::max<int64_t>(500000000LL, want));
    }
    // Chunks are cut from the device's regions in order of their longest read (descending): a chunk then holds a
    // narrow band of read lengths, its reads share a few kernel classes, and those classes are popular enough in the
    // chunk to keep their exact rows-per-lane (Planner: fine_len) instead of rounding up to the coarse grid.  Results
    // are scattered by region index, so the order is free.  (Ragged calls only: equal-length regions keep their order.)
    static const bool sort_regions = env_i64("FCS_PHMM_SORT_REGIONS", 1) != 0;  // developer knob
    if (sort_regions && total > 2 * (uint64_t)limit && part[d].size() > 8)
      std::stable_sort(part[d].begin(), part[d].end(), [&](int64_t a, int64_t b) { return rc_maxrl[(size_t)a] > rc_maxrl[(size_t)b]; });
    uint64_t cells = 0, pairs = 0, bytes = 0;
    std::vector<int64_t> cur;
    auto& chunks = work[d].chunks;
    for (int64_t g : part[d]) {
      const size_t k = (size_t)g;
      static const int ramp_style = (int)env_i64("FCS_PHMM_RAMP", 2);
      static const int64_t ramp_floor = env_i64("FCS_PHMM_RAMP_FLOOR_CELLS", 125000000LL);  // developer knob: smallest ramp chunk
      static const int64_t ramp0_div = std::max<int64_t>(1, env_i64("FCS_PHMM_RAMP0_DIV", 4));  // developer knob: first round = limit / this
      int64_t lim_now = limit;
      // ramp: the first round of chunks (one per packing thread) is small so that the device starts after a fraction of
      // a millisecond of planning + packing, later rounds double.  Style 2 adds a round at 1/16 for calls whose regular
      // chunk is large (config 3: 3 Gcells = 1.1 ms of host work before the first launch otherwise).
      const int round = (int)chunks.size() / sched_threads;
      if (ramp && ramp_style == 2 && limit >= 1500000000LL) {
        if (round == 0) lim_now = std::max<int64_t>(limit / 16, ramp_floor);
        else if (round == 1) lim_now = std::max<int64_t>(limit / 4, ramp_floor);
        else if (round == 2) lim_now = std::max<int64_t>(limit / 2, ramp_floor);
      } else if (ramp && round == 0) lim_now = std::max<int64_t>(limit / ramp0_div, ramp_floor);
      else if (ramp && ramp_style >= 1 && round == 1) lim_now = std::max<int64_t>(limit / 2, ramp_floor);
      if (!cur.empty() && cells > 0 &&
          ((int64_t)(cells + rc_cells[k]) > lim_now || pairs + rc_pairs[k] > 0x7fffffffULL || bytes + rc_bytes[k] > (1ull << 31))) {
        chunks.emplace_back(std::move(cur));
        cur.clear();
        cells = pairs = bytes = 0;
      }
      cur.push_back(g);
      cells += rc_cells[k]; pairs += rc_pairs[k]; bytes += rc_bytes[k];
    }
    if (!cur.empty()) chunks.emplace_back(std::move(cur));
    work[d].threads = (int)std::min<size_t>({(size_t)pack_threads_, chunks.size(), devs_[d]->slots.size() / 2});
    // Which thread gets which chunk is decided at run time, so any slot may receive the largest chunk: give
    // every slot of the call room for it now.  Growing a slot later means cudaFreeHost + cudaHostAlloc +
    // cudaFree + cudaMalloc in the middle of the pipeline (measured: calls of 4-700 ms instead of 1.6 ms
    // until all eight slots had met the largest chunk).  A shortfall (striped-path scratch) is still handled
    // by the worker's own ensure_buffers.
    {
      size_t need_in = 0, need_out = 0, need_tot = 0;
      for (const auto& ch : chunks) {
        uint64_t ub = 0, prs = 0;
        for (int64_t g : ch) { ub += rc_ub_in[(size_t)g]; prs += rc_pairs[(size_t)g]; }
        const size_t in_b = (size_t)ub + 8192, out_b = (size_t)prs * 13 + 4096;
        need_in = std::max(need_in, in_b);
        need_out = std::max(need_out, out_b);
        need_tot = std::max(need_tot, in_b + (size_t)prs * 8 + out_b + 8192);
      }
      if (cudaSetDevice(devs_[d]->ordinal) != cudaSuccess) return set_error(FCS_PHMM_ECUDA, "cudaSetDevice failed");
      for (int sidx = 0; sidx < 2 * work[d].threads; ++sidx) {
        Slot& sl = devs_[d]->slots[(size_t)sidx];
        if (need_in > sl.h_in_cap || need_out > sl.h_out_cap || need_tot > sl.d_cap) {
          int rc = retire_slot(*devs_[d], sl);  // the previous batch's chunk may still be using the buffers
          if (rc == FCS_PHMM_OK) rc = ensure_buffers(sl, need_in, need_out, need_tot);
          if (rc != FCS_PHMM_OK) return rc;
        }
      }
    }
    for (int w = 0; w < work[d].threads; ++w) jobs.emplace_back((int)d, w);
    n_jobs += work[d].threads;
  }
  tl_mark("sized+chunked", -1, 0);
  std::atomic<int> first_rc{FCS_PHMM_OK};
  std::mutex err_mu;
  std::string err_text;
  auto fail_with = [&](int rc) {
    int expect = FCS_PHMM_OK;
    if (first_rc.compare_exchange_strong(expect, rc)) {
      std::lock_guard<std::mutex> l2(err_mu);
      err_text = last_error();
    }
  };
  const std::function<void(int)> worker = [&](int job) {
   try {
    Device& d = *devs_[(size_t)jobs[(size_t)job].first];
    DevWork& dw = work[(size_t)jobs[(size_t)job].first];
    const int w = jobs[(size_t)job].second;
    if (cudaSetDevice(d.ordinal) != cudaSuccess) { set_error(FCS_PHMM_ECUDA, "cudaSetDevice failed"); fail_with(FCS_PHMM_ECUDA); return; }
    if (d.has_node_cpus && t_pool_thread) sched_setaffinity(0, sizeof(cpu_set_t), &d.node_cpus);  // never the caller's own thread
    int& use = d.use[(size_t)w];  // (one worker per (device, w) at a time: the front phase is exclusive)
    while (first_rc.load() == FCS_PHMM_OK) {
      const size_t c = dw.next.fetch_add(1);
      if (c >= dw.chunks.size()) break;
      Slot& s = d.slots[(size_t)2 * w + (size_t)(use++ & 1)];
      tl_mark("take", w, c);
      int rc = retire_slot(d, s);
      if (rc != FCS_PHMM_OK) { fail_with(rc); break; }
      tl_mark("retired", w, c);
      const double t0 = now_ms();
      size_t next = 0;
      Planner pl{in, s, use_double_, keep_raw_, INT64_MAX, hs_cols, d.sm_count, c + 1 == dw.chunks.size()};
      pl.finalize = fin_on_.load();
      rc = pl.run(dw.chunks[c], 0, next);
      if (rc == FCS_PHMM_OK && next != dw.chunks[c].size()) rc = set_error(FCS_PHMM_EUNSUPPORTED, "a single region exceeds the chunk limits (2^31 pairs / 2 GiB)");
      if (rc != FCS_PHMM_OK) { fail_with(rc); break; }
      if (s.plan.n_pairs == 0) continue;
      rc = ensure_buffers(s, s.plan.in_bytes, s.plan.total_bytes - s.plan.off_out, s.plan.total_bytes);
      if (rc != FCS_PHMM_OK) { fail_with(rc); break; }
      const double t1 = now_ms();
      tl_mark("planned", w, c);
      rc = pack_chunk(s, in);
      if (rc != FCS_PHMM_OK) { fail_with(rc); break; }
      {
        std::lock_guard<std::mutex> lk2(stats_.mu);
        stats_.plan_ms += t1 - t0;
        stats_.pack_ms += now_ms() - t1;
      }
      s.input = &in;
      tl_mark("packed", w, c);
      rc = launch_chunk(d, s, true, true, chunk_timing);
      if (rc != FCS_PHMM_OK) { cudaStreamSynchronize(s.stream); fail_with(rc); break; }
      {
        std::lock_guard<std::mutex> sl(s.mu);
        ctx.pending.fetch_add(1);
        s.owner = &ctx;
        s.busy = true;
      }
      tl_mark("launched", w, c);
    }
    // Nobody queued behind this batch: drain this worker's own slots here, in parallel with the other workers (the
    // scatter of a large call is ~0.1 ms per chunk).  Otherwise leave the chunks in flight to the next batch's workers,
    // whose front phase starts as soon as this one ends, and to the back phase.
    if (comb_waiting_.load() == 0)
      for (int k = 0; k < 2; ++k) {
        int r2 = retire_slot(d, d.slots[(size_t)2 * w + k], &ctx);
        if (r2 != FCS_PHMM_OK) fail_with(r2);
        tl_mark("drained", w, (size_t)k);
      }
   } catch (...) {  // pool threads have no caller to unwind to (std::bad_alloc from the planner's vectors is the realistic case)
    set_error(FCS_PHMM_ENOMEM, "host allocation failed in a packing thread");
    fail_with(FCS_PHMM_ENOMEM);
    // this worker's chunks in flight still point at the caller's arrays: let them finish, then forget them
    Device& d = *devs_[(size_t)jobs[(size_t)job].first];
    for (int k = 0; k < 2; ++k) {
      Slot& s = d.slots[(size_t)2 * jobs[(size_t)job].second + k];
      std::lock_guard<std::mutex> sl(s.mu);
      if (s.busy && s.owner == &ctx) {
        cudaEventSynchronize(s.ev_done);
        s.busy = false;
        s.owner = nullptr;
        std::lock_guard<std::mutex> l(ctx.mu);
        if (ctx.pending.fetch_sub(1) == 1) ctx.cv.notify_all();
      }
    }
   }
  };
  if (n_jobs == 1) worker(0);
  else pool_->run(n_jobs, worker);
  if (timeline) fprintf(stderr, "[fcs_phmm timeline, ms]%s end@%.3f%s\n", tl_text.c_str(), now_ms() - tl0, g_tl_gpu.c_str());
  if (first_rc.load() != FCS_PHMM_OK) return set_error(first_rc.load(), err_text);
  return FCS_PHMM_OK;
}

// ---------------------------------------------------------------------------------------
int Engine::submit(std::unique_ptr<Input> in, std::shared_ptr<void> keepalive, fcs_phmm_ticket* t) {
  std::unique_ptr<Pending> p(new Pending());
  Pending* raw = p.get();
  std::shared_ptr<Input> sin(in.release());
  raw->th = std::thread([this, raw, sin, keepalive] {
    try {
      raw->rc = compute(*sin);
      if (raw->rc != FCS_PHMM_OK) raw->err = last_error();
    } catch (...) {
      raw->rc = FCS_PHMM_ENOMEM;
    }
  });
  std::lock_guard<std::mutex> lk(tickets_mu_);
  *t = next_ticket_++;
  tickets_[*t] = std::move(p);
  return FCS_PHMM_OK;
}

int Engine::wait(fcs_phmm_ticket t) {
  std::unique_ptr<Pending> p;
  {
    std::lock_guard<std::mutex> lk(tickets_mu_);
    auto it = tickets_.find(t);
    if (it == tickets_.end()) return set_error(FCS_PHMM_ETICKET, "unknown or already-waited ticket");
    p = std::move(it->second);
    tickets_.erase(it);
  }
  if (p->th.joinable()) p->th.join();
  if (p->rc != FCS_PHMM_OK) return set_error(p->rc, p->err);
  return FCS_PHMM_OK;
}

// ---------------------------------------------------------------------------------------
// device-resident batches
namespace {
class FlatInput : public Input {
 public:
  FlatInput(const fcs_phmm_flat_batch& b, double* out, uint8_t* used, float* raw, uint8_t* poorly = nullptr)
      : b_(b), out_(out), used_(used), raw_(raw), poorly_(poorly) {}
  uint8_t* poorly(int64_t g) const override { return poorly_ ? poorly_ + b_.reg_read0[g] : nullptr; }
  int64_t n_regions() const override { return b_.n_regions; }
  void shape(int64_t g, int32_t& nr, int32_t& nh) const override { nr = b_.reg_nreads[g]; nh = b_.reg_nhaps[g]; }
  InRead read(int64_t g, int32_t i) const override {
    const int64_t r = (int64_t)b_.reg_read0[g] + i;
    const int64_t o = b_.rd_off[r];
    return InRead{b_.read_bases + o, b_.read_q + o, b_.read_i + o, b_.read_d + o, b_.read_c + o, b_.rd_len[r]};
  }
  InHap hap(int64_t g, int32_t j) const override {
    const int64_t h = (int64_t)b_.reg_hap0[g] + j;
    return InHap{b_.hap_bases + b_.hp_off[h], b_.hp_len[h]};
  }
  double* out(int64_t g) const override { return out_ ? out_ + b_.reg_out0[g] : nullptr; }
  uint8_t* used(int64_t g) const override { return used_ ? used_ + b_.reg_out0[g] : nullptr; }
  float* raw(int64_t g) const override { return raw_ ? raw_ + b_.reg_out0[g] : nullptr; }
  void sum_lens(int64_t g, uint64_t& sr, uint64_t& sh, uint32_t& max_rl) const override {
    const int32_t* rl = b_.rd_len + b_.reg_read0[g];
    const int32_t* hl = b_.hp_len + b_.reg_hap0[g];
    uint64_t a = 0, b = 0;
    int32_t m = 0;
    for (int32_t i = 0; i < b_.reg_nreads[g]; ++i) { a += (uint64_t)(rl[i] > 0 ? rl[i] : 0); m = rl[i] > m ? rl[i] : m; }
    for (int32_t j = 0; j < b_.reg_nhaps[g]; ++j) b += (uint64_t)(hl[j] > 0 ? hl[j] : 0);
    sr = a;
    sh = b;
    max_rl = (uint32_t)m;
  }

 private:
  fcs_phmm_flat_batch b_;
  double* out_;
  uint8_t* used_;
  float* raw_;
  uint8_t* poorly_;
};
}  // namespace

std::unique_ptr<Input> make_flat_input(const fcs_phmm_flat_batch& b, double* out, uint8_t* used, float* raw, uint8_t* poorly) {
  return std::unique_ptr<Input>(new FlatInput(b, out, used, raw, poorly));
}

int Engine::batch_create(const fcs_phmm_flat_batch* fb, int device_index, Batch** out) {
  *out = nullptr;
  if (!fb) return set_error(FCS_PHMM_EINVAL, "null batch");
  if (device_index < 0 || device_index >= (int)devs_.size()) return set_error(FCS_PHMM_EINVAL, "device index out of range");
  Device& d = *devs_[device_index];
  std::lock_guard<std::mutex> lk(d.mu);
  CK(cudaSetDevice(d.ordinal));
  std::unique_ptr<Batch> b(new Batch());
  b->device_index = device_index;
  // planning needs non-null out pointers only as a validity check; use a dummy base
  static double dummy_out;
  b->input.reset(new FlatInput(*fb, &dummy_out, nullptr, nullptr));
  Slot& s = b->slot;
  CK(init_slot(s));
  std::vector<int64_t> regs((size_t)fb->n_regions);
  std::iota(regs.begin(), regs.end(), (int64_t)0);
  size_t next = 0;
  const uint32_t hs_cols = (uint32_t)env_i64("FCS_PHMM_HS_COLS", 640);
  Planner pl{*b->input, s, use_double_, keep_raw_, INT64_MAX, hs_cols, d.sm_count, true, false};  // resident: one launch set, exact classes
  pl.finalize = fin_on_.load();
  int rc = pl.run(regs, 0, next);
  if (rc == FCS_PHMM_OK && next != regs.size())
    rc = set_error(FCS_PHMM_EUNSUPPORTED, "batch too large for one resident chunk (2^31 pairs / 2 GiB of reads+haplotypes)");
  if (rc == FCS_PHMM_OK) rc = ensure_buffers(s, s.plan.in_bytes, s.plan.total_bytes - s.plan.off_out, s.plan.total_bytes);
  if (rc == FCS_PHMM_OK) rc = pack_chunk(s, *b->input);
  if (rc == FCS_PHMM_OK && s.plan.in_bytes) {
    cudaError_t e = cudaMemcpyAsync(s.d_buf, s.h_in, s.plan.in_bytes, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) rc = set_error(FCS_PHMM_ECUDA, std::string("batch upload: ") + cudaGetErrorString(e));
  }
  if (rc != FCS_PHMM_OK) {
    const std::string saved = last_error();
    free_slot(s);
    return set_error(rc, saved);
  }
  for (size_t k = 0; k < s.plan.regions.size(); ++k) {
    const int64_t g = s.plan.regions[k];
    b->flat_out0.push_back(fb->reg_out0[g]);
    b->reg_pairs.push_back((uint64_t)fb->reg_nreads[g] * (uint64_t)fb->reg_nhaps[g]);
  }
  b->input.reset();  // the caller's arrays are not needed any more
  *out = b.release();
  return FCS_PHMM_OK;
}

int Engine::batch_run(Batch* b, bool timed, float* total_ms, float* main_ms) {
  if (!b) return set_error(FCS_PHMM_EINVAL, "null batch");
  Device& d = *devs_[b->device_index];
  std::lock_guard<std::mutex> lk(d.mu);
  CK(cudaSetDevice(d.ordinal));
  int rc = launch_chunk(d, b->slot, false, false, true);
  if (rc != FCS_PHMM_OK) return rc;
  if (timed) {
    CK(cudaEventSynchronize(b->slot.ev_done));
    float a = 0.f, m = 0.f;
    CK(cudaEventElapsedTime(&a, b->slot.ev_k0, b->slot.ev_k2));
    CK(cudaEventElapsedTime(&m, b->slot.ev_k0, b->slot.ev_k1));
    if (total_ms) *total_ms = a;
    if (main_ms) *main_ms = m;
  }
  return FCS_PHMM_OK;
}

int Engine::batch_sync(Batch* b) {
  if (!b) return set_error(FCS_PHMM_EINVAL, "null batch");
  CK(cudaSetDevice(devs_[b->device_index]->ordinal));
  CK(cudaStreamSynchronize(b->slot.stream));
  return FCS_PHMM_OK;
}

int Engine::batch_download(Batch* b, double* out, uint8_t* used, float* raw) {
  if (!b) return set_error(FCS_PHMM_EINVAL, "null batch");
  Device& d = *devs_[b->device_index];
  std::lock_guard<std::mutex> lk(d.mu);
  CK(cudaSetDevice(d.ordinal));
  Slot& s = b->slot;
  const ChunkPlan& P = s.plan;
  if (!P.n_pairs) return FCS_PHMM_OK;
  CK(cudaMemcpyAsync(s.h_out, s.d_buf + P.off_out, P.total_bytes - P.off_out, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  const double* ho = reinterpret_cast<const double*>(s.h_out);
  const uint8_t* hu = s.h_out + (P.off_used - P.off_out);
  const float* hr = reinterpret_cast<const float*>(s.h_out + (P.off_raw - P.off_out));
  if (raw && !keep_raw_) return set_error(FCS_PHMM_EINVAL, "raw_f32 requested but the handle was created without keep_raw_f32");
  for (size_t k = 0; k < P.regions.size(); ++k) {
    const uint64_t n = b->reg_pairs[k], o = P.reg_out0[k];
    const int64_t f = b->flat_out0[k];
    if (!n) continue;
    if (out) std::memcpy(out + f, ho + o, n * sizeof(double));
    if (used) std::memcpy(used + f, hu + o, n);
    if (raw) std::memcpy(raw + f, hr + o, n * sizeof(float));
  }
  return FCS_PHMM_OK;
}

void Engine::batch_destroy(Batch* b) {
  if (!b) return;
  cudaSetDevice(devs_[b->device_index]->ordinal);
  free_slot(b->slot);
  delete b;
}

// Host-only: plan + pack a batch as one chunk exactly as compute() would (no device needed), check that
// the tasks and the generic pair list cover every (read, hap) pair exactly once, and report what the
// batcher decided.  Used by the CPU tests and to profile the host path.
int plan_check(const fcs_phmm_flat_batch* fb, int sm_count, fcs_phmm_plan_info* out) {
  std::call_once(g_cls_once, build_len_tables);
  std::memset(out, 0, sizeof(*out));
  static double dummy_out;
  std::unique_ptr<Input> in = make_flat_input(*fb, &dummy_out, nullptr, nullptr);
  Slot s;
  std::vector<int64_t> regs((size_t)fb->n_regions);
  std::iota(regs.begin(), regs.end(), (int64_t)0);
  size_t next = 0;
  const uint32_t hs_cols = (uint32_t)env_i64("FCS_PHMM_HS_COLS", 640);
  const double t0 = now_ms();
  Planner pl{*in, s, false, false, INT64_MAX, hs_cols, sm_count > 0 ? sm_count : 148};
  int rc = pl.run(regs, 0, next);
  if (rc != FCS_PHMM_OK) return rc;
  if (next != regs.size()) return set_error(FCS_PHMM_EUNSUPPORTED, "batch too large for one chunk");
  const double t1 = now_ms();
  if (g_plan_prof) {
    fprintf(stderr, "[fcs_phmm plan prof, ms] quality scans %.2f  per-region checks+hap scan %.2f  sort %.2f  tasks %.2f  fp64 caps %.2f  post (task sort, layout) %.2f\n",
            t_prof.scan, t_prof.region, t_prof.sort, t_prof.tasks, t_prof.f64, t_prof.post);
    t_prof = PlanProf();
  }
  ChunkPlan& P = s.plan;
  std::vector<uint8_t> buf(P.in_bytes + 256);  // touched here, so the pack time below excludes page faults
  s.h_in = buf.data();
  const double t1b = now_ms();
  rc = Engine::pack_chunk_static(s, *in);  // touches nothing but the slot
  s.h_in = nullptr;
  if (rc != FCS_PHMM_OK) return rc;
  const double t2 = now_ms();
  out->plan_ms = t1 - t0;
  out->pack_ms = t2 - t1b;
  out->n_pairs = (int64_t)P.n_pairs;
  out->n_tasks = (int64_t)P.n_tasks;
  out->n_generic_pairs = (int64_t)P.n_gen;
  out->n_launches_f32 = 0;
  out->n_tasks_general = out->n_tasks_uniform_gcp = out->n_tasks_all_uniform = out->n_tasks_hap_pairs = 0;
  for (const auto& r : P.f32) {
    out->n_launches_f32 += r.n_tasks ? 1 : 0;
    (r.tk->form == 3 ? out->n_tasks_hap_pairs : (r.tk->form == 2 ? out->n_tasks_all_uniform : (r.tk->form == 1 ? out->n_tasks_uniform_gcp : out->n_tasks_general))) += (int64_t)r.n_tasks;
  }
  out->n_launches_f32 += P.n_gen ? 1 : 0;
  out->n_launches_f64 = (int32_t)P.f64.size() + (P.gen64_cap ? 1 : 0);
  out->n_sym = (int32_t)P.n_sym;
  out->latency_mode = P.latency_mode ? 1 : 0;
  out->in_bytes = (int64_t)P.in_bytes;
  // ---- coverage: every (read, hap) pair exactly once
  const ReadMeta* rm = reinterpret_cast<const ReadMeta*>(buf.data() + P.off_rmeta);
  const HapMeta* hm = reinterpret_cast<const HapMeta*>(buf.data() + P.off_hmeta);
  std::vector<uint8_t> seen((size_t)P.n_pairs, 0);
  double swept = 0, useful = 0;
  size_t max_smem = 0;
  for (const F32Range& r : P.f32) {
    max_smem = std::max(max_smem, r.smem);
    const TaskBucket& bk = s.buckets[r.bucket];
    double r_swept = 0, r_useful = 0;
    for (const Task& t : bk.tasks) {
      const ClassDesc& cd = r.tk->classes[t.cls];
      if ((int)t.n_reads > 32 / cd.G || t.n_reads == 0 || t.n_haps == 0) return set_error(FCS_PHMM_EINVAL, "plan_check: task shape");
      double cols = 0, hl = 0, rl = 0;
      const bool pr = r.tk->form == 3;  // haplotype-pair kernels sweep two columns per step
      for (uint32_t j = 0; j < t.n_haps; ++j) hl += hm[t.hap0 + j].len;
      for (uint32_t j = 0; j < t.n_haps; j += pr ? 2 : 1)
        cols += std::max(hm[t.hap0 + j].len, (pr && j + 1 < t.n_haps) ? hm[t.hap0 + j + 1].len : 0u) + cd.G - 1;
      for (uint32_t i = 0; i < t.n_reads; ++i) {
        const ReadMeta& m = rm[t.read0 + i];
        const uint32_t len = read_len_of(m);
        if ((int)len + 1 > cd.G * cd.R) return set_error(FCS_PHMM_EINVAL, "plan_check: class does not cover the read");
        rl += len;
        for (uint32_t j = 0; j < t.n_haps; ++j) {
          const uint32_t oi = m.out_off + hm[t.hap0 + j].col;
          if (oi >= P.n_pairs || seen[oi]++) return set_error(FCS_PHMM_EINVAL, "plan_check: pair covered twice or out of range");
        }
      }
      swept += (pr ? 64.0 : 32.0) * cd.R * cols;
      useful += rl * hl;
      r_swept += (pr ? 64.0 : 32.0) * cd.R * cols;
      r_useful += rl * hl;
    }
    if (env_i64("FCS_PHMM_DEBUG", 0))  // developer probe: where the cells of a chunk go, launch by launch
      fprintf(stderr, "[fcs_phmm plan_check] launch tier %d form %d: %u tasks, %.3f Gcells useful, geometric efficiency %.3f\n", r.tk->tier, r.tk->form, r.n_tasks,
              r_useful / 1e9, r_swept > 0 ? r_useful / r_swept : 1.0);
  }
  for (const RerunEntry& e2 : s.genlist) {
    const ReadMeta& m = rm[e2.read];
    const uint32_t oi = m.out_off + hm[e2.hap].col;
    if (oi >= P.n_pairs || seen[oi]++) return set_error(FCS_PHMM_EINVAL, "plan_check: generic pair covered twice or out of range");
  }
  for (uint8_t v : seen)
    if (v != 1) return set_error(FCS_PHMM_EINVAL, "plan_check: a pair is not covered");
  out->geometric_efficiency = swept > 0 ? useful / swept : 1.0;
  out->max_smem_bytes = (int64_t)max_smem;
  return FCS_PHMM_OK;
}

int Engine::get_stats(fcs_phmm_stats* s) {
  if (!s) return set_error(FCS_PHMM_EINVAL, "null stats");
  s->pairs = stats_.pairs;
  s->cells = stats_.cells;
  s->fp64_pairs = stats_.fp64_pairs;
  s->kernel_launches = stats_.launches;
  s->h2d_bytes = stats_.h2d;
  s->d2h_bytes = stats_.d2h;
  s->chunks = stats_.chunks;
  std::lock_guard<std::mutex> lk(stats_.mu);
  s->kernel_ms = stats_.kernel_ms;
  s->main_kernel_ms = stats_.main_ms;
  s->host_plan_ms = stats_.plan_ms;
  s->host_pack_ms = stats_.pack_ms;
  s->host_wait_ms = stats_.wait_ms;
  s->host_scatter_ms = stats_.scatter_ms;
  return FCS_PHMM_OK;
}

void Engine::reset_stats() {
  stats_.pairs = 0; stats_.cells = 0; stats_.fp64_pairs = 0; stats_.launches = 0;
  stats_.h2d = 0; stats_.d2h = 0; stats_.chunks = 0;
  std::lock_guard<std::mutex> lk(stats_.mu);
  stats_.kernel_ms = 0;
  stats_.main_ms = 0;
  stats_.plan_ms = stats_.pack_ms = stats_.wait_ms = stats_.scatter_ms = 0;
}

}  // namespace fcsphmm
