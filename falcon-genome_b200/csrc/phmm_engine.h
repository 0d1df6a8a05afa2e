// phmm_engine.h — host library behind the C ABI: packer (batcher), per-device chunk
// pipeline and multi-GPU dispatcher.  See include/fcs_pairhmm.h for the contract.
#pragma once
#include <cuda_runtime.h>

#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fcs_pairhmm.h"
#include "phmm_registry.h"
#include "phmm_types.h"

namespace fcsphmm {

int set_error(int code, const std::string& msg);
const char* last_error();

struct InRead {
  const uint8_t *b, *q, *i, *d, *c;
  int32_t len;
};
struct InHap {
  const uint8_t* b;
  int32_t len;
};

// Uniform view over the two input forms (array of region structs / flat batch).
class Input {
 public:
  virtual ~Input() {}
  virtual int64_t n_regions() const = 0;
  virtual void shape(int64_t g, int32_t& nr, int32_t& nh) const = 0;
  virtual InRead read(int64_t g, int32_t i) const = 0;
  virtual InHap hap(int64_t g, int32_t j) const = 0;
  virtual double* out(int64_t g) const = 0;
  virtual uint8_t* used(int64_t g) const = 0;
  virtual float* raw(int64_t g) const = 0;
  // poorly-modelled flags of region g's reads (caller's read order), or null (finalize epilogue)
  virtual uint8_t* poorly(int64_t g) const { (void)g; return nullptr; }
  // Sum of the (non-negative) read and haplotype lengths of region g and its longest read: the sizing pass of a call
  // walks every read once, so inputs override this with a loop over their own arrays (no virtual call per read).
  virtual void sum_lens(int64_t g, uint64_t& sr, uint64_t& sh, uint32_t& max_rl) const {
    int32_t nr = 0, nh = 0;
    shape(g, nr, nh);
    sr = sh = 0;
    max_rl = 0;
    for (int32_t i = 0; i < nr; ++i) { const int32_t l = read(g, i).len; sr += (uint64_t)(l > 0 ? l : 0); if (l > 0 && (uint32_t)l > max_rl) max_rl = (uint32_t)l; }
    for (int32_t j = 0; j < nh; ++j) { const int32_t l = hap(g, j).len; sh += (uint64_t)(l > 0 ? l : 0); }
  }
};

struct F32Range {
  const TierKernel* tk;
  uint32_t task0, n_tasks, hs_cap, hap_stage;
  uint32_t bucket;  // index of the TaskBucket the tasks come from
  int gcp;          // >= 0: uniform launch: gcp | ins << 8 | del << 16 (ins / del for the all-uniform form), -1: general form
  size_t smem;      // dynamic shared memory: max over the classes present
  uint32_t max_task_cost;
};
// Tasks of one launch: one tier kernel x one uniform-quality key (or -1 = general form).
struct TaskBucket {
  const TierKernel* tk = nullptr;
  int gcp = -1;
  uint32_t hs = 0, stage = 0;
  uint64_t cls_mask = 0;  // classes of the tier that occur
  uint32_t max_task_cost = 0;
  std::vector<Task> tasks;
};
struct F64Queue {
  uint32_t qid, cap, maxlh;
};
// One FP64 launch: the queues whose class lives in this tier kernel, one CTA segment per queue.
struct F64Range {
  const TierKernel* tk;
  uint32_t n_seg, hs_cap, hap_stage;
  uint16_t seg_cls[32], seg_qid[32], seg_G[32];
  uint32_t seg_cap[32], seg_min[32], seg_max[32];
  size_t smem;
};

// Host-side description of one packed chunk (what a slot currently holds).
struct ChunkPlan {
  std::vector<int64_t> regions;      // indices into the Input
  std::vector<uint64_t> reg_out0;    // pair offset of each region inside the chunk
  uint64_t n_reads = 0, n_haps = 0, n_pairs = 0, cells = 0;
  size_t off_reads = 0, off_haps = 0, off_rmeta = 0, off_hmeta = 0, off_tasks = 0, off_rbase = 0, off_rcount = 0;
  size_t off_rerun = 0, in_bytes = 0;  // input part = [0, in_bytes)
  size_t off_out = 0, off_used = 0, off_raw = 0, total_bytes = 0;
  bool finalize = false;             // per-read cap + poorly-modelled flag epilogue
  size_t off_rnh = 0, off_poor = 0;  // finalize: haplotypes per read (input, u32), flags per read (output, u8)
  size_t n_tasks = 0;
  size_t off_genlist = 0, off_scratch = 0;
  uint32_t n_gen = 0, gen64_cap = 0, gen_ctas = 0, scratch_cols = 0;  // striped generic path
  std::vector<F32Range> f32;
  std::vector<F64Range> f64;
  std::vector<F64Queue> queues;
  bool force_double = false;
  uint32_t n_sym = 6;         // prior-table symbol rows: 5 (no N in any haplotype), 6 (N), 7 + e (haplotype bytes outside ACGTN)
  uint64_t extra_bytes = 0;   // byte e = haplotype byte value that owns symbol row kCodeExtra0 + e (it also occurs in some read)
  bool latency_mode = false;  // under-filled chunk: widest lane groups, one haplotype per task
  int f64_gcp = -1;  // >= 0: every read of the chunk shares this gap-continuation quality
  int launches() const;
};

// One batch of compute() between its front phase (plan + pack + launch, exclusive) and the moment its last chunk has been
// retired (results scattered).  Chunks are retired by whoever needs their slot next -- a worker of the same batch, a worker
// of the NEXT batch (whose front phase overlaps this batch's tail on the devices) or the batch's own leader in its back phase.
struct BatchCtx {
  std::atomic<int> pending{0};  // chunks launched and not yet retired
  std::mutex mu;
  std::condition_variable cv;
  int rc = 0;                   // first error met while retiring a chunk of this batch (guarded by mu)
  std::string err;
};

struct Slot {
  std::mutex mu;                // guards busy / owner / the retirement of the chunk in flight
  BatchCtx* owner = nullptr;    // batch of the chunk in flight (busy)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_k2 = nullptr, ev_done = nullptr;
  // kernels of different classes are forked onto side streams so that their tails overlap
  static constexpr int kSide = 3;
  cudaStream_t side[kSide] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_side[kSide] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr;
  // the FP64 rerun launches go to high-priority streams: their CTAs (mostly empty: a handful of pairs per
  // chunk) are then scheduled as soon as a slot frees up instead of queueing behind the FP32 CTAs of the
  // chunks that were launched in the meantime, which would hold back this chunk's download
  static constexpr int kHp = 2;
  cudaStream_t hp[kHp] = {nullptr, nullptr};
  cudaEvent_t ev_hp[kHp] = {nullptr, nullptr};
  uint8_t* h_in = nullptr;
  size_t h_in_cap = 0;
  uint8_t* h_out = nullptr;
  size_t h_out_cap = 0;
  uint8_t* d_buf = nullptr;
  size_t d_cap = 0;
  bool busy = false;
  bool timed = false;  // the chunk in flight recorded its timing events (ev_k0 / ev_k1 / ev_k2)
  ChunkPlan plan;
  const Input* input = nullptr;
  // packer scratch (reused)
  std::vector<TaskBucket> buckets;
  std::vector<uint32_t> order;
  std::vector<uint32_t> rlayout;       // per read (same order as ukeys): blob layout flags (phmm_types.h read_layout)
  std::vector<uint32_t> hap_order;     // per haplotype of the chunk (region by region): the caller's index of the haplotype packed at this position
  std::vector<int> ukeys;              // per read (region by region, caller's order): gcp | ins << 8 | del << 16 if all three are constant, else -1
  std::vector<RerunEntry> genlist;     // (read, hap) pairs of the striped generic path
  std::vector<uint8_t> gen_flags;      // per chunk-wide read: takes the generic path
};

struct Device {
  int ordinal = 0;
  int sm_count = 0;
  void* d_ph2pr_f = nullptr;
  void* d_mm_f = nullptr;
  void* d_ph2pr_d = nullptr;
  void* d_mm_d = nullptr;
  std::vector<Slot> slots;
  std::vector<int> use;  // per packing thread: which of its two slots comes next (kept across batches: the older chunk retires first)
  std::mutex mu;  // one call at a time drives a device's slots
  // FCS_PHMM_NUMA_BIND=1 (opt-in): the cores of the NUMA node the GPU hangs off, intersected with the process's
  // affinity mask.  Pool threads that work for this device run there and its pinned staging is allocated from there.
  cpu_set_t node_cpus;
  bool has_node_cpus = false;
};

struct Stats {
  std::atomic<uint64_t> pairs{0}, cells{0}, fp64_pairs{0}, launches{0}, h2d{0}, d2h{0}, chunks{0};
  std::mutex mu;
  double kernel_ms = 0, main_ms = 0;
  double plan_ms = 0, pack_ms = 0, wait_ms = 0, scatter_ms = 0;  // host-side phases of compute()
};

// Persistent host threads: run(n, fn) executes fn(0..n-1) on the pool plus the calling thread.
class WorkerPool {
 public:
  explicit WorkerPool(int n_threads);
  ~WorkerPool();
  void run(int n_jobs, const std::function<void(int)>& fn);

 private:
  // One run() call.  Workers snapshot the descriptor under mu_ and keep it alive through the shared_ptr, so a
  // thread that is late (preempted between taking an index and testing it) only ever compares that index with
  // the job count of the run it belongs to and can never execute a job of a later run.
  struct Run {
    const std::function<void(int)>* fn = nullptr;
    int n_jobs = 0;
    int pending = 0;  // guarded by mu_
    std::atomic<int> next{0};
  };
  void loop();
  void drain(const std::shared_ptr<Run>& r);
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::shared_ptr<Run> cur_;  // guarded by mu_
  std::atomic<uint64_t> gen_{0};
  bool stop_ = false;
};

struct Batch;  // device-resident batch
class CaptureWriter;

class Engine {
 public:
  static int create(const fcs_phmm_config* cfg, Engine** out);
  ~Engine();
  int compute(const Input& in);       // coalesces concurrent callers into one device batch
  int compute_front(const Input& in, BatchCtx& ctx);  // one batch: plan, pack and launch every chunk (exclusive: front_mu_)
  int compute_back(BatchCtx& ctx);                    // wait until every chunk of the batch has been retired
  int compute_front_noexcept(const Input& in, BatchCtx& ctx);
  int compute_whole(const Input& in);                 // front + back
  struct BackGuard;
  int submit(std::unique_ptr<Input> in, std::shared_ptr<void> keepalive, fcs_phmm_ticket* t);
  int wait(fcs_phmm_ticket t);
  int batch_create(const fcs_phmm_flat_batch* b, int device_index, Batch** out);
  int batch_run(Batch* b, bool timed, float* total_ms, float* main_ms);
  int batch_sync(Batch* b);
  int batch_download(Batch* b, double* out, uint8_t* used, float* raw);
  void batch_destroy(Batch* b);
  int get_stats(fcs_phmm_stats* s);
  void reset_stats();
  int device_count() const { return (int)devs_.size(); }
  static int pack_chunk_static(Slot& s, const Input& in);
  int set_capture(const char* path);  // nullptr / empty = stop capturing
  void set_finalize(bool on, double log10_mismap, double err_rate) { fin_on_ = on; fin_mismap_ = log10_mismap; fin_err_ = err_rate; }

 private:
  Engine();
  int init(const fcs_phmm_config* cfg);
  int pack_chunk(Slot& s, const Input& in);
  int launch_chunk(Device& d, Slot& s, bool upload, bool download, bool timing);
  int retire_slot(Device& d, Slot& s, BatchCtx* only_owner = nullptr);
  int retire_locked(Slot& s);  // wait for the chunk, scatter its results (caller holds s.mu)
  int ensure_buffers(Slot& s, size_t in_bytes, size_t out_bytes, size_t dev_bytes);
  void fill_kparams(const Device& d, const Slot& s, KParams& p, bool f64) const;

  std::vector<std::unique_ptr<Device>> devs_;
  bool use_double_ = false;
  bool keep_raw_ = false;
  std::atomic<bool> fin_on_{false};
  double fin_mismap_ = -4.5, fin_err_ = 0.02;
  int pack_threads_ = 0;
  int64_t max_chunk_cells_ = 0;
  Stats stats_;
  std::mutex tickets_mu_;
  struct Pending {
    std::thread th;
    int rc = 0;
    std::string err;
  };
  std::map<fcs_phmm_ticket, std::unique_ptr<Pending>> tickets_;
  fcs_phmm_ticket next_ticket_ = 1;
  std::unique_ptr<CaptureWriter> capture_;
  std::unique_ptr<WorkerPool> pool_;
  std::mutex front_mu_;  // one batch at a time plans, packs and launches (its chunks may still be on the devices when the next one starts)
  // flat combining of concurrent compute() calls: the caller that finds no leader becomes one and runs
  // everything queued so far as ONE batch; the others sleep until their call is marked done
  struct PendingCall {
    const Input* in = nullptr;
    int rc = 0;
    std::string err;
    bool done = false;
  };
  std::mutex comb_mu_;
  std::condition_variable comb_cv_;
  std::vector<PendingCall*> comb_queue_;
  bool comb_leader_ = false;
  std::atomic<int> comb_waiting_{0};  // calls queued and not yet taken by a leader
  friend struct Batch;
};

struct Batch {
  int device_index = 0;
  Slot slot;  // owns the buffers; plan describes the single chunk
  std::unique_ptr<Input> input;
  std::vector<int64_t> flat_out0;   // caller's reg_out0 per planned region
  std::vector<uint64_t> reg_pairs;  // pairs per planned region
};

int prepare_read(const uint8_t* bases, const uint8_t* raw_q, int32_t len, int32_t mapq, const uint8_t* bam_ins, const uint8_t* bam_del,
                 const fcs_phmm_prep_params* pp, uint8_t* out_q, uint8_t* out_i, uint8_t* out_d, uint8_t* out_c);
int finalize_region(double* l, int32_t n_reads, int32_t n_haps, const int32_t* read_len, double log10_mismap, double err_rate, uint8_t* poorly);

int plan_check(const fcs_phmm_flat_batch* fb, int sm_count, fcs_phmm_plan_info* out);

std::unique_ptr<Input> make_flat_input(const fcs_phmm_flat_batch& b, double* out, uint8_t* used, float* raw, uint8_t* poorly = nullptr);

}  // namespace fcsphmm
