/*
 * fcs_pairhmm.h — C ABI of libfcs_pairhmm, the B200-native PairHMM forward-likelihood
 * library for the fcs-genome HaplotypeCaller / Mutect2 path.
 *
 * What it replaces.  falcon-genome itself holds no PairHMM arithmetic; `fcs-genome htc`
 * and `fcs-genome mutect2` launch one GATK JVM per genome partition
 *   /root/reference/src/workers/HTCWorker.cpp:48-113     (command line; PairHMM thread knob at :85 / :105)
 *   /root/reference/src/workers/Mutect2Worker.cpp:109-192
 * next to a Blaze NAM accelerator daemon
 *   /root/reference/src/worker-htc.cpp:99-112, src/worker-mutect2.cpp:152-165,
 *   src/workers/BlazeWorker.cpp:9-26, src/BackgroundExecutor.cpp:13-84
 * and the read-vs-haplotype likelihoods are computed inside the JVM by
 *   VectorLoglessPairHMM.computeLog10Likelihoods -> JNI
 *   Java_com_intel_gkl_pairhmm_IntelPairHmm_{initNative,computeLikelihoodsNative,doneNative}
 * [upstream GATK / Intel GKL, not vendored in the reference].  The entry points below are
 * what a JNI shim with those three symbol names binds (INTEGRATION.md shows the shim):
 *
 *   initNative(readClass, hapClass, use_double, max_threads)  -> fcs_pairhmm_create
 *   computeLikelihoodsNative(reads[], haps[], double out[])   -> fcs_pairhmm_compute  (one region)
 *   doneNative()                                              -> fcs_pairhmm_destroy
 *
 * Semantics: out[r * n_haps + h] = log10 P(read r | hap h) of the
 * logless forward algorithm; quals are masked with & 127; 'N' in the read or the haplotype
 * matches anything; float first (K = 2^120), recomputed in double (K = 2^1020) when the raw
 * float sum is < 1e-28f, decided per pair.  Every (read, hap) result is independent of how
 * the caller batches, orders or splits regions.
 *
 * Rules of the ABI: plain pointers and sizes only; all memory is caller-owned and never
 * retained past the return of a blocking call (or of fcs_pairhmm_wait for a ticket);
 * no exception crosses the boundary and the library never calls exit().  There is NO CPU
 * fallback: without a usable sm_100 device every compute call fails with FCS_PHMM_ENODEV,
 * as a missing NAM binary is fatal in the reference (src/workers/BlazeWorker.cpp:17-20).
 */
#ifndef FCS_PAIRHMM_H
#define FCS_PAIRHMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCS_PHMM_ABI_VERSION 1

#if defined(__GNUC__)
#define FCS_PHMM_API __attribute__((visibility("default")))
#else
#define FCS_PHMM_API
#endif

/* error codes (0 = ok, negative = failure; fcs_pairhmm_last_error() has the text) */
#define FCS_PHMM_OK 0
#define FCS_PHMM_EINVAL (-1)       /* bad argument: null pointer, non-positive length */
#define FCS_PHMM_ENODEV (-2)       /* no usable CUDA device (no CPU fallback exists) */
#define FCS_PHMM_ECUDA (-3)        /* a CUDA runtime call or kernel failed */
#define FCS_PHMM_ENOMEM (-4)       /* host or device allocation failed */
#define FCS_PHMM_EUNSUPPORTED (-5) /* shape outside what the kernels cover (FCS_PHMM_MAX_READ_LEN; > 8 distinct non-ACGTN byte values shared by reads and haplotypes of a chunk) */
#define FCS_PHMM_ETICKET (-6)      /* unknown or already-waited ticket */

#define FCS_PHMM_MAX_READ_LEN 65535
#define FCS_PHMM_MAX_HAP_LEN 65535

typedef struct fcs_phmm_handle fcs_phmm_handle;
typedef struct fcs_phmm_batch fcs_phmm_batch;
typedef int64_t fcs_phmm_ticket;

/* One read: five parallel byte arrays of `len` entries (GKL testcase fields rs, q, i, d, c). */
typedef struct {
  const uint8_t* bases;  /* ASCII; anything other than A C G T N mismatches every haplotype base but N */
  const uint8_t* base_q; /* Phred, NOT +33 */
  const uint8_t* ins_q;
  const uint8_t* del_q;
  const uint8_t* gcp;
  int32_t len;
} fcs_phmm_read;

typedef struct {
  const uint8_t* bases; /* ASCII A C G T N only */
  int32_t len;
} fcs_phmm_hap;

/* One active region = one computeLikelihoodsNative call of the reference path. */
typedef struct {
  const fcs_phmm_read* reads;
  int32_t n_reads;
  const fcs_phmm_hap* haps;
  int32_t n_haps;
  double* out_log10;      /* n_reads * n_haps, index r * n_haps + h (read-major, as GKL) */
  uint8_t* out_used_fp64; /* optional (may be NULL), same shape: 1 if the pair took the double path */
} fcs_phmm_region;

typedef struct {
  uint32_t struct_size;       /* sizeof(fcs_phmm_config), for forward compatibility */
  int32_t n_devices;          /* 0 = every visible device */
  const int32_t* devices;     /* n_devices CUDA ordinals, or NULL for 0..n_devices-1 */
  int32_t use_double;         /* GKL initNative(use_double): force the double path for every pair */
  int32_t max_threads;        /* GKL initNative(max_threads): host packing threads per device (0 = default: 4, or fewer if this process has fewer than 4 cores per device) */
  int32_t slots_per_device;   /* in-flight chunk pipelines (streams) per device; at least 2 per packing thread */
  int64_t max_chunk_cells;    /* split a call into chunks of about this many DP cells, 0 = default */
  int32_t keep_raw_f32;       /* 1 = also return the raw float sums through fcs_pairhmm_compute_flat (tests) */
  int32_t reserved;
} fcs_phmm_config;

/* Counters since creation (or the last reset), summed over devices. */
typedef struct {
  uint64_t pairs;
  uint64_t cells;          /* sum of read_len * hap_len */
  uint64_t fp64_pairs;     /* pairs that took the double path */
  uint64_t kernel_launches;
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t chunks;
  double kernel_ms;        /* CUDA-event time of the kernels (per chunk: first launch to last), summed */
  double main_kernel_ms;   /* of which: the FP32 wavefront kernels */
  double host_plan_ms;     /* host phases of fcs_pairhmm_compute, summed over chunks: task planning, */
  double host_pack_ms;     /*   copying reads/haplotypes into pinned staging,                          */
  double host_wait_ms;     /*   blocked on the device,                                                */
  double host_scatter_ms;  /*   scattering results into the caller's arrays                            */
} fcs_phmm_stats;

/*
 * Flat (structure-of-arrays) batch: the layout a JNI shim builds once per call instead of
 * n_reads small structs, and the layout of the on-disk capture format.
 *   read r of the batch:  rd_len[r] bytes at rd_off[r] in each of the five planes
 *   hap h of the batch:   hp_len[h] bytes at hp_off[h] in hap_bases
 *   region g:             reads reg_read0[g] .. +reg_nreads[g], haps reg_hap0[g] .. +reg_nhaps[g],
 *                         results at out[reg_out0[g] + r * reg_nhaps[g] + h]
 */
typedef struct {
  const uint8_t* read_bases;
  const uint8_t* read_q;
  const uint8_t* read_i;
  const uint8_t* read_d;
  const uint8_t* read_c;
  const int64_t* rd_off;
  const int32_t* rd_len;
  int64_t n_reads;
  const uint8_t* hap_bases;
  const int64_t* hp_off;
  const int32_t* hp_len;
  int64_t n_haps;
  const int32_t* reg_read0;
  const int32_t* reg_nreads;
  const int32_t* reg_hap0;
  const int32_t* reg_nhaps;
  const int64_t* reg_out0;
  int64_t n_regions;
} fcs_phmm_flat_batch;

/* ---- lifecycle -------------------------------------------------------------------- */
FCS_PHMM_API int fcs_pairhmm_abi_version(void);
/* cfg may be NULL (defaults).  On failure *out is NULL and the text is in fcs_pairhmm_last_error(NULL). */
FCS_PHMM_API int fcs_pairhmm_create(const fcs_phmm_config* cfg, fcs_phmm_handle** out);
FCS_PHMM_API void fcs_pairhmm_destroy(fcs_phmm_handle* h);
/* Thread-local message of the last failure on this thread; h may be NULL. */
FCS_PHMM_API const char* fcs_pairhmm_last_error(const fcs_phmm_handle* h);
FCS_PHMM_API int fcs_pairhmm_device_count(const fcs_phmm_handle* h);

/* ---- the reference-facing call -------------------------------------------------------
 * Blocking; many regions per call (one region = one GKL computeLikelihoodsNative).  Packs
 * the caller's host arrays, partitions regions over the handle's devices by cell count,
 * runs H2D -> FP32 wavefront -> FP64 rerun -> D2H per chunk and scatters into out_log10.
 * Thread-safe on one handle (GATK's --native-pair-hmm-threads, /root/reference/src/workers/HTCWorker.cpp:85): calls that arrive
 * while a batch is being planned are merged into the next batch, and a batch is planned and packed while the previous one's last
 * chunks are still on the devices.  The caller's arrays must stay untouched until its own call returns; an input the batcher
 * refuses fails only the call that carried it. */
FCS_PHMM_API int fcs_pairhmm_compute(fcs_phmm_handle* h, const fcs_phmm_region* regions, int32_t n_regions);
/* Same work from the flat layout.  used_fp64 and raw_f32 may be NULL; raw_f32 needs keep_raw_f32. */
FCS_PHMM_API int fcs_pairhmm_compute_flat(fcs_phmm_handle* h, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64,
                             float* raw_f32);
/* Asynchronous variant: the regions and everything they point to must stay valid until wait returns. */
FCS_PHMM_API int fcs_pairhmm_submit(fcs_phmm_handle* h, const fcs_phmm_region* regions, int32_t n_regions,
                       fcs_phmm_ticket* ticket);
FCS_PHMM_API int fcs_pairhmm_wait(fcs_phmm_handle* h, fcs_phmm_ticket ticket);

/* ---- device-resident batches (kernel-only timing, repeated scoring of one batch) -------
 * batch_create packs and uploads once (to the handle's device `device_index`);
 * batch_run enqueues only the kernels; batch_run_timed brackets them with CUDA events on
 * the launching stream and returns the elapsed milliseconds (total and FP32-main share);
 * batch_download copies the results back. */
FCS_PHMM_API int fcs_pairhmm_batch_create(fcs_phmm_handle* h, const fcs_phmm_flat_batch* b, int32_t device_index,
                             fcs_phmm_batch** out);
FCS_PHMM_API int fcs_pairhmm_batch_run(fcs_phmm_handle* h, fcs_phmm_batch* b);
FCS_PHMM_API int fcs_pairhmm_batch_run_timed(fcs_phmm_handle* h, fcs_phmm_batch* b, float* total_ms, float* main_ms);
FCS_PHMM_API int fcs_pairhmm_batch_sync(fcs_phmm_handle* h, fcs_phmm_batch* b);
FCS_PHMM_API int fcs_pairhmm_batch_download(fcs_phmm_handle* h, fcs_phmm_batch* b, double* out, uint8_t* used_fp64,
                               float* raw_f32);
FCS_PHMM_API int64_t fcs_pairhmm_batch_pairs(const fcs_phmm_batch* b);
FCS_PHMM_API int64_t fcs_pairhmm_batch_cells(const fcs_phmm_batch* b);
FCS_PHMM_API int32_t fcs_pairhmm_batch_launches(const fcs_phmm_batch* b); /* kernels one batch_run enqueues */
FCS_PHMM_API void fcs_pairhmm_batch_destroy(fcs_phmm_handle* h, fcs_phmm_batch* b);

/* ---- capture / replay of testcases at this boundary -----------------
 * set_capture: append every region passed to compute / compute_flat / submit to `path` (format in
 * falcon-genome_b200/csrc/phmm_capture.h); NULL or "" stops.  The environment variable
 * FCS_PHMM_CAPTURE=<path> enables it at create time (e.g. under a JVM through the JNI shim).
 * capture_load: read such a file into a flat batch that points into `*owner`; free with capture_free.
 * Both are host-only (no device needed). */
FCS_PHMM_API int fcs_pairhmm_set_capture(fcs_phmm_handle* h, const char* path);
FCS_PHMM_API int fcs_pairhmm_capture_load(const char* path, fcs_phmm_flat_batch* out, void** owner);
/* Same, from capture blocks in memory (no file header): what the fcs-pairhmm-nam daemon receives. */
FCS_PHMM_API int fcs_pairhmm_capture_parse(const void* blocks, uint64_t n_bytes, fcs_phmm_flat_batch* out, void** owner);
FCS_PHMM_API void fcs_pairhmm_capture_free(void* owner);

/* ---- GATK-side steps either side of the kernel (host-only) -------------
 * [upstream GATK4 PairHMMLikelihoodCalculationEngine, restated]  prepare_read turns a raw read into the
 * four qual arrays the kernel takes: base quals capped by mapq (mapq < 0 = no cap), then
 * q < base_q_threshold -> min_usable_q; ins/del quals from the BAM tags or default_indel_q, lowered by
 * the PCR indel model (0 none, 1 hostile, 2 aggressive, 3 conservative) inside tandem repeats;
 * gcp constant.  params == NULL: {18, 6, 45, 10, 3}.  finalize_region caps every read's row at
 * best + log10_global_mismapping_rate (GATK: -4.5) and flags reads whose best likelihood is below
 * min(2, ceil(len * expected_error_rate_per_base (0.02))) * -4.0. */
typedef struct {
  int32_t base_q_threshold, min_usable_q, default_indel_q, gcp, pcr_model;
} fcs_phmm_prep_params;
FCS_PHMM_API int fcs_pairhmm_prepare_read(const uint8_t* bases, const uint8_t* raw_base_q, int32_t len, int32_t mapq,
                                          const uint8_t* bam_ins_q, const uint8_t* bam_del_q, const fcs_phmm_prep_params* params,
                                          uint8_t* out_base_q, uint8_t* out_ins_q, uint8_t* out_del_q, uint8_t* out_gcp);
FCS_PHMM_API int fcs_pairhmm_finalize_region(double* log10_likelihoods, int32_t n_reads, int32_t n_haps, const int32_t* read_len,
                                             double log10_global_mismapping_rate, double expected_error_rate_per_base,
                                             uint8_t* out_poorly_modeled);

/* The post-processing half fused into the device pipeline ("cheap to fuse on GPU").  While
 * enabled on a handle, every compute call caps each read's row at  best + log10_global_mismapping_rate  on the
 * device, after the FP64 reruns and before the download -- out_log10 then holds what GATK's
 * normalizeLikelihoods would leave -- and fcs_pairhmm_compute_flat_finalized also returns the poorly-modelled flag
 * per read (best < min(2, ceil(len * expected_error_rate_per_base)) * -4).  Same arithmetic as
 * fcs_pairhmm_finalize_region, so both give the same bits.  p == NULL or p->enabled == 0 switches it off. */
typedef struct {
  int32_t enabled;
  int32_t reserved;
  double log10_global_mismapping_rate;  /* GATK default -4.5 */
  double expected_error_rate_per_base;  /* GATK default 0.02 */
} fcs_phmm_finalize_params;
FCS_PHMM_API int fcs_pairhmm_set_finalize(fcs_phmm_handle* h, const fcs_phmm_finalize_params* p);
/* out_poorly_modeled: one byte per read of the batch (index = read index in the flat batch), may be NULL. */
FCS_PHMM_API int fcs_pairhmm_compute_flat_finalized(fcs_phmm_handle* h, const fcs_phmm_flat_batch* b, double* out_log10, uint8_t* out_used_fp64,
                                                    uint8_t* out_poorly_modeled);

/* ---- service seam: client of the fcs-pairhmm-nam daemon -----------------
 * The daemon (falcon-genome_b200/csrc/fcs_pairhmm_nam.cpp) owns the GPUs for the lifetime of a stage, as the
 * Blaze NAM does in the reference (src/worker-htc.cpp:99-112, src/BackgroundExecutor.cpp:13-84).  These
 * symbols live in libfcs_pairhmm_client.so, which has no CUDA dependency. */
typedef struct fcs_phmm_remote fcs_phmm_remote;
FCS_PHMM_API int fcs_pairhmm_remote_open(const char* socket_path, fcs_phmm_remote** out);
FCS_PHMM_API int fcs_pairhmm_remote_compute_flat(fcs_phmm_remote* r, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64);
FCS_PHMM_API const char* fcs_pairhmm_remote_last_error(const fcs_phmm_remote* r);
FCS_PHMM_API void fcs_pairhmm_remote_close(fcs_phmm_remote* r);
/* Zero-copy variant: build the batch IN the connection's segment.  reserve() lays out a batch of the announced shape
 * (read_bytes = bytes of each of the five read planes, hap_bytes = haplotype bytes, n_pairs = sum over regions of
 * reads x haplotypes; reads and haplotypes indexed densely in region order) and hands out writable views; the caller
 * fills every array, then compute_reserved() rings the daemon.  Results are read in place (out_log10[reg_out0 + r*nh + h]
 * with reg_out0 the running sum of pairs) and stay valid until the next reserve / compute call on the connection.
 * This is what the JNI shim of a JVM behind the daemon does: GetByteArrayRegion straight into the planes. */
typedef struct {
  uint8_t *read_bases, *read_q, *read_i, *read_d, *read_c;
  int64_t* rd_off;
  int32_t* rd_len;
  uint8_t* hap_bases;
  int64_t* hp_off;
  int32_t* hp_len;
  int32_t *reg_read0, *reg_nreads, *reg_hap0, *reg_nhaps;
  const double* out_log10;
  const uint8_t* out_used_fp64;
} fcs_phmm_remote_views;
FCS_PHMM_API int fcs_pairhmm_remote_reserve(fcs_phmm_remote* r, int64_t n_regions, int64_t n_reads, int64_t n_haps, uint64_t read_bytes,
                                            uint64_t hap_bytes, uint64_t n_pairs, fcs_phmm_remote_views* views);
FCS_PHMM_API int fcs_pairhmm_remote_compute_reserved(fcs_phmm_remote* r);
/* 1 while requests travel through the connection's shared-memory segment (the default; the batch is written
 * once into a sealed memfd that the daemon maps, csrc/phmm_shm.h), 0 on the byte-stream protocol
 * (FCS_PHMM_REMOTE_SHM=0, or a daemon that declined the segment). */
FCS_PHMM_API int fcs_pairhmm_remote_uses_shm(const fcs_phmm_remote* r);

/* ---- introspection ------------------------------------------------------------------ */
FCS_PHMM_API int fcs_pairhmm_get_stats(fcs_phmm_handle* h, fcs_phmm_stats* out);
FCS_PHMM_API int fcs_pairhmm_reset_stats(fcs_phmm_handle* h);
/* Host-only: run the batcher on a batch (one chunk, no device needed), verify that its tasks cover every
 * (read, hap) pair exactly once with classes that fit the reads, and report its decisions. */
typedef struct {
  int64_t n_pairs, n_tasks, n_generic_pairs, in_bytes, max_smem_bytes;
  int32_t n_launches_f32, n_launches_f64, n_sym, latency_mode;
  double geometric_efficiency; /* useful cells / cells swept by the tiles of the FP32 tasks */
  double plan_ms, pack_ms;
  /* FP32 tasks per kernel form: general, uniform gap-continuation quality, all transition qualities uniform */
  int64_t n_tasks_general, n_tasks_uniform_gcp, n_tasks_all_uniform;
  /* FP32 tasks of the haplotype-pair kernels (uniform gap-continuation quality, two haplotypes per lane in packed f32x2 arithmetic) */
  int64_t n_tasks_hap_pairs;
} fcs_phmm_plan_info;
FCS_PHMM_API int fcs_pairhmm_plan_check(const fcs_phmm_flat_batch* b, int32_t sm_count, fcs_phmm_plan_info* out);
/* The transition / prior lookup tables the kernels use, for bit-level checks against the oracle
 * (host-side, needs no device): ph2pr[q], matchToMatch(i,d) in float and double. */
FCS_PHMM_API float fcs_pairhmm_lut_ph2pr_f32(int q);
FCS_PHMM_API double fcs_pairhmm_lut_ph2pr_f64(int q);
FCS_PHMM_API float fcs_pairhmm_lut_mm_f32(int ins_q, int del_q);
FCS_PHMM_API double fcs_pairhmm_lut_mm_f64(int ins_q, int del_q);
/* Kernel class (lanes per read G, rows per lane R) the batcher picks for a read length;
 * returns 0 and fills G/R, or FCS_PHMM_EUNSUPPORTED. */
FCS_PHMM_API int fcs_pairhmm_kernel_class(int32_t read_len, int32_t fp64, int32_t* lanes_per_read, int32_t* rows_per_lane);

#ifdef __cplusplus
}
#endif
#endif /* FCS_PAIRHMM_H */
