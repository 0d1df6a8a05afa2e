// phmm_api.cu — the extern "C" surface declared in include/fcs_pairhmm.h.
// No exception leaves this file; every failure is a negative code + thread-local text.
#include <cstring>
#include <new>

#include "phmm_capture.h"
#include "phmm_engine.h"
#include "phmm_luts.h"

using namespace fcsphmm;

struct fcs_phmm_handle {
  Engine* e;
};
struct fcs_phmm_batch {
  Batch* b;
};

namespace {

class RegionInput : public Input {
 public:
  RegionInput(const fcs_phmm_region* r, int64_t n) : r_(r), n_(n) {}
  int64_t n_regions() const override { return n_; }
  void shape(int64_t g, int32_t& nr, int32_t& nh) const override { nr = r_[g].n_reads; nh = r_[g].n_haps; }
  InRead read(int64_t g, int32_t i) const override {
    const fcs_phmm_read& x = r_[g].reads[i];
    return InRead{x.bases, x.base_q, x.ins_q, x.del_q, x.gcp, x.len};
  }
  InHap hap(int64_t g, int32_t j) const override {
    const fcs_phmm_hap& x = r_[g].haps[j];
    return InHap{x.bases, x.len};
  }
  double* out(int64_t g) const override { return r_[g].out_log10; }
  uint8_t* used(int64_t g) const override { return r_[g].out_used_fp64; }
  float* raw(int64_t) const override { return nullptr; }
  void sum_lens(int64_t g, uint64_t& sr, uint64_t& sh, uint32_t& max_rl) const override {
    const fcs_phmm_region& x = r_[g];
    uint64_t a = 0, b = 0;
    int32_t m = 0;
    for (int32_t i = 0; i < x.n_reads; ++i) { const int32_t l = x.reads[i].len; a += (uint64_t)(l > 0 ? l : 0); m = l > m ? l : m; }
    for (int32_t j = 0; j < x.n_haps; ++j) b += (uint64_t)(x.haps[j].len > 0 ? x.haps[j].len : 0);
    sr = a;
    sh = b;
    max_rl = (uint32_t)m;
  }

 private:
  const fcs_phmm_region* r_;
  int64_t n_;
};

int check_regions(const fcs_phmm_region* regions, int32_t n) {
  if (n < 0) return set_error(FCS_PHMM_EINVAL, "negative region count");
  if (n > 0 && !regions) return set_error(FCS_PHMM_EINVAL, "null regions");
  for (int32_t g = 0; g < n; ++g) {
    if (regions[g].n_reads > 0 && !regions[g].reads) return set_error(FCS_PHMM_EINVAL, "null reads array");
    if (regions[g].n_haps > 0 && !regions[g].haps) return set_error(FCS_PHMM_EINVAL, "null haps array");
  }
  return FCS_PHMM_OK;
}

int check_flat(const fcs_phmm_flat_batch* b) {
  if (!b) return set_error(FCS_PHMM_EINVAL, "null batch");
  if (b->n_regions < 0 || b->n_reads < 0 || b->n_haps < 0) return set_error(FCS_PHMM_EINVAL, "negative count");
  if (b->n_regions > 0 && (!b->reg_read0 || !b->reg_nreads || !b->reg_hap0 || !b->reg_nhaps || !b->reg_out0))
    return set_error(FCS_PHMM_EINVAL, "null region table");
  if (b->n_reads > 0 && (!b->read_bases || !b->read_q || !b->read_i || !b->read_d || !b->read_c || !b->rd_off || !b->rd_len))
    return set_error(FCS_PHMM_EINVAL, "null read plane");
  if (b->n_haps > 0 && (!b->hap_bases || !b->hp_off || !b->hp_len)) return set_error(FCS_PHMM_EINVAL, "null haplotype plane");
  for (int64_t g = 0; g < b->n_regions; ++g) {
    if (b->reg_nreads[g] < 0 || b->reg_nhaps[g] < 0) return set_error(FCS_PHMM_EINVAL, "negative read or haplotype count");
    if (b->reg_read0[g] < 0 || (int64_t)b->reg_read0[g] + b->reg_nreads[g] > b->n_reads)
      return set_error(FCS_PHMM_EINVAL, "region read range outside the batch");
    if (b->reg_hap0[g] < 0 || (int64_t)b->reg_hap0[g] + b->reg_nhaps[g] > b->n_haps)
      return set_error(FCS_PHMM_EINVAL, "region haplotype range outside the batch");
  }
  return FCS_PHMM_OK;
}

}  // namespace

#define API_TRY try {
#define API_CATCH                                                         \
  }                                                                       \
  catch (const std::bad_alloc&) {                                         \
    return set_error(FCS_PHMM_ENOMEM, "host allocation failed");          \
  }                                                                       \
  catch (const std::exception& ex) {                                      \
    return set_error(FCS_PHMM_EINVAL, std::string("internal: ") + ex.what()); \
  }                                                                       \
  catch (...) {                                                           \
    return set_error(FCS_PHMM_EINVAL, "internal: unknown exception");    \
  }

extern "C" {

int fcs_pairhmm_abi_version(void) { return FCS_PHMM_ABI_VERSION; }

int fcs_pairhmm_create(const fcs_phmm_config* cfg, fcs_phmm_handle** out) {
  if (!out) return set_error(FCS_PHMM_EINVAL, "null out pointer");
  *out = nullptr;
  API_TRY
  Engine* e = nullptr;
  int rc = Engine::create(cfg, &e);
  if (rc != FCS_PHMM_OK) return rc;
  fcs_phmm_handle* h = new fcs_phmm_handle();
  h->e = e;
  *out = h;
  return FCS_PHMM_OK;
  API_CATCH
}

void fcs_pairhmm_destroy(fcs_phmm_handle* h) {
  if (!h) return;
  try {
    delete h->e;
  } catch (...) {
  }
  delete h;
}

const char* fcs_pairhmm_last_error(const fcs_phmm_handle*) { return last_error(); }

int fcs_pairhmm_device_count(const fcs_phmm_handle* h) { return h ? h->e->device_count() : 0; }

int fcs_pairhmm_compute(fcs_phmm_handle* h, const fcs_phmm_region* regions, int32_t n_regions) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  API_TRY
  int rc = check_regions(regions, n_regions);
  if (rc != FCS_PHMM_OK) return rc;
  RegionInput in(regions, n_regions);
  return h->e->compute(in);
  API_CATCH
}

int fcs_pairhmm_compute_flat(fcs_phmm_handle* h, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64, float* raw_f32) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  API_TRY
  int rc = check_flat(b);
  if (rc != FCS_PHMM_OK) return rc;
  if (!out) return set_error(FCS_PHMM_EINVAL, "null out");
  std::unique_ptr<Input> in = make_flat_input(*b, out, used_fp64, raw_f32);
  return h->e->compute(*in);
  API_CATCH
}

int fcs_pairhmm_set_finalize(fcs_phmm_handle* h, const fcs_phmm_finalize_params* p) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  if (p && p->enabled && !(p->log10_global_mismapping_rate <= 0.0) ) return set_error(FCS_PHMM_EINVAL, "log10 mismapping rate must be <= 0");
  if (p && p->enabled) h->e->set_finalize(true, p->log10_global_mismapping_rate, p->expected_error_rate_per_base);
  else h->e->set_finalize(false, -4.5, 0.02);
  return FCS_PHMM_OK;
}

int fcs_pairhmm_compute_flat_finalized(fcs_phmm_handle* h, const fcs_phmm_flat_batch* b, double* out, uint8_t* used_fp64, uint8_t* poorly) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  API_TRY
  int rc = check_flat(b);
  if (rc != FCS_PHMM_OK) return rc;
  if (!out) return set_error(FCS_PHMM_EINVAL, "null out");
  std::unique_ptr<Input> in = make_flat_input(*b, out, used_fp64, nullptr, poorly);
  return h->e->compute(*in);
  API_CATCH
}

int fcs_pairhmm_submit(fcs_phmm_handle* h, const fcs_phmm_region* regions, int32_t n_regions, fcs_phmm_ticket* ticket) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  if (!ticket) return set_error(FCS_PHMM_EINVAL, "null ticket");
  API_TRY
  int rc = check_regions(regions, n_regions);
  if (rc != FCS_PHMM_OK) return rc;
  return h->e->submit(std::unique_ptr<Input>(new RegionInput(regions, n_regions)), nullptr, ticket);
  API_CATCH
}

int fcs_pairhmm_wait(fcs_phmm_handle* h, fcs_phmm_ticket ticket) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  API_TRY
  return h->e->wait(ticket);
  API_CATCH
}

int fcs_pairhmm_batch_create(fcs_phmm_handle* h, const fcs_phmm_flat_batch* b, int32_t device_index, fcs_phmm_batch** out) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  if (!out) return set_error(FCS_PHMM_EINVAL, "null out pointer");
  *out = nullptr;
  API_TRY
  int rc = check_flat(b);
  if (rc != FCS_PHMM_OK) return rc;
  Batch* bb = nullptr;
  rc = h->e->batch_create(b, device_index, &bb);
  if (rc != FCS_PHMM_OK) return rc;
  fcs_phmm_batch* w = new fcs_phmm_batch();
  w->b = bb;
  *out = w;
  return FCS_PHMM_OK;
  API_CATCH
}

int fcs_pairhmm_batch_run(fcs_phmm_handle* h, fcs_phmm_batch* b) {
  if (!h || !b) return set_error(FCS_PHMM_EINVAL, "null handle or batch");
  API_TRY
  return h->e->batch_run(b->b, false, nullptr, nullptr);
  API_CATCH
}

int fcs_pairhmm_batch_run_timed(fcs_phmm_handle* h, fcs_phmm_batch* b, float* total_ms, float* main_ms) {
  if (!h || !b) return set_error(FCS_PHMM_EINVAL, "null handle or batch");
  API_TRY
  return h->e->batch_run(b->b, true, total_ms, main_ms);
  API_CATCH
}

int fcs_pairhmm_batch_sync(fcs_phmm_handle* h, fcs_phmm_batch* b) {
  if (!h || !b) return set_error(FCS_PHMM_EINVAL, "null handle or batch");
  API_TRY
  return h->e->batch_sync(b->b);
  API_CATCH
}

int fcs_pairhmm_batch_download(fcs_phmm_handle* h, fcs_phmm_batch* b, double* out, uint8_t* used_fp64, float* raw_f32) {
  if (!h || !b) return set_error(FCS_PHMM_EINVAL, "null handle or batch");
  API_TRY
  return h->e->batch_download(b->b, out, used_fp64, raw_f32);
  API_CATCH
}

int64_t fcs_pairhmm_batch_pairs(const fcs_phmm_batch* b) { return b ? (int64_t)b->b->slot.plan.n_pairs : 0; }
int64_t fcs_pairhmm_batch_cells(const fcs_phmm_batch* b) { return b ? (int64_t)b->b->slot.plan.cells : 0; }
int32_t fcs_pairhmm_batch_launches(const fcs_phmm_batch* b) { return b ? b->b->slot.plan.launches() : 0; }

void fcs_pairhmm_batch_destroy(fcs_phmm_handle* h, fcs_phmm_batch* b) {
  if (!h || !b) return;
  try {
    h->e->batch_destroy(b->b);
  } catch (...) {
  }
  delete b;
}

int fcs_pairhmm_get_stats(fcs_phmm_handle* h, fcs_phmm_stats* out) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  return h->e->get_stats(out);
}

int fcs_pairhmm_reset_stats(fcs_phmm_handle* h) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  h->e->reset_stats();
  return FCS_PHMM_OK;
}

int fcs_pairhmm_set_capture(fcs_phmm_handle* h, const char* path) {
  if (!h) return set_error(FCS_PHMM_EINVAL, "null handle");
  API_TRY
  return h->e->set_capture(path);
  API_CATCH
}

int fcs_pairhmm_capture_load(const char* path, fcs_phmm_flat_batch* out, void** owner) {
  if (!path || !out || !owner) return set_error(FCS_PHMM_EINVAL, "null argument");
  *owner = nullptr;
  API_TRY
  LoadedCapture* c = nullptr;
  int rc = load_capture(path, &c);
  if (rc != FCS_PHMM_OK) return rc;
  c->view(out);
  *owner = c;
  return FCS_PHMM_OK;
  API_CATCH
}

int fcs_pairhmm_capture_parse(const void* blocks, uint64_t n_bytes, fcs_phmm_flat_batch* out, void** owner) {
  if (!blocks || !out || !owner) return set_error(FCS_PHMM_EINVAL, "null argument");
  *owner = nullptr;
  API_TRY
  std::unique_ptr<LoadedCapture> c(new LoadedCapture());
  if (!parse_blocks(static_cast<const uint8_t*>(blocks), (size_t)n_bytes, *c)) return set_error(FCS_PHMM_EINVAL, "truncated or corrupt capture blocks");
  c->view(out);
  *owner = c.release();
  return FCS_PHMM_OK;
  API_CATCH
}

void fcs_pairhmm_capture_free(void* owner) { delete static_cast<LoadedCapture*>(owner); }

int fcs_pairhmm_prepare_read(const uint8_t* bases, const uint8_t* raw_base_q, int32_t len, int32_t mapq, const uint8_t* bam_ins_q,
                             const uint8_t* bam_del_q, const fcs_phmm_prep_params* params, uint8_t* out_base_q, uint8_t* out_ins_q,
                             uint8_t* out_del_q, uint8_t* out_gcp) {
  API_TRY
  return prepare_read(bases, raw_base_q, len, mapq, bam_ins_q, bam_del_q, params, out_base_q, out_ins_q, out_del_q, out_gcp);
  API_CATCH
}

int fcs_pairhmm_finalize_region(double* log10_likelihoods, int32_t n_reads, int32_t n_haps, const int32_t* read_len,
                                double log10_global_mismapping_rate, double expected_error_rate_per_base, uint8_t* out_poorly_modeled) {
  API_TRY
  return finalize_region(log10_likelihoods, n_reads, n_haps, read_len, log10_global_mismapping_rate, expected_error_rate_per_base,
                         out_poorly_modeled);
  API_CATCH
}

int fcs_pairhmm_plan_check(const fcs_phmm_flat_batch* b, int32_t sm_count, fcs_phmm_plan_info* out) {
  if (!out) return set_error(FCS_PHMM_EINVAL, "null out");
  API_TRY
  int rc = check_flat(b);
  if (rc != FCS_PHMM_OK) return rc;
  return plan_check(b, sm_count, out);
  API_CATCH
}

float fcs_pairhmm_lut_ph2pr_f32(int q) { return luts().ph2pr_f[q & 127]; }
double fcs_pairhmm_lut_ph2pr_f64(int q) { return luts().ph2pr_d[q & 127]; }
float fcs_pairhmm_lut_mm_f32(int i, int d) { return luts().mm_f[mm_index(i & 127, d & 127)]; }
double fcs_pairhmm_lut_mm_f64(int i, int d) { return luts().mm_d[mm_index(i & 127, d & 127)]; }

int fcs_pairhmm_kernel_class(int32_t read_len, int32_t fp64, int32_t* lanes_per_read, int32_t* rows_per_lane) {
  const ClassRef* k = select_class(fp64 != 0, false, read_len);
  if (!k) return set_error(FCS_PHMM_EUNSUPPORTED, "no compiled kernel class covers this read length");
  if (lanes_per_read) *lanes_per_read = k->G;
  if (rows_per_lane) *rows_per_lane = k->R;
  return FCS_PHMM_OK;
}

}  // extern "C"
