/*
 * fcs_pairhmm_jni.c — JNI shim that gives libfcs_pairhmm.so the symbol names GATK's
 * VectorLoglessPairHMM binds through Intel GKL's com.intel.gkl.pairhmm.IntelPairHmm
 * [upstream; SURVEY.md §8(b), §8(f) row f1]:
 *
 *   Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative(readDataHolderClass, haplotypeDataHolderClass,
 *                                                      use_double, max_threads)
 *   Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative(Object[] reads, Object[] haps, double[] out)
 *   Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative()
 *
 * The JVM that /root/reference/src/workers/HTCWorker.cpp:51-58 (and Mutect2Worker.cpp:113-121)
 * launches loads this as libgkl_pairhmm.so; every call is a 1:1 adapter onto the C ABI in
 * include/fcs_pairhmm.h.  NOT COMPILED IN THIS REPO'S BUILD: the image has no JDK (no jni.h).
 * Build where a JDK exists:
 *   gcc -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I include \
 *       falcon-genome_b200/jni/fcs_pairhmm_jni.c -L falcon-genome_b200 -lfcs_pairhmm -lpthread -o libgkl_pairhmm.so
 *
 * Field names follow GATK's ReadDataHolder {readBases, readQuals, insertionGOP, deletionGOP, overallGCP}
 * and HaplotypeDataHolder {haplotypeBases} [upstream].
 *
 * How a region crosses the boundary.  The byte arrays are copied with GetByteArrayRegion straight into the
 * planes of a one-region fcs_phmm_flat_batch (thread-local, grown on demand) and every local reference is
 * deleted as soon as its bytes are out: at most six references are alive at any time, however deep the pileup
 * (a Mutect2 region holds up to ~2000 reads = 12 000 references otherwise, far beyond the 16 a native frame
 * is guaranteed).  HotSpot copies on Get<Type>ArrayElements anyway, so this is one copy, not an extra one, and
 * nothing stays pinned across the GPU call.  The result goes back with one SetDoubleArrayRegion.
 */
#include <jni.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "fcs_pairhmm.h"

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static fcs_phmm_handle* g_handle; /* shared by every IntelPairHmm instance of the JVM */
static int g_users;               /* initNative calls not yet matched by doneNative */
static jfieldID g_readBases, g_readQuals, g_insGOP, g_delGOP, g_gcp, g_hapBases;

static void throw_new(JNIEnv* env, const char* cls_name, const char* msg) {
  if ((*env)->ExceptionCheck(env)) return; /* keep the first exception */
  jclass ex = (*env)->FindClass(env, cls_name);
  if (ex) {
    (*env)->ThrowNew(env, ex, msg);
    (*env)->DeleteLocalRef(env, ex);
  }
}

/* GetFieldID + check: a renamed holder field leaves NoSuchFieldError pending and returns NULL */
static int field_id(JNIEnv* env, jclass cls, const char* name, jfieldID* out) {
  *out = (*env)->GetFieldID(env, cls, name, "[B");
  return *out != NULL && !(*env)->ExceptionCheck(env);
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative(JNIEnv* env, jclass cls, jclass readDataHolder,
                                                                           jclass hapDataHolder, jboolean use_double,
                                                                           jint max_threads) {
  (void)cls;
  jfieldID rb, rq, ri, rd, rc, hb;
  if (!field_id(env, readDataHolder, "readBases", &rb) || !field_id(env, readDataHolder, "readQuals", &rq) ||
      !field_id(env, readDataHolder, "insertionGOP", &ri) || !field_id(env, readDataHolder, "deletionGOP", &rd) ||
      !field_id(env, readDataHolder, "overallGCP", &rc) || !field_id(env, hapDataHolder, "haplotypeBases", &hb)) {
    /* the JVM's NoSuchFieldError stays pending; make sure SOMETHING is thrown even if the VM returned NULL silently */
    throw_new(env, "java/lang/NoSuchFieldError", "ReadDataHolder / HaplotypeDataHolder do not have the fields libfcs_pairhmm expects");
    return;
  }
  pthread_mutex_lock(&g_mu);
  g_readBases = rb; g_readQuals = rq; g_insGOP = ri; g_delGOP = rd; g_gcp = rc; g_hapBases = hb;
  int rcode = FCS_PHMM_OK;
  if (!g_handle) {
    fcs_phmm_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.use_double = use_double ? 1 : 0;
    cfg.max_threads = max_threads;
    rcode = fcs_pairhmm_create(&cfg, &g_handle);
  }
  if (rcode == FCS_PHMM_OK) ++g_users;
  pthread_mutex_unlock(&g_mu);
  if (rcode != FCS_PHMM_OK) /* no CPU fallback: surface the failure to the JVM, as a missing NAM is fatal in the reference */
    throw_new(env, "java/lang/RuntimeException", fcs_pairhmm_last_error(NULL));
}

/* thread-local staging of one region (GATK may drive several IntelPairHmm instances from several threads) */
typedef struct {
  uint8_t* plane[6]; /* read bases, base quals, insertion, deletion, gcp; haplotype bases */
  size_t cap[6];
  int64_t *rd_off, *hp_off;
  int32_t *rd_len, *hp_len;
  size_t cap_reads, cap_haps;
  double* out;
  size_t cap_out;
} Staging;
static __thread Staging t_st;

static int grow(void** p, size_t* cap, size_t need, size_t elem) {
  if (need <= *cap) return 1;
  size_t n = *cap ? *cap : 1024;
  while (n < need) n *= 2;
  void* q = realloc(*p, n * elem);
  if (!q) return 0;
  *p = q;
  *cap = n;
  return 1;
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative(JNIEnv* env, jobject obj,
                                                                                         jobjectArray reads,
                                                                                         jobjectArray haps,
                                                                                         jdoubleArray out) {
  (void)obj;
  /* the handle first: after a failed initNative nothing is touched, copied or pinned */
  pthread_mutex_lock(&g_mu);
  fcs_phmm_handle* h = g_handle;
  pthread_mutex_unlock(&g_mu);
  if (!h) {
    throw_new(env, "java/lang/IllegalStateException", "libfcs_pairhmm: initNative failed or was not called (no handle)");
    return;
  }
  if (!reads || !haps || !out) {
    throw_new(env, "java/lang/NullPointerException", "libfcs_pairhmm: null array argument");
    return;
  }
  if ((*env)->EnsureLocalCapacity(env, 16) != 0) return; /* OutOfMemoryError pending */
  Staging* st = &t_st;
  const jsize nr = (*env)->GetArrayLength(env, reads), nh = (*env)->GetArrayLength(env, haps);
  const size_t n_out = (size_t)nr * (size_t)nh;
  if ((size_t)(*env)->GetArrayLength(env, out) < n_out) {
    throw_new(env, "java/lang/IllegalArgumentException", "libfcs_pairhmm: likelihood array shorter than reads x haplotypes");
    return;
  }
  void *p_ro = st->rd_off, *p_rl = st->rd_len, *p_ho = st->hp_off, *p_hl = st->hp_len, *p_out = st->out;
  size_t c1 = st->cap_reads, c2 = st->cap_reads, c3 = st->cap_haps, c4 = st->cap_haps;
  int ok = grow(&p_ro, &c1, (size_t)nr + 1, sizeof(int64_t)) && grow(&p_rl, &c2, (size_t)nr + 1, sizeof(int32_t)) &&
           grow(&p_ho, &c3, (size_t)nh + 1, sizeof(int64_t)) && grow(&p_hl, &c4, (size_t)nh + 1, sizeof(int32_t)) &&
           grow(&p_out, &st->cap_out, n_out + 1, sizeof(double));
  st->rd_off = (int64_t*)p_ro; st->rd_len = (int32_t*)p_rl; st->hp_off = (int64_t*)p_ho; st->hp_len = (int32_t*)p_hl; st->out = (double*)p_out;
  st->cap_reads = c1 < c2 ? c1 : c2;
  st->cap_haps = c3 < c4 ? c3 : c4;
  size_t rpos = 0, hpos = 0;
  for (jsize r = 0; ok && r < nr; ++r) {
    jobject o = (*env)->GetObjectArrayElement(env, reads, r);
    if (!o) { ok = 0; throw_new(env, "java/lang/NullPointerException", "libfcs_pairhmm: null read holder"); break; }
    const jfieldID f[5] = {g_readBases, g_readQuals, g_insGOP, g_delGOP, g_gcp};
    jsize len = 0;
    for (int k = 0; ok && k < 5; ++k) {
      jbyteArray a = (jbyteArray)(*env)->GetObjectField(env, o, f[k]);
      if (!a) { ok = 0; throw_new(env, "java/lang/NullPointerException", "libfcs_pairhmm: null byte array in a read holder"); break; }
      const jsize l = (*env)->GetArrayLength(env, a);
      if (k == 0) len = l;
      if (l < len) { ok = 0; throw_new(env, "java/lang/IllegalArgumentException", "libfcs_pairhmm: quality array shorter than the read"); }
      void* pl = st->plane[k];
      if (ok && !grow(&pl, &st->cap[k], rpos + (size_t)len + 1, 1)) ok = 0;
      st->plane[k] = (uint8_t*)pl;
      if (ok) (*env)->GetByteArrayRegion(env, a, 0, len, (jbyte*)(st->plane[k] + rpos));
      (*env)->DeleteLocalRef(env, a);
      if (ok && (*env)->ExceptionCheck(env)) ok = 0;
    }
    (*env)->DeleteLocalRef(env, o);
    st->rd_off[r] = (int64_t)rpos;
    st->rd_len[r] = len;
    rpos += (size_t)len;
  }
  for (jsize j = 0; ok && j < nh; ++j) {
    jobject o = (*env)->GetObjectArrayElement(env, haps, j);
    if (!o) { ok = 0; throw_new(env, "java/lang/NullPointerException", "libfcs_pairhmm: null haplotype holder"); break; }
    jbyteArray a = (jbyteArray)(*env)->GetObjectField(env, o, g_hapBases);
    if (!a) { ok = 0; throw_new(env, "java/lang/NullPointerException", "libfcs_pairhmm: null haplotype bases"); }
    if (ok) {
      const jsize len = (*env)->GetArrayLength(env, a);
      void* pl = st->plane[5];
      if (!grow(&pl, &st->cap[5], hpos + (size_t)len + 1, 1)) ok = 0;
      st->plane[5] = (uint8_t*)pl;
      if (ok) (*env)->GetByteArrayRegion(env, a, 0, len, (jbyte*)(st->plane[5] + hpos));
      st->hp_off[j] = (int64_t)hpos;
      st->hp_len[j] = len;
      hpos += (size_t)len;
      if (ok && (*env)->ExceptionCheck(env)) ok = 0;
    }
    if (a) (*env)->DeleteLocalRef(env, a);
    (*env)->DeleteLocalRef(env, o);
  }
  if (!ok) {
    throw_new(env, "java/lang/OutOfMemoryError", "libfcs_pairhmm: cannot stage the region");
    return;
  }
  if (n_out == 0) return;
  const int32_t zero32 = 0, nr32 = nr, nh32 = nh;
  const int64_t zero64 = 0;
  fcs_phmm_flat_batch b;
  memset(&b, 0, sizeof(b));
  b.read_bases = st->plane[0]; b.read_q = st->plane[1]; b.read_i = st->plane[2]; b.read_d = st->plane[3]; b.read_c = st->plane[4];
  b.rd_off = st->rd_off; b.rd_len = st->rd_len; b.n_reads = nr;
  b.hap_bases = st->plane[5]; b.hp_off = st->hp_off; b.hp_len = st->hp_len; b.n_haps = nh;
  b.reg_read0 = &zero32; b.reg_nreads = &nr32; b.reg_hap0 = &zero32; b.reg_nhaps = &nh32; b.reg_out0 = &zero64; b.n_regions = 1;
  const int rc = fcs_pairhmm_compute_flat(h, &b, st->out, NULL, NULL);
  if (rc != FCS_PHMM_OK) {
    throw_new(env, "java/lang/RuntimeException", fcs_pairhmm_last_error(h));
    return;
  }
  (*env)->SetDoubleArrayRegion(env, out, 0, (jsize)n_out, st->out);
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative(JNIEnv* env, jobject obj) {
  (void)env; (void)obj;
  pthread_mutex_lock(&g_mu);
  fcs_phmm_handle* h = NULL;
  if (g_users > 0 && --g_users == 0) { h = g_handle; g_handle = NULL; }
  pthread_mutex_unlock(&g_mu);
  if (h) fcs_pairhmm_destroy(h);
}
