/*
 * tests/jni_stub/fake_jvm.c — a tiny stand-in for the JVM side of the JNI shim (TEST INFRASTRUCTURE).
 *
 * The image has no JDK, so falcon-genome_b200/jni/fcs_pairhmm_jni.c can only be compiled against the stub
 * tests/jni_stub/jni.h.  This file implements that stub's JNIEnv function table over a minimal object model
 * (byte arrays, a double array, object arrays, and "holder" objects with the byte-array fields GATK's
 * ReadDataHolder / HaplotypeDataHolder carry [upstream]) and drives the shim's three GKL entry points the way
 * VectorLoglessPairHMM does: initNative once, computeLikelihoodsNative per region, doneNative at the end.
 * Like a copying JVM it hands out COPIES from Get<Type>ArrayElements and only writes a double array back
 * on Release with mode 0, so the test also sees whether the shim releases what it pins and commits the output.
 * Local references are counted: every object the "VM" returns (array elements, fields, classes) is one live
 * reference until DeleteLocalRef; a native frame is only guaranteed 16 (more after EnsureLocalCapacity), and the
 * test reads back the high-water mark.  Modes: 1 = the holder classes lack the expected fields (GetFieldID leaves
 * NoSuchFieldError pending), 2 = computeLikelihoodsNative is called although initNative was never called.
 *
 * It is not a JVM and proves nothing about ABI compatibility with one; it checks the shim's logic.
 */
#include <dlfcn.h>
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { K_CLASS, K_BYTES, K_DOUBLES, K_OBJECTS, K_HOLDER };
struct _jobject {
  int kind;
  jsize len;
  void* data;                 /* K_BYTES: jbyte[len]; K_DOUBLES: jdouble[len]; K_OBJECTS: jobject[len] */
  struct _jobject* field[5];  /* K_HOLDER */
  const char* name;           /* K_CLASS */
};
struct _jfieldID {
  int idx;
};

static struct _jfieldID g_fid[5] = {{0}, {1}, {2}, {3}, {4}};
static int g_pinned;           /* Get*ArrayElements minus Release*ArrayElements */
static int g_bad_release;      /* releases of pointers that were never handed out, or byte releases that would copy back */
static char g_exception[512];  /* pending exception: "<class>: <message>" */
static int g_live_refs, g_max_live_refs, g_ref_capacity = 16;
static int g_calls_with_pending; /* JNI calls that are illegal while an exception is pending */
static int g_break_fields;
static jobject new_ref(jobject o) {
  if (o && ++g_live_refs > g_max_live_refs) g_max_live_refs = g_live_refs;
  return o;
}
#define PENDING_GUARD() do { if (g_exception[0]) ++g_calls_with_pending; } while (0)

static jfieldID f_GetFieldID(JNIEnv* env, jclass cls, const char* name, const char* sig) {
  (void)env;
  PENDING_GUARD();
  static const char* read_fields[5] = {"readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP"};
  if (g_break_fields && strcmp(name, "insertionGOP") == 0) {
    snprintf(g_exception, sizeof(g_exception), "java/lang/NoSuchFieldError: %s", name);
    return NULL;
  }
  if (!cls || cls->kind != K_CLASS || strcmp(sig, "[B") != 0) return NULL;
  if (strcmp(cls->name, "ReadDataHolder") == 0) {
    for (int i = 0; i < 5; ++i)
      if (strcmp(name, read_fields[i]) == 0) return &g_fid[i];
  } else if (strcmp(cls->name, "HaplotypeDataHolder") == 0 && strcmp(name, "haplotypeBases") == 0) {
    return &g_fid[0];
  }
  return NULL;
}
static jobject f_GetObjectField(JNIEnv* env, jobject o, jfieldID f) {
  (void)env;
  PENDING_GUARD();
  return new_ref((o && f && o->kind == K_HOLDER) ? o->field[f->idx] : NULL);
}
static jsize f_GetArrayLength(JNIEnv* env, jarray a) {
  (void)env;
  return a ? a->len : 0;
}
static jobject f_GetObjectArrayElement(JNIEnv* env, jobjectArray a, jsize i) {
  (void)env;
  PENDING_GUARD();
  return new_ref((a && a->kind == K_OBJECTS && i >= 0 && i < a->len) ? ((jobject*)a->data)[i] : NULL);
}
static jbyte* f_GetByteArrayElements(JNIEnv* env, jbyteArray a, jboolean* is_copy) {
  (void)env;
  if (!a || a->kind != K_BYTES) return NULL;
  jbyte* c = (jbyte*)malloc((size_t)a->len + 1);
  memcpy(c, a->data, (size_t)a->len);
  if (is_copy) *is_copy = 1;
  ++g_pinned;
  return c;
}
static void f_ReleaseByteArrayElements(JNIEnv* env, jbyteArray a, jbyte* p, jint mode) {
  (void)env;
  if (!a || a->kind != K_BYTES || !p) { ++g_bad_release; return; }
  if (mode != JNI_ABORT) ++g_bad_release; /* the inputs are read-only: anything but JNI_ABORT copies them back for nothing */
  free(p);
  --g_pinned;
}
static jdouble* f_GetDoubleArrayElements(JNIEnv* env, jdoubleArray a, jboolean* is_copy) {
  (void)env;
  if (!a || a->kind != K_DOUBLES) return NULL;
  jdouble* c = (jdouble*)malloc(sizeof(jdouble) * ((size_t)a->len + 1));
  memcpy(c, a->data, sizeof(jdouble) * (size_t)a->len);
  if (is_copy) *is_copy = 1;
  ++g_pinned;
  return c;
}
static void f_ReleaseDoubleArrayElements(JNIEnv* env, jdoubleArray a, jdouble* p, jint mode) {
  (void)env;
  if (!a || a->kind != K_DOUBLES || !p) { ++g_bad_release; return; }
  if (mode != JNI_ABORT) memcpy(a->data, p, sizeof(jdouble) * (size_t)a->len); /* 0 = copy back and free */
  free(p);
  --g_pinned;
}
static struct _jobject g_exc_class = {K_CLASS, 0, NULL, {0}, NULL};
static jclass f_FindClass(JNIEnv* env, const char* name) {
  (void)env;
  PENDING_GUARD();
  g_exc_class.name = name;
  return new_ref(&g_exc_class);
}
static jint f_ThrowNew(JNIEnv* env, jclass cls, const char* msg) {
  (void)env;
  PENDING_GUARD();
  snprintf(g_exception, sizeof(g_exception), "%s: %s", cls && cls->name ? cls->name : "?", msg ? msg : "");
  return 0;
}
static jboolean f_ExceptionCheck(JNIEnv* env) {
  (void)env;
  return g_exception[0] != 0;
}
static void f_DeleteLocalRef(JNIEnv* env, jobject o) {
  (void)env;
  if (o) --g_live_refs;
}
static jint f_EnsureLocalCapacity(JNIEnv* env, jint n) {
  (void)env;
  if (n > g_ref_capacity) g_ref_capacity = n;
  return 0;
}
static void f_GetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize start, jsize len, jbyte* buf) {
  (void)env;
  PENDING_GUARD();
  if (!a || a->kind != K_BYTES || start < 0 || len < 0 || start + len > a->len) {
    snprintf(g_exception, sizeof(g_exception), "java/lang/ArrayIndexOutOfBoundsException: GetByteArrayRegion");
    return;
  }
  memcpy(buf, (const jbyte*)a->data + start, (size_t)len);
}
static void f_SetDoubleArrayRegion(JNIEnv* env, jdoubleArray a, jsize start, jsize len, const jdouble* buf) {
  (void)env;
  PENDING_GUARD();
  if (!a || a->kind != K_DOUBLES || start < 0 || len < 0 || start + len > a->len) {
    snprintf(g_exception, sizeof(g_exception), "java/lang/ArrayIndexOutOfBoundsException: SetDoubleArrayRegion");
    return;
  }
  memcpy((jdouble*)a->data + start, buf, sizeof(jdouble) * (size_t)len);
}

static const struct JNINativeInterface_ g_table = {
    f_GetFieldID,          f_GetObjectField,          f_GetArrayLength,           f_GetObjectArrayElement, f_GetByteArrayElements,
    f_ReleaseByteArrayElements, f_GetDoubleArrayElements, f_ReleaseDoubleArrayElements, f_FindClass,             f_ThrowNew,
    f_ExceptionCheck,      f_DeleteLocalRef,          f_EnsureLocalCapacity,      f_GetByteArrayRegion,    f_SetDoubleArrayRegion};

typedef void (*init_fn)(JNIEnv*, jclass, jclass, jclass, jboolean, jint);
typedef void (*compute_fn)(JNIEnv*, jobject, jobjectArray, jobjectArray, jdoubleArray);
typedef void (*done_fn)(JNIEnv*, jobject);

static jobject new_bytes(const uint8_t* p, int32_t len) {
  jobject o = (jobject)calloc(1, sizeof(*o));
  o->kind = K_BYTES;
  o->len = len;
  o->data = malloc((size_t)len + 1);
  memcpy(o->data, p, (size_t)len);
  return o;
}
static void free_obj(jobject o) {
  if (!o) return;
  free(o->data);
  free(o);
}

/* test knobs / read-backs */
__attribute__((visibility("default"))) void fake_jvm_set_mode(int break_fields) { g_break_fields = break_fields; }
__attribute__((visibility("default"))) int fake_jvm_max_live_refs(void) { return g_max_live_refs; }
__attribute__((visibility("default"))) int fake_jvm_ref_capacity(void) { return g_ref_capacity; }
__attribute__((visibility("default"))) int fake_jvm_calls_with_pending_exception(void) { return g_calls_with_pending; }

/*
 * Runs one session against the shim at `shim_path`: initNative, `repeats` x computeLikelihoodsNative on the
 * region given as concatenated planes (read r = rd_len[r] bytes at rd_off[r] of each plane; hap h likewise),
 * doneNative.  out[n_reads * n_haps] receives what the "Java" double array holds afterwards.
 * Returns 0, or -1 with the reason in err (pending exception, unbalanced pins, missing symbol).
 */
__attribute__((visibility("default"))) int fake_jvm_run(const char* shim_path, const uint8_t* bases, const uint8_t* q, const uint8_t* ins,
                                                        const uint8_t* del, const uint8_t* gcp, const int64_t* rd_off, const int32_t* rd_len,
                                                        int32_t n_reads, const uint8_t* hap_bases, const int64_t* hp_off,
                                                        const int32_t* hp_len, int32_t n_haps, int32_t use_double, int32_t max_threads,
                                                        int32_t repeats, double* out, char* err, int32_t err_cap) {
  g_pinned = g_bad_release = 0;
  g_live_refs = g_max_live_refs = g_calls_with_pending = 0;
  g_ref_capacity = 16;
  g_exception[0] = 0;
  if (err_cap > 0) err[0] = 0;
  void* so = dlopen(shim_path, RTLD_NOW | RTLD_LOCAL);
  if (!so) {
    snprintf(err, (size_t)err_cap, "dlopen: %s", dlerror());
    return -1;
  }
  init_fn init = (init_fn)dlsym(so, "Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative");
  compute_fn compute = (compute_fn)dlsym(so, "Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative");
  done_fn done = (done_fn)dlsym(so, "Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative");
  if (!init || !compute || !done) {
    snprintf(err, (size_t)err_cap, "the shim does not export the three GKL entry points");
    dlclose(so);
    return -1;
  }
  const struct JNINativeInterface_* envp = &g_table;
  JNIEnv* env = &envp;
  struct _jobject read_cls = {K_CLASS, 0, NULL, {0}, "ReadDataHolder"}, hap_cls = {K_CLASS, 0, NULL, {0}, "HaplotypeDataHolder"};
  struct _jobject self = {K_HOLDER, 0, NULL, {0}, NULL};
  int rc = 0;
  const int skip_init = max_threads < 0; /* test mode: computeLikelihoodsNative without a successful initNative */
  if (!skip_init) init(env, &read_cls, &read_cls, &hap_cls, (jboolean)(use_double != 0), max_threads);
  if (g_exception[0]) {
    snprintf(err, (size_t)err_cap, "%s", g_exception);
    dlclose(so);
    return -1;
  }
  /* the "Java" objects of one region */
  struct _jobject reads = {K_OBJECTS, n_reads, calloc((size_t)n_reads + 1, sizeof(jobject)), {0}, NULL};
  struct _jobject haps = {K_OBJECTS, n_haps, calloc((size_t)n_haps + 1, sizeof(jobject)), {0}, NULL};
  struct _jobject outa = {K_DOUBLES, n_reads * n_haps, calloc((size_t)n_reads * (size_t)n_haps + 1, sizeof(jdouble)), {0}, NULL};
  const uint8_t* planes[5] = {bases, q, ins, del, gcp};
  for (int32_t r = 0; r < n_reads; ++r) {
    jobject h = (jobject)calloc(1, sizeof(*h));
    h->kind = K_HOLDER;
    for (int k = 0; k < 5; ++k) h->field[k] = new_bytes(planes[k] + rd_off[r], rd_len[r]);
    ((jobject*)reads.data)[r] = h;
  }
  for (int32_t j = 0; j < n_haps; ++j) {
    jobject h = (jobject)calloc(1, sizeof(*h));
    h->kind = K_HOLDER;
    h->field[0] = new_bytes(hap_bases + hp_off[j], hp_len[j]);
    ((jobject*)haps.data)[j] = h;
  }
  for (int32_t it = 0; it < repeats && !g_exception[0]; ++it) {
    for (jsize i = 0; i < outa.len; ++i) ((jdouble*)outa.data)[i] = 1.0; /* a log10 likelihood is never positive */
    compute(env, &self, &reads, &haps, &outa);
  }
  memcpy(out, outa.data, sizeof(double) * (size_t)outa.len);
  done(env, &self);
  if (g_exception[0]) {
    snprintf(err, (size_t)err_cap, "%s", g_exception);
    rc = -1;
  } else if (g_pinned != 0 || g_bad_release != 0) {
    snprintf(err, (size_t)err_cap, "array elements not released properly: %d still pinned, %d bad releases", g_pinned, g_bad_release);
    rc = -1;
  } else if (g_live_refs != 0 || g_max_live_refs > g_ref_capacity) {
    snprintf(err, (size_t)err_cap, "local references: %d leaked, high-water mark %d of a capacity of %d", g_live_refs, g_max_live_refs, g_ref_capacity);
    rc = -1;
  } else if (g_calls_with_pending) {
    snprintf(err, (size_t)err_cap, "%d JNI calls were made while an exception was pending", g_calls_with_pending);
    rc = -1;
  }
  for (int32_t r = 0; r < n_reads; ++r) {
    jobject h = ((jobject*)reads.data)[r];
    for (int k = 0; k < 5; ++k) free_obj(h->field[k]);
    free(h);
  }
  for (int32_t j = 0; j < n_haps; ++j) {
    jobject h = ((jobject*)haps.data)[j];
    free_obj(h->field[0]);
    free(h);
  }
  free(reads.data);
  free(haps.data);
  free(outa.data);
  dlclose(so);
  return rc;
}
