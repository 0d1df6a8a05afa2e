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
 *       falcon-genome_b200/jni/fcs_pairhmm_jni.c -L falcon-genome_b200 -lfcs_pairhmm -o libgkl_pairhmm.so
 *
 * Field names follow GATK's ReadDataHolder {readBases, readQuals, insertionGOP, deletionGOP, overallGCP}
 * and HaplotypeDataHolder {haplotypeBases} [upstream].
 */
#include <jni.h>
#include <stdlib.h>

#include "fcs_pairhmm.h"

static fcs_phmm_handle* g_handle;
static jfieldID g_readBases, g_readQuals, g_insGOP, g_delGOP, g_gcp, g_hapBases;

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_initNative(JNIEnv* env, jclass cls, jclass readDataHolder,
                                                                           jclass hapDataHolder, jboolean use_double,
                                                                           jint max_threads) {
  (void)cls;
  g_readBases = (*env)->GetFieldID(env, readDataHolder, "readBases", "[B");
  g_readQuals = (*env)->GetFieldID(env, readDataHolder, "readQuals", "[B");
  g_insGOP = (*env)->GetFieldID(env, readDataHolder, "insertionGOP", "[B");
  g_delGOP = (*env)->GetFieldID(env, readDataHolder, "deletionGOP", "[B");
  g_gcp = (*env)->GetFieldID(env, readDataHolder, "overallGCP", "[B");
  g_hapBases = (*env)->GetFieldID(env, hapDataHolder, "haplotypeBases", "[B");
  fcs_phmm_config cfg = {0};
  cfg.struct_size = sizeof(cfg);
  cfg.use_double = use_double ? 1 : 0;
  cfg.max_threads = max_threads;
  if (!g_handle && fcs_pairhmm_create(&cfg, &g_handle) != FCS_PHMM_OK) {
    /* no CPU fallback: surface the failure to the JVM, as a missing NAM is fatal in the reference */
    jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
    (*env)->ThrowNew(env, ex, fcs_pairhmm_last_error(NULL));
  }
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_computeLikelihoodsNative(JNIEnv* env, jobject obj,
                                                                                         jobjectArray reads,
                                                                                         jobjectArray haps,
                                                                                         jdoubleArray out) {
  (void)obj;
  const jsize nr = (*env)->GetArrayLength(env, reads), nh = (*env)->GetArrayLength(env, haps);
  fcs_phmm_read* R = (fcs_phmm_read*)calloc((size_t)nr, sizeof(*R));
  fcs_phmm_hap* H = (fcs_phmm_hap*)calloc((size_t)nh, sizeof(*H));
  jbyteArray* ra = (jbyteArray*)calloc((size_t)nr * 5 + (size_t)nh, sizeof(jbyteArray));
  for (jsize r = 0; r < nr; ++r) {
    jobject o = (*env)->GetObjectArrayElement(env, reads, r);
    jfieldID f[5] = {g_readBases, g_readQuals, g_insGOP, g_delGOP, g_gcp};
    const uint8_t** dst[5] = {&R[r].bases, &R[r].base_q, &R[r].ins_q, &R[r].del_q, &R[r].gcp};
    for (int k = 0; k < 5; ++k) {
      ra[r * 5 + k] = (jbyteArray)(*env)->GetObjectField(env, o, f[k]);
      *dst[k] = (const uint8_t*)(*env)->GetByteArrayElements(env, ra[r * 5 + k], NULL);
    }
    R[r].len = (*env)->GetArrayLength(env, ra[r * 5]);
  }
  for (jsize h = 0; h < nh; ++h) {
    jobject o = (*env)->GetObjectArrayElement(env, haps, h);
    ra[nr * 5 + h] = (jbyteArray)(*env)->GetObjectField(env, o, g_hapBases);
    H[h].bases = (const uint8_t*)(*env)->GetByteArrayElements(env, ra[nr * 5 + h], NULL);
    H[h].len = (*env)->GetArrayLength(env, ra[nr * 5 + h]);
  }
  jdouble* o = (*env)->GetDoubleArrayElements(env, out, NULL);
  fcs_phmm_region reg = {R, nr, H, nh, (double*)o, NULL};
  const int rc = fcs_pairhmm_compute(g_handle, &reg, 1);
  (*env)->ReleaseDoubleArrayElements(env, out, o, 0);
  for (jsize r = 0; r < nr; ++r) {
    const uint8_t* src[5] = {R[r].bases, R[r].base_q, R[r].ins_q, R[r].del_q, R[r].gcp};
    for (int k = 0; k < 5; ++k) (*env)->ReleaseByteArrayElements(env, ra[r * 5 + k], (jbyte*)src[k], JNI_ABORT);
  }
  for (jsize h = 0; h < nh; ++h) (*env)->ReleaseByteArrayElements(env, ra[nr * 5 + h], (jbyte*)H[h].bases, JNI_ABORT);
  free(R); free(H); free(ra);
  if (rc != FCS_PHMM_OK) {
    jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
    (*env)->ThrowNew(env, ex, fcs_pairhmm_last_error(g_handle));
  }
}

JNIEXPORT void JNICALL Java_com_intel_gkl_pairhmm_IntelPairHmm_doneNative(JNIEnv* env, jobject obj) {
  (void)env; (void)obj;
  fcs_pairhmm_destroy(g_handle);
  g_handle = NULL;
}
