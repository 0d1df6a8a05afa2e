// tools/sass/one_class.cu — ONE (G, R, FORM) class of the FP32 wavefront kernel as a kernel of its own, built with the
// product's flags, so that its SASS loop body can be listed and counted (tools/sass_audit.py).  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I falcon-genome_b200/csrc -DGG=8 -DRR=19 -DFORM_=1 -DMINB=12 -cubin ...
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
#ifndef FORM_
#define FORM_ 1
#endif
#ifndef TT
#define TT float
#endif
#define LIST1(X) X(0, GG, RR)
#if defined(PAIR_KERNEL)
PHMM_DEFINE_PAIR_KERNEL(one_class_kernel, float, 3, MINB, LIST1)
#elif defined(QUEUE_KERNEL)
PHMM_DEFINE_QUEUE_KERNEL(one_class_kernel, TT, FORM_, MINB, LIST1)
#else
PHMM_DEFINE_TASK_KERNEL(one_class_kernel, TT, FORM_, MINB, LIST1)
#endif
}  // namespace fcsphmm
