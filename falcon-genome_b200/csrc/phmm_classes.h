// phmm_classes.h — the (G, R) classes compiled into the library, one list per translation
// unit so the build parallelises.  A read of length L needs G*R >= L+1 rows.
//   FP32 general: R in 4..24 (8 registers per row); FP32 uniform-GCP: R up to 28 (7 per row);
//   FP64: R in 4..12 (16 / 14 registers per row).
#pragma once
#define PHMM_F32_G4(X)  X(4,4) X(4,6) X(4,8) X(4,10) X(4,12) X(4,14) X(4,16) X(4,18) X(4,20) X(4,22) X(4,24)
#define PHMM_F32_G8(X)  X(8,13) X(8,14) X(8,15) X(8,16) X(8,17) X(8,18) X(8,19) X(8,20) X(8,21) X(8,22) X(8,23) X(8,24)
#define PHMM_F32_G16(X) X(16,13) X(16,14) X(16,15) X(16,16) X(16,17) X(16,18) X(16,19) X(16,20) X(16,21) X(16,22) X(16,23) X(16,24)
#define PHMM_F32_G32(X) X(32,13) X(32,14) X(32,15) X(32,16) X(32,17) X(32,18) X(32,19) X(32,20) X(32,21) X(32,22) X(32,23) X(32,24)
#define PHMM_F32U_G4(X)  PHMM_F32_G4(X)
#define PHMM_F32U_G8(X)  PHMM_F32_G8(X)
#define PHMM_F32U_G16(X) PHMM_F32_G16(X)
#define PHMM_F32U_G32(X) PHMM_F32_G32(X)
#define PHMM_F64_G4(X)  X(4,4) X(4,6) X(4,8) X(4,10) X(4,12)
#define PHMM_F64_G8(X)  X(8,7) X(8,8) X(8,9) X(8,10) X(8,11) X(8,12)
#define PHMM_F64_G16(X) X(16,7) X(16,8) X(16,9) X(16,10) X(16,11) X(16,12)
#define PHMM_F64_G32(X) X(32,7) X(32,8) X(32,9) X(32,10) X(32,11) X(32,12)
