/* Shim: SparseFrame_amd is dead code (call commented out, CH.c:1935). */
#pragma once
#define AMD_CONTROL 5
#define AMD_INFO 20
#define AMD_DENSE 0
#define AMD_AGGRESSIVE 1
static inline void amd_l2(long n, long *Pe, long *Iw, long *Len, long iwlen, long pfree, long *Nv, long *Next,
                          long *Last, long *Head, long *Elen, long *Degree, long *W, double *Control, double *Info) {}
