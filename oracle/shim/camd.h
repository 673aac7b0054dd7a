/* Shim: SparseFrame_camd is dead code (call commented out, CH.c:1936). */
#pragma once
static inline void camd_l2(long n, long *Pe, long *Iw, long *Len, long iwlen, long pfree, long *Nv, long *Next,
                           long *Last, long *Head, long *Elen, long *Degree, long *W, double *Control, double *Info,
                           const long *C, long *BucketSet) {}
