/* libmetis_static.a (CUDA 12.9) references the pre-glibc-2.33 __xstat64 entry point. */
#include <sys/stat.h>
int __xstat64 ( int ver, const char *path, struct stat *buf ) { (void) ver; return stat ( path, buf ); }
