/* forwarding header kept for source compatibility with code that includes the reference's macro.h */
#include "SparseFrame.h"
