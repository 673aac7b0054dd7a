/* forwarding header kept for source compatibility with code that includes the reference's arch.h */
#include "SparseFrame.h"
