/* forwarding header kept for source compatibility with code that includes the reference's info.h */
#include "SparseFrame.h"
