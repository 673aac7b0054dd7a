/* forwarding header kept for source compatibility with code that includes the reference's type.h */
#include "SparseFrame.h"
