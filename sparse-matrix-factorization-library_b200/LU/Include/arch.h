/* LU/Include/arch.h -- LU flavour of the shared header (the reference keeps two edited copies of every header; the
 * differences -- ten more matrix_info members, getrf prototypes, mappedSubtract's signature -- are selected by SPARSEFRAME_LU) */
#ifndef SPARSEFRAME_LU
#define SPARSEFRAME_LU 1
#endif
#include "../../Cholesky/Include/arch.h"
