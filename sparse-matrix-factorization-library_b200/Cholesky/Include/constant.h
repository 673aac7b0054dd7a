/* Cholesky/Include/constant.h -- BLAS scalar constants as (re, im) pairs (reference: Include/constant.h:6-8).
 * static: the reference defines them non-static, which forces library and executable into separate link units. */
#ifndef INCLUDE_CONSTANT_H
#define INCLUDE_CONSTANT_H

#include "arch.h"

static const Float zero[] = { 0, 0 };
static const Float one[] = { 1, 0 };
static const Float minus_one[] = { -1, 0 };

#endif
