/* Cholesky/Include/type.h -- enumerations stored in matrix_info_struct (reference: Include/type.h). */
#ifndef INCLUDE_TYPE_H
#define INCLUDE_TYPE_H

enum FactorizeType { TYPE_CHOLESKY, TYPE_QR, TYPE_LU };
enum PermMethod { PERM_IDENTITY, PERM_AMD, PERM_METIS };

#endif
