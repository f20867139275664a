/*
 * Drop-in replacement for the reference's src/include/hcore.h:30-44: the HCORE_enum type
 * and the operation constants (numerically equal to the CBLAS enums).  Unlike the
 * reference header this one does not pull in <cblas.h>/<lapacke.h>: the B200 library
 * has no BLAS/LAPACK dependency.
 */
#ifndef HCORE_B200_DROPIN_HCORE_H
#define HCORE_B200_DROPIN_HCORE_H

typedef int HCORE_enum;

enum {
    HCORE_NoTrans = 111, HCORE_Trans = 112, HCORE_ConjTrans = 113,
    HCORE_Upper = 121, HCORE_Lower = 122, HCORE_UpperLower = 123,
    HCORE_NonUnit = 131, HCORE_Unit = 132,
    HCORE_Left = 141, HCORE_Right = 142
};

#ifndef hcore_min
#define hcore_min(x, y) (((x) < (y)) ? (x) : (y))
#endif
#ifndef hcore_max
#define hcore_max(x, y) (((x) < (y)) ? (y) : (x))
#endif
#endif
