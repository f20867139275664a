/*
 * ABI mirror of the reference's flop accumulator.  Every legacy kernel signature carries
 * a `flop_counter *` (reference src/misc/include/flop_util_structs.h:5-36) and the
 * kernels add their operation counts to `update` (HCORE_dgemm, reference
 * src/compute/hcore_dgemm.c:126,179,208,224,434,493,595,666) or `syrk` (HCORE_dsyrk,
 * reference src/compute/hcore_dsyrk.c:188).  Field order, names and types must match the
 * reference exactly so that existing callers (HiCMA codelets) link and read the same
 * offsets; the layout is checked by tests/test_abi.py against the compiled reference.
 */
#ifndef HCORE_B200_FLOPS_H
#define HCORE_B200_FLOPS_H
#include <stdint.h>

#define HCORE_B200_FLOP_FIELDS(X)                                                          \
    X(unsigned long int, __pad1)      X(unsigned long int, __pad2)                         \
    X(unsigned long int, copy)        X(unsigned long int, flop)                           \
    X(int, tid)                                                                            \
    X(unsigned long int, potrf)       X(unsigned long int, trsm)                           \
    X(unsigned long int, syrk)        X(unsigned long int, update)                         \
    X(unsigned long int, geqrfu)      X(unsigned long int, geqrfv)                         \
    X(unsigned long int, geqrfv_gemm) X(unsigned long int, gesvd)                          \
    X(unsigned long int, gesvd_gemm)  X(unsigned long int, gesvd_lasetu)                   \
    X(unsigned long int, gesvd_lasetv) X(unsigned long int, gesvd_lacpyu)                  \
    X(unsigned long int, gesvd_lacpyv) X(unsigned long int, gesvd_dscal)                   \
    X(unsigned long int, ormqru)      X(unsigned long int, ormqru_laset)                   \
    X(unsigned long int, ormqru_lacpy) X(unsigned long int, ormqrv)                        \
    X(unsigned long int, ormqrv_laset) X(unsigned long int, ormqrv_lacpy)                  \
    X(unsigned long int, omatcopy)                                                         \
    X(int, numthreads)                                                                     \
    X(unsigned long int, __pad3)      X(unsigned long int, __pad4)

#define HCORE_B200_DECL_FIELD(type, name) type name;
typedef struct flop_counter_t { HCORE_B200_FLOP_FIELDS(HCORE_B200_DECL_FIELD) } flop_counter;
#undef HCORE_B200_DECL_FIELD

#ifndef FLOP_NUMTHREADS
#define FLOP_NUMTHREADS 256
#endif
#endif
