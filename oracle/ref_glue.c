/* TEST INFRASTRUCTURE ONLY.
 * Glue compiled into oracle/_ref/libhcore_ref.so next to the unmodified reference:
 *  - defines `use_scratch`, which the reference declares extern and expects the
 *    application to provide (reference src/compute/hcore_dgemm.c:38,
 *    testing/testing_dgemm.c:26);
 *  - exposes the BLAS thread control and build string so the bench can report them.
 */
#include "cblas.h"
int use_scratch = 0;
void oracle_ref_set_blas_threads(int n) { scipy_openblas_set_num_threads(n); }
const char *oracle_ref_blas_config(void) { return scipy_openblas_get_config(); }
