/* TEST INFRASTRUCTURE ONLY: some reference files include <lapacke_utils.h>. */
#include "lapacke.h"
