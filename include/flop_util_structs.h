/* Drop-in name of the reference header (src/misc/include/flop_util_structs.h). */
#include "hcore_b200_flops.h"
