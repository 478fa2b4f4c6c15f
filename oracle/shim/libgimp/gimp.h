/* shim: see dctc_shim_types.h (test infrastructure, not reference code) */
#include "dctc_shim_types.h"
