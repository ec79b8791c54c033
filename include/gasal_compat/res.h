/* Forwarding header: the reference's res.h, provided by the agatha_b200 shim. */
#include "gasal_compat.h"
