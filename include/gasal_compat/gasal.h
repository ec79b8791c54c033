/* Forwarding header: the reference's gasal.h, provided by the agatha_b200 shim. */
#include "gasal_compat.h"
