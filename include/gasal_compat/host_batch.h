/* Forwarding header: the reference's host_batch.h, provided by the agatha_b200 shim. */
#include "gasal_compat.h"
