/* Forwarding header: the reference's args_parser.h, provided by the agatha_b200 shim. */
#include "gasal_compat.h"
