// Explicit instantiations of the extension kernel (see extend_launch.cuh); one file per shape group for parallel builds.
#define AGATHA_DEFINE_LAUNCH
#include "extend_launch.cuh"

namespace agatha {
AGATHA_INSTANTIATE(24, 1, false, -1)
AGATHA_INSTANTIATE(24, 1, true, -1)
AGATHA_INSTANTIATE(24, 1, true, 7)
AGATHA_INSTANTIATE(24, 1, true, 15)
AGATHA_INSTANTIATE(24, 1, true, 23)
}  // namespace agatha
