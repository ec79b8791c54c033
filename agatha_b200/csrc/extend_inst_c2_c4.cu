// Explicit instantiations of the extension kernel (see extend_launch.cuh); one file per shape group for parallel builds.
#define AGATHA_DEFINE_LAUNCH
#include "extend_launch.cuh"

namespace agatha {
AGATHA_INSTANTIATE(2, 1, false, -1)
AGATHA_INSTANTIATE(2, 1, true, -1)
AGATHA_INSTANTIATE(2, 1, true, 1)
AGATHA_INSTANTIATE(4, 1, false, -1)
AGATHA_INSTANTIATE(4, 1, true, -1)
AGATHA_INSTANTIATE(4, 1, true, 3)
}  // namespace agatha
