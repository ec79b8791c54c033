// Explicit instantiations of the packed extension kernel (see extend_launch.cuh); one file per shape group for parallel builds.
#define AGATHA_DEFINE_LAUNCH
#include "extend_launch.cuh"

namespace agatha {
AGATHA_INSTANTIATE16(32, 2, 7)
AGATHA_INSTANTIATE16(32, 2, 15)
AGATHA_INSTANTIATE16(32, 2, 23)
AGATHA_INSTANTIATE16(32, 2, 31)
}  // namespace agatha
