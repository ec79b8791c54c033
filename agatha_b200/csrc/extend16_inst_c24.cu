// Explicit instantiations of the packed extension kernel (see extend_launch.cuh); one file per shape group for parallel builds.
#define AGATHA_DEFINE_LAUNCH
#include "extend_launch.cuh"

namespace agatha {
AGATHA_INSTANTIATE16(24, 1, 7)
AGATHA_INSTANTIATE16(24, 1, 15)
AGATHA_INSTANTIATE16(24, 1, 23)
}  // namespace agatha
