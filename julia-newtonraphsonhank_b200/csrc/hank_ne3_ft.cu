// Instantiates the forward tangent launcher for n_e = 3 (see hank_launch.cuh).
#include "hank_launch.cuh"
namespace hank {
template int Sweeps<3>::forward_tangent(hank_ctx*, int, int, const double*, double*, int*);
}
