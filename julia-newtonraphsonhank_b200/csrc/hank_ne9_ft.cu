// Instantiates the forward tangent launcher for n_e = 9 (see hank_launch.cuh).
#include "hank_launch.cuh"
namespace hank {
template int Sweeps<9>::forward_tangent(hank_ctx*, int, int, const double*, double*, int*);
}
