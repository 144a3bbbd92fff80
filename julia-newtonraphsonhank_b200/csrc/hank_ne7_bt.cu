// Instantiates the backward tangent launcher for n_e = 7 (see hank_launch.cuh).
#include "hank_launch.cuh"
namespace hank {
template int Sweeps<7>::backward_tangent(hank_ctx*, int, int, const double*, const double*, const double*, double*, double*);
}
