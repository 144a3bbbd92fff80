// Instantiates the primal sweep launchers for n_e = 5 (see hank_launch.cuh); one translation unit per
// launcher group so that the build parallelises.
#include "hank_launch.cuh"
namespace hank {
template int Sweeps<5>::backward_primal(hank_ctx*, int, const double*, const double*, const double*);
template int Sweeps<5>::forward_primal(hank_ctx*, int, const double*, const double*, double*);
template int Sweeps<5>::primal_both(hank_ctx*, int, const double*, const double*, const double*, const double*);
template int Sweeps<5>::lanes_per_cta(hank_ctx*, int);
}
