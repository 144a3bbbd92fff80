// Instantiates the primal sweep launchers for n_e = 7 (see hank_launch.cuh); one translation unit per
// launcher group so that the build parallelises.
#include "hank_launch.cuh"
namespace hank {
template int Sweeps<7>::backward_primal(hank_ctx*, int, const double*, const double*, const double*);
template int Sweeps<7>::forward_primal(hank_ctx*, int, const double*, const double*, double*);
template int Sweeps<7>::primal_both(hank_ctx*, int, const double*, const double*, const double*, const double*);
template int Sweeps<7>::lanes_per_cta(hank_ctx*, int);
}
