// Instantiates the primal sweep launchers for n_e = 9 (see hank_launch.cuh); one translation unit per
// launcher group so that the build parallelises.
#include "hank_launch.cuh"
namespace hank {
template int Sweeps<9>::backward_primal(hank_ctx*, int, const double*, const double*, const double*);
template int Sweeps<9>::forward_primal(hank_ctx*, int, const double*, const double*, double*);
template int Sweeps<9>::primal_both(hank_ctx*, int, const double*, const double*, const double*, const double*);
template int Sweeps<9>::lanes_per_cta(hank_ctx*, int);
}
