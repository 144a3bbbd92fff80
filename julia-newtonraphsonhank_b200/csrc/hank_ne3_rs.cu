// Instantiates the row-split cluster tangent launchers for n_e = 3 (see hank_launch_rs.cuh).
#include "hank_launch_rs.cuh"
namespace hank {
template int Sweeps<3>::backward_tangent_rs(hank_ctx*, int, int, int, int, int, int, const double*, const double*, double*);
template int Sweeps<3>::rs_max_clusters(hank_ctx*, int, int, int, int);
template int Sweeps<3>::forward_tangent_rs(hank_ctx*, int, int, int, int, int, int, const double*, double*);
}
