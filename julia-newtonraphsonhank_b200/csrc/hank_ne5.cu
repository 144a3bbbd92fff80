// Instantiates the sweep launchers for n_e = 5 (see hank_launch.cuh).
#include "hank_launch.cuh"
namespace hank { template struct Sweeps<5>; }
