// Host-visible interface of the MPC kernels (device buffers + launchers).
#pragma once
#include <string>
#include <vector>

#include "dev_common.cuh"

namespace qmb {

struct MpcBuffers {
  int B = 0, nmax = 0;
};
bool mpc_alloc(MpcBuffers& m, int B, int nmax, std::string& err, std::vector<void*>& allocs);

}  // namespace qmb
