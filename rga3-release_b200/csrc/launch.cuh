// Kernel launch helper: every kernel of the path is launched with programmatic dependent launch
// (PDL) enabled, so the next kernel's CTAs are scheduled -- and run their set-up (barrier init, TMEM
// allocation, tensor-map prefetch) -- while the previous kernel drains; each kernel executes
// griddepcontrol.wait before it touches global memory produced (or still read) by its predecessor.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace b200 {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VIT_PDL");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                          Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace b200
