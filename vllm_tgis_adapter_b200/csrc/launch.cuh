// Kernel launch helper: every hot-path kernel is launched with Programmatic Dependent Launch (PDL) so that its
// prologue (and, for the GEMM, its weight prefetch) overlaps the tail of the previous kernel in the stream.
// Contract for kernels launched through here:
//   * call griddep_launch() early (lets the NEXT kernel start its prologue), and
//   * call griddep_wait() before the first access to any buffer another kernel of the step writes or reads
//     (weights, cos/sin table and tensor maps are static and may be touched before the wait).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

namespace tgis {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TGIS_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same, with a thread-block cluster of `cluster_x` consecutive CTAs (runtime cluster size).
template <class... KArgs, class... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace tgis
