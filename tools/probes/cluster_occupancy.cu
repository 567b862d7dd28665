// How many clusters of 2 / 4 / 8 CTAs (one CTA per SM: 226 KB of dynamic shared memory, 448 threads) can be co-resident on this GPU?
// Decides whether a 4-CTA weight-multicast variant of mlp_tc_kernel could still cover all 148 SMs (DESIGN.md, K3+K4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_occupancy tools/probes/cluster_occupancy.cu && /tmp/cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe_kernel(int* out) { extern __shared__ char s[]; if (out) out[0] = s[0]; }
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"device\": \"%s\", \"sms\": %d", p.name, p.multiProcessorCount);
  const int smem = 232000;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.multiProcessorCount / cs * cs); cfg.blockDim = dim3(448); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf(", \"clusters_of_%d\": %d, \"ctas_of_%d\": %d", cs, e == cudaSuccess ? n : -1, cs, e == cudaSuccess ? n * cs : -1);
  }
  printf("}\n");
  return 0;
}
