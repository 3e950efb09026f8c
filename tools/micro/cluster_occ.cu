// dev microbenchmark: how many clusters of a 320-thread, ~220 KB-smem kernel are co-resident, by cluster size
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 6, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d (= %d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
