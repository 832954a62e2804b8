// How many thread-block clusters of a one-CTA-per-SM kernel (320 threads, ~226 KB of dynamic shared memory: the shape of
// the persistent LSTM kernels) can be co-resident on this GPU, by cluster size?  (cudaOccupancyMaxActiveClusters)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o probe_cluster_occupancy.bin probe_cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(320, 1) big_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
}

int main() {
  const int smem = 231000;
  cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(big_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int shapes[][2] = {{1, 1}, {2, 1}, {1, 4}, {2, 4}, {8, 1}, {16, 1}};
  for (auto& sh : shapes) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16 * sh[0], 4 * sh[1], 8);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a;
    a.id = cudaLaunchAttributeClusterDimension;
    a.val.clusterDim.x = sh[0]; a.val.clusterDim.y = sh[1]; a.val.clusterDim.z = 1;
    cfg.attrs = &a; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, big_kernel, &cfg);
    printf("cluster %2d x %d (%2d CTAs): max active clusters %3d = %3d CTAs  (%s)\n", sh[0], sh[1], sh[0] * sh[1], n,
           n * sh[0] * sh[1], cudaGetErrorString(e));
  }
  return 0;
}
