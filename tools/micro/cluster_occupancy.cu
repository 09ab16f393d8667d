// How many thread-block clusters of 2 / 4 / 8 / 16 CTAs (1 CTA per SM: 200 KB of shared memory, 512 threads) can
// be resident on this GPU at once?  (Behind the decoder / BiLSTM design choices: DESIGN.md section 10.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/cluster_occupancy.cu -o /tmp/co && /tmp/co
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; if (p) p[0] = s[0]; }
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: %3d clusters resident = %3d of %d SMs (%s)\n", cs, n, n * cs, sms, cudaGetErrorString(e));
  }
  return 0;
}
