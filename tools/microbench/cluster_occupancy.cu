// How many thread-block clusters of size 2 / 4 / 8 (one CTA per SM: ~200 KB of dynamic smem each) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / cs * cs);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("SMs %d cluster size %2d: max active clusters %d -> %d SMs busy (%s)\n", sms, cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
