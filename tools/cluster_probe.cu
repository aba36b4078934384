// cluster_probe.cu - how many clusters of S CTAs (576 threads, ~200 KB dynamic shared memory each: the footprint of
// plane_conv_umma_fold) can be resident at once on this GPU?  Feeds the planner's cluster-slot table (kernels_umma.cu).
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(576, 1) dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int S = 1; S <= 16; ++S) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(S * 64); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute a; a.id = cudaLaunchAttributeClusterDimension; a.val.clusterDim.x = S; a.val.clusterDim.y = 1; a.val.clusterDim.z = 1;
        cfg.attrs = &a; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("S=%2d max active clusters %3d (= %3d CTAs)  %s\n", S, n, n * S, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
