// development probe: which u8 tiled tensor-map variants does UTMALDG accept on this GPU?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include "../pixel_art_remaster_gpu_b200/csrc/common.cuh"
using namespace par;
__global__ void k3(const __grid_constant__ CUtensorMap m, int bytes, int c0, int c1, int c2, unsigned* out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_3d(sm, &m, &bar, c0, c1, c2); }
    mbar_wait(&bar, 0);
    unsigned s = 0; for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += sm[i];
    atomicAdd(out, s);
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    int only = argc > 1 ? atoi(argv[1]) : -1; int vi = -1;
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Enc enc = (Enc)fn;
    const int W = 288, H = 80, F = 2;
    uint8_t* d; cudaMalloc(&d, W * H * F); cudaMemset(d, 1, W * H * F);
    unsigned* out; cudaMalloc(&out, 4);
    struct V { const char* name; unsigned b0, b1; CUtensorMapL2promotion l2; int c0, c1; } vs[] = {
        {"64x16 c0=-16 c1=-1", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, -16, -1},
        {"64x16 c0=0 c1=-1", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0, -1},
        {"64x16 c0=13 c1=0", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 13, 0},
        {"64x16 c0=-3 c1=0", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, -3, 0},
        {"64x16 c0=16 c1=5", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 16, 5},
        {"64x16 c0=272 c1=70", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 272, 70},
        {"64x16 c0=4 c1=0", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 4, 0},
        {"64x16 c0=8 c1=0", 64, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 8, 0},
    };
    for (auto& v : vs) {
        vi++; if (only >= 0 && vi != only) continue;
        CUtensorMap m; cuuint64_t dims[3] = {W, H, F}; cuuint64_t st[2] = {W, (cuuint64_t)W * H};
        cuuint32_t box[3] = {v.b0, v.b1, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, v.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemset(out, 0, 4);
        k3<<<1, 128, v.b0 * v.b1>>>(m, v.b0 * v.b1, v.c0, v.c1, 1, out);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned h = 0; if (e == cudaSuccess) cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
        printf("%-22s encode=%d run=%s sum=%u\n", v.name, (int)r, cudaGetErrorString(e), h);
        if (e != cudaSuccess) { printf("context dead, stopping\n"); return 1; }
    }
    return 0;
}
