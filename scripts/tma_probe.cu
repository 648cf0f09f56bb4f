// Stand-alone probe of cp.async.bulk.tensor tile loads on uint8 images (debug aid, not part of the library).
// usage: tma_probe rank x y boxw boxh [l2promo] ; prints OK/MISMATCH or the CUDA error.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int rank, int x, int y, int z, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(sm)), "l"((uint64_t)&tmap), "r"(s32(&bar)), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(sm)), "l"((uint64_t)&tmap), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), x = atoi(argv[2]), y = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]);
    int promo = argc > 6 ? atoi(argv[6]) : 2;
    const int W = 640, H = 480, F = 4, z = 2;
    std::vector<uint8_t> h((size_t)W * H * F);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (!p) { printf("no encoder\n"); return 1; }
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[3] = {W, H, F}; cuuint64_t str[2] = {W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((enc_fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, rank == 2 ? d + (size_t)z * W * H : d, dims, str, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    int bytes = bw * bh;
    probe<<<1, 128, bytes + 128>>>(tm, rank, x, y, z, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rank %d x %d y %d box %dx%d promo %d: CUDA error: %s\n", rank, x, y, bw, bh, promo, cudaGetErrorString(e)); return 2; }
    std::vector<uint8_t> got(bytes); cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
        int gx = x + xx, gy = y + yy;
        uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)z * H + gy) * W + gx] : 0;
        bad += got[yy * bw + xx] != want;
    }
    printf("rank %d x %d y %d box %dx%d promo %d: %s (%d bad)\n", rank, x, y, bw, bh, promo, bad ? "MISMATCH" : "OK", bad);
    return 0;
}
