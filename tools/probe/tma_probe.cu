// Probe: TMA tile loads of FLOAT64 boxes with out-of-range coordinates (what eu_tile.cuh relies on).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k3(const __grid_constant__ CUtensorMap m, int x, int y, int z, int n, double* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned bar = (unsigned)__cvta_generic_to_shared(sm);
    unsigned dst = bar + 128;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(n*8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(dst), "l"(&m), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" :: "r"(bar) : "memory");
    const double* s = reinterpret_cast<const double*>(sm + 128);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}
__global__ void k4(const __grid_constant__ CUtensorMap m, int x, int y, int z, int w, int n, double* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned bar = (unsigned)__cvta_generic_to_shared(sm);
    unsigned dst = bar + 128;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(n*8) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     :: "r"(dst), "l"(&m), "r"(x), "r"(y), "r"(z), "r"(w), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" :: "r"(bar) : "memory");
    const double* s = reinterpret_cast<const double*>(sm + 128);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv)
{
    // usage: tma_probe <rank 3|4> <x> <y> <z> [boxx boxy] : one TMA load per process (an illegal instruction kills the context)
    const int rank = atoi(argv[1]), x = atoi(argv[2]), y = atoi(argv[3]), z = atoi(argv[4]);
    const int bx = argc > 5 ? atoi(argv[5]) : 34, by = argc > 6 ? atoi(argv[6]) : 10;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int nx = 64, ny = 16, nz = 6, n = nx*ny*nz;
    std::vector<double> h(3*n); for (int i = 0; i < 3*n; ++i) h[i] = i + 1;
    double *d, *out; cudaMalloc(&d, 3*n*8); cudaMalloc(&out, 65536); cudaMemcpy(d, h.data(), 3*n*8, cudaMemcpyHostToDevice);
    cuuint32_t ones[5] = {1,1,1,1,1};
    const int nb = bx*by;
    CUtensorMap m;
    CUresult r;
    if (rank == 3) {
        cuuint64_t dims[3] = {nx, ny, nz}; cuuint64_t str[2] = {nx*8, nx*ny*8}; cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
        r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        k3<<<1, 128, 128 + nb*8>>>(m, x, y, z, nb, out);
    } else {
        cuuint64_t dims[4] = {nx, ny, nz, 3}; cuuint64_t str[3] = {nx*8, nx*ny*8, (cuuint64_t)n*8}; cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1};
        r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        k4<<<1, 128, 128 + nb*8>>>(m, x, y, z, 2, nb, out);
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> o(nb, -7.0);
    if (e == cudaSuccess) cudaMemcpy(o.data(), out, nb*8, cudaMemcpyDeviceToHost);
    auto at = [&](int gx, int gy, int gz) { return (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nz) ? h[(rank == 4 ? 2*(size_t)n : 0) + gx + nx*(gy + ny*gz)] : 0.0; };
    int bad = 0;
    for (int j = 0; j < by; ++j) for (int i = 0; i < bx; ++i) if (o[j*bx + i] != at(x + i, y + j, z)) ++bad;
    printf("rank %d box %dx%d at (%d,%d,%d): encode %d, %s, mismatches %d of %d\n", rank, bx, by, x, y, z, (int)r, cudaGetErrorString(e), bad, nb);
    return 0;
}
