// tma_box_probe.cu -- stand-alone probe of the TMA box load used by csrc/fold2.cuh (same descriptor parameters, same PTX wrappers):
// loads a 34 x 34 box of doubles at (-1, -1), (31, 31) and at the far corner of a (P0 x ny) array and checks values + zero fill.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tests/experiments/tma_box_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, double *out, int nbox, int *status, int variant, const double *src, const CUtensorMap *gmap)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = smem_u32(&bar), raw = smem_u32(sm), dst = variant == 4 ? raw : (raw + 127u) & ~127u;   // variant 4: trust the declared alignment
    if (threadIdx.x == 0) status[1] = (int)(raw & 127u);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(variant == 1 ? 0 : (variant == 2 ? 4096 : nbox * 8)) : "memory");
        if (variant == 1) { /* barrier instructions only */ }
        else if (variant == 2)   // plain (non-tensor) bulk copy of 4096 bytes
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(4096), "r"(b) : "memory");
        else if (N == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(variant == 6 ? (unsigned long long)gmap : (unsigned long long)&map),
                         "r"(b), "r"(c0), "r"(c1)
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                         "l"((unsigned long long)&map), "r"(b), "r"(c0), "r"(c1), "r"(c2)
                         : "memory");
    }
    uint32_t ok = 0;
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        if (!ok && clock64() - t0 > 400000000ll) break;
    } while (!ok);
    if (threadIdx.x == 0) *status = ok ? 1 : -1;
    __syncthreads();
    const double *s = reinterpret_cast<const double *>(sm + (dst - raw));
    for (int j = threadIdx.x; j < nbox; j += blockDim.x) out[j] = ok ? s[j] : -777.0;
}

typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;   // 0 tensor f64, 1 barrier only, 2 plain bulk copy, 3 tensor described as f32 (box 68 wide)
    const int abx = argc > 2 ? atoi(argv[2]) : 34, ac0 = argc > 3 ? atoi(argv[3]) : -1;
    printf("variant %d box x %d coord %d\n", variant, abx, ac0);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    enc_t enc = (enc_t)p;
    for (int N = 2; N <= 3; ++N) {
        const long long P0 = 288, ny = N == 2 ? 259 : 25, nz = N == 2 ? 1 : 15;
        const size_t n = (size_t)P0 * ny * nz;
        std::vector<double> h(n);
        for (size_t i = 0; i < n; ++i) h[i] = (double)i + 0.5;
        double *d = nullptr, *dout = nullptr;
        int *dst = nullptr;
        cudaMalloc(&d, n * 8);
        cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
        const int nbox = N == 2 ? (variant == 5 ? 32 * 32 : (variant == 7 ? abx * 32 : 34 * 34)) : 34 * 10 * 6;
        cudaMalloc(&dout, nbox * 8);
        cudaMalloc(&dst, 8);
        const cuuint64_t dims[3] = {(cuuint64_t)P0, (cuuint64_t)ny, (cuuint64_t)nz};
        const cuuint64_t strides[2] = {(cuuint64_t)P0 * 8, (cuuint64_t)P0 * ny * 8};
        const cuuint32_t box3[3] = {34, 10, 6}, es[3] = {1, 1, 1};
        CUtensorMap m;
        const cuuint64_t dims32[3] = {(cuuint64_t)P0 * 2, (cuuint64_t)ny, (cuuint64_t)nz};
        cuuint32_t box2f[2] = {68, 34}, box3f[3] = {68, 10, 6};
        cuuint32_t box2[2] = {34, 34};
        if (variant == 5) { box2f[0] = 64; box2f[1] = 32; }
        if (variant == 7) { box2[0] = abx; box2[1] = 32; }
        CUresult r = (variant == 3 || variant == 5) ? enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)N, d, dims32, strides, N == 2 ? box2f : box3f, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                                  : enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)N, d, dims, strides, N == 2 ? box2 : box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("N=%d encode -> %d\n", N, (int)r);
        int coords[3][3] = {{-1, -1, -1}, {31, 7, 3}, {(int)P0 - 20, (int)ny - 5, (int)nz - 3}};
        if (variant == 5) { coords[0][0] = coords[0][1] = coords[0][2] = 0; }
        if (variant == 7) { coords[0][0] = coords[0][1] = ac0; }
        CUtensorMap *gm = nullptr;
        cudaMalloc(&gm, sizeof(CUtensorMap));
        cudaMemcpy(gm, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        if (N == 3 && variant >= 5) break;
        for (int t = 0; t < 3; ++t) {
            const int c0 = coords[t][0], c1 = coords[t][1], c2 = N == 3 ? coords[t][2] : 0;
            const int smem = (nbox * 8 + 127) / 128 * 128 + 128;
            const int cc0 = variant == 3 ? 2 * c0 : c0;
            if (N == 2) probe<2><<<1, 256, smem>>>(m, cc0, c1, c2, dout, nbox, dst, variant, d, gm);
            else probe<3><<<1, 256, smem>>>(m, cc0, c1, c2, dout, nbox, dst, variant, d, gm);
            cudaError_t ee = cudaDeviceSynchronize();
            int st2[2] = {0, 0};
            std::vector<double> o(nbox);
            cudaMemcpy(st2, dst, 8, cudaMemcpyDeviceToHost);
            const int st = st2[0];
            printf("  dynamic smem base & 127 = %d\n", st2[1]);
            cudaMemcpy(o.data(), dout, nbox * 8, cudaMemcpyDeviceToHost);
            long bad = 0;
            const int BX = 34, BY = N == 2 ? 34 : 10, BZ = N == 2 ? 1 : 6;
            for (int jz = 0; jz < BZ; ++jz)
                for (int jy = 0; jy < BY; ++jy)
                    for (int jx = 0; jx < BX; ++jx) {
                        const long long gx = c0 + jx, gy = c1 + jy, gz = c2 + jz;
                        const bool in = gx >= 0 && gx < P0 && gy >= 0 && gy < ny && gz >= 0 && gz < nz;
                        const double want = in ? (double)(gx + P0 * (gy + ny * gz)) + 0.5 : 0.0;
                        if (o[(jz * BY + jy) * BX + jx] != want) ++bad;
                    }
            printf("N=%d box at (%d,%d,%d): sync=%s status=%d mismatches=%ld of %d (o[0]=%g o[35]=%g)\n", N, c0, c1, c2, cudaGetErrorString(ee), st, bad, nbox, o[0], o[35]);
        }
        cudaFree(d); cudaFree(dout); cudaFree(dst);
    }
    return 0;
}
