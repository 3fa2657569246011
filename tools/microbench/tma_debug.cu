// Bring-up check for tensor-map TMA loads on the reference spectrogram layout (see tma_box_stream.cu): one CTA, one
// box, result compared element by element.  mode 0: plain 2-D map of a 16-byte-aligned matrix; mode 1: the 5-D
// (frame, row-group, q, plane, clip) map with the base moved back to a 16-byte boundary.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_debug tma_debug.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int n, int c0, int c1, int c2, int c3, int c4) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
    float* dst = reinterpret_cast<float*>(smem + 1024);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 4) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                             smem_u32(dst)),
                         "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                         : "memory");
        else
            asm volatile(
                "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                    smem_u32(dst)),
                "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(bar)), "r"(0)
        : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int T = argc > 2 ? atoi(argv[2]) : 862;
    const int rcls = argc > 3 ? atoi(argv[3]) : 1;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn encode = (EncodeFn)fn;
    printf("mode %d T %d rcls %d entry %p qres %d\n", mode, T, rcls, fn, (int)qres);
    const int clips = 2, rows = 1024;
    const size_t elems = (size_t)clips * 3 * rows * T;
    std::vector<float> h(elems);
    for (size_t i = 0; i < elems; ++i) h[i] = (float)(i % 1000003);
    float* d;
    CK(cudaMalloc(&d, elems * 4 + 64));
    CK(cudaMemcpy(d, h.data(), elems * 4, cudaMemcpyHostToDevice));
    float* out;
    CK(cudaMalloc(&out, 65536));
    CUtensorMap map;
    int n = 0;
    std::vector<float> ref;
    if (mode == 0) {
        // rows of 4 rows each: [rows/4 * 3 * clips][4*T], box 16 x 8
        cuuint64_t dims[2] = {(cuuint64_t)4 * T, (cuuint64_t)clips * 3 * rows / 4};
        cuuint64_t strides[1] = {(cuuint64_t)16 * T};
        cuuint32_t box[2] = {16, 8}, es[2] = {1, 1};
        CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode rc %d\n", (int)rc);
        n = 128;
        CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        k<2><<<1, 128, 32768>>>(map, out, n, 32, 5, 0, 0, 0);
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 16; ++c) ref.push_back(h[(size_t)(5 + r) * 4 * T + 32 + c]);
    } else {
        const long long a = ((long long)rcls * T * 4) & 15;
        char* base = (char*)d + (long long)rcls * T * 4 - a;
        cuuint64_t dims[5] = {(cuuint64_t)(T + a / 4), 8, 32, 3, (cuuint64_t)clips};
        cuuint64_t strides[4] = {(cuuint64_t)4 * T * 4, (cuuint64_t)32 * T * 4, (cuuint64_t)rows * T * 4, (cuuint64_t)3 * rows * T * 4};
        cuuint32_t box[5] = {16, 1, 32, 3, 1}, es[5] = {1, 1, 1, 1, 1};
        CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode rc %d base %p a %lld\n", (int)rc, (void*)base, a);
        n = 16 * 32 * 3;
        const int t0 = 48, jr = 3, b = 1;
        CK(cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        k<5><<<1, 128, 32768>>>(map, out, n, t0 + (int)(a / 4), jr, 0, 0, b);
        for (int pl = 0; pl < 3; ++pl)
            for (int q = 0; q < 32; ++q)
                for (int c = 0; c < 16; ++c) {
                    const size_t row = rcls + 4 * jr + 32 * q;
                    ref.push_back(h[(((size_t)b * 3 + pl) * rows + row) * T + t0 + c]);
                }
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> got(n);
    CK(cudaMemcpy(got.data(), out, n * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += got[i] != ref[i];
    printf("mismatches %d of %d (got[0] %.0f ref[0] %.0f)\n", bad, n, got[0], ref[0]);
    return 0;
}
