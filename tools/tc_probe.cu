// tc_probe.cu -- stand-alone bring-up test of the tcgen05 (UMMA) TF32 path used for the SSD contractions:
//   D[128 x N] = A[128 x K] * B[N x K]^T,  A/B fp32 read as TF32, fp32 accumulation in TMEM.
// Operands are written to shared memory by ordinary threads (the SSD kernels apply decay / dt factors on the way in,
// so TMA cannot stage them) in the canonical NO-SWIZZLE K-major UMMA layout
//   byte offset(row r, k) = (k / 4) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 4) * 4,   SBO = 128, LBO = rows * 16
// (8-row x 16-byte core matrices; cute/atom/mma_traits_sm100.hpp "LayoutType::INTERLEAVE ((8,n),2):((1,SBO),LBO)").
// Inputs are small integers, exactly representable in TF32, so the result must match the CPU bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/tc_probe tools/tc_probe.cu && tools/_build/tc_probe [N] [K]
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    return d;                                          // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, unsigned parity, int max_iter) {
    const unsigned a = smem_u32(bar);
    for (int it = 0; it < max_iter; ++it) {
        unsigned ok;
        asm volatile(
            "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int K, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    float* sA = reinterpret_cast<float*>(smem);                 // 128 x K
    float* sB = reinterpret_cast<float*>(smem) + 128 * K;        // N x K
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr unsigned NCOLS = N < 32 ? 32 : N;                  // power of two >= 32

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(NCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operands -> canonical layout; lanes along rows so that a warp store is 512 contiguous bytes
    const unsigned lboA = 128 * 16, lboB = N * 16;
    for (int idx = tid; idx < 128 * (K / 4); idx += 128) {
        const int r = idx % 128, kc = idx / 128;
        const float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + kc * 4);
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(sA) + kc * lboA + (r / 8) * 128 + (r % 8) * 16) = v;
    }
    for (int idx = tid; idx < N * (K / 4); idx += 128) {
        const int r = idx % N, kc = idx / N;
        const float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + kc * 4);
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(sB) + kc * lboB + (r / 8) * 128 + (r % 8) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;

    if (tid == 0) {
        // instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < K / 8; ++ks) {   // one MMA = 8 TF32 along K = two 16-byte core columns
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * lboA, lboA, 128);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * lboB, lboB, 128);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                :
                : "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const bool done = mbar_wait_bounded(&bar, 0, 1 << 20);
    if (!done && tid == 0) *status = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        // warp w reads TMEM lanes 32w .. 32w+31 (= rows), 8 columns at a time
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
#pragma unroll
            for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NCOLS));
}

template <int N>
static int run(int K) {
    std::vector<float> hA(128 * K), hB((size_t)N * K), hD(128 * N), ref(128 * N);
    srand(1);
    for (auto& v : hA) v = (float)(rand() % 9 - 4);
    for (auto& v : hB) v = (float)(rand() % 9 - 4);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)hA[m * K + k] * hB[(size_t)n * K + k];
            ref[m * N + n] = (float)s;
        }
    float *dA, *dB, *dD;
    int* dS;
    cudaMalloc(&dA, hA.size() * 4);
    cudaMalloc(&dB, hB.size() * 4);
    cudaMalloc(&dD, hD.size() * 4);
    cudaMalloc(&dS, 4);
    cudaMemset(dS, 0, 4);
    cudaMemset(dD, 0xff, hD.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(128 + N) * K * 4;
    cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, K, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128 * N; ++i)
        if (hD[i] != ref[i]) {
            if (bad < 5) printf("  mismatch at (%d, %d): got %g want %g\n", i / N, i % N, hD[i], ref[i]);
            ++bad;
        }
    printf("tcgen05 tf32 probe N=%d K=%d: cuda=%s status=%d mismatches=%d/%d\n", N, K, cudaGetErrorString(e), st, bad, 128 * N);
    return (e != cudaSuccess || st || bad) ? 1 : 0;
}

int main(int argc, char** argv) {
    const int K = argc > 2 ? atoi(argv[2]) : 64;
    const int N = argc > 1 ? atoi(argv[1]) : 64;
    if (N == 64) return run<64>(K);
    if (N == 128) return run<128>(K);
    if (N == 256) return run<256>(K);
    printf("N must be 64, 128 or 256\n");
    return 2;
}
