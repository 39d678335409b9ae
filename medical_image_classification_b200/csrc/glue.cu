// glue.cu -- SS_Conv_SSM block tail: out = channel_shuffle(cat(left, x), groups = 2) + input, forward and backward.
//
// Replaces `left.permute(0, 2, 3, 1)`, `x.to(left.dtype)`, `torch.cat`, the channel shuffle's transpose-copy and the
// residual add of the reference block (MedMamba.py:486-499, 533-538): SURVEY.md section 8(f) rank 2.  With c = C / 2
//   out[b, p, 2 j]     = left[b, j, p] + input[b, p, 2 j]        (left: the conv branch's (B, c, H, W) planes)
//   out[b, p, 2 j + 1] = x[b, p, j]    + input[b, p, 2 j + 1]    (x: the SS2D branch, channels-last)
// i.e. a 32 x 32 plane-to-channels-last transpose fused with an interleave and an add: every global access is a
// full coalesced row (lanes along pixels for the planes, along channels for the channels-last tensors).
#include "common.cuh"

namespace b200 {

template <typename TL>
__global__ void __launch_bounds__(256) shuffle_cat_add_fwd_kernel(const TL* __restrict__ left, const TL* __restrict__ x,
                                                                  const float* __restrict__ input, float* __restrict__ out, int c, int P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, j0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {   // planes: lanes along pixels
        const int j = j0 + r, p = p0 + tx;
        tile[r][tx] = (j < c && p < P) ? ldg_stream(left + ((size_t)b * c + j) * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {   // channels-last: lanes along channels
        const int p = p0 + r, j = j0 + tx;
        if (p < P && j < c) {
            const size_t row = (size_t)b * P + p;
            const float2 in2 = __ldcs(reinterpret_cast<const float2*>(input + row * 2 * c) + j);
            const float xv = ldg_stream(x + row * c + j);
            __stcs(reinterpret_cast<float2*>(out + row * 2 * c) + j, make_float2(tile[tx][r] + in2.x, xv + in2.y));
        }
    }
}

template <typename TL>
__global__ void __launch_bounds__(256) shuffle_cat_add_bwd_kernel(const float* __restrict__ dout, TL* __restrict__ dleft, TL* __restrict__ dx, int c,
                                                                  int P) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, j0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, j = j0 + tx;
        float2 g = make_float2(0.f, 0.f);
        if (p < P && j < c) {
            const size_t row = (size_t)b * P + p;
            g = __ldcs(reinterpret_cast<const float2*>(dout + row * 2 * c) + j);
            stg_stream(dx + row * c + j, g.y);
        }
        tile[r][tx] = g.x;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, p = p0 + tx;
        if (j < c && p < P) stg_stream(dleft + ((size_t)b * c + j) * P + p, tile[tx][r]);
    }
}

// left already channels-last (B, P, c): a pure interleave, two rows of c -> one row of 2 c
template <typename TL>
__global__ void __launch_bounds__(256) shuffle_cat_add_cl_fwd_kernel(const TL* __restrict__ left, const TL* __restrict__ x,
                                                                     const float* __restrict__ input, float* __restrict__ out, int c, int64_t rows) {
    const int64_t total = rows * c;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const float2 in2 = __ldcs(reinterpret_cast<const float2*>(input) + idx);   // (row, 2 j), (row, 2 j + 1)
        __stcs(reinterpret_cast<float2*>(out) + idx, make_float2(ldg_stream(left + idx) + in2.x, ldg_stream(x + idx) + in2.y));
    }
}
template <typename TL>
__global__ void __launch_bounds__(256) shuffle_cat_add_cl_bwd_kernel(const float* __restrict__ dout, TL* __restrict__ dleft, TL* __restrict__ dx,
                                                                     int c, int64_t rows) {
    const int64_t total = rows * c;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const float2 g = __ldcs(reinterpret_cast<const float2*>(dout) + idx);
        stg_stream(dleft + idx, g.x);
        stg_stream(dx + idx, g.y);
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_shuffle_cat_add_fwd(const void* left, int32_t left_channels_last, const void* x, int32_t lx_dtype, const float* input,
                                        float* out, int32_t B, int32_t c, int32_t P, b200_stream_t stream) {
    B200_REQUIRE(left && x && input && out, "b200_shuffle_cat_add_fwd: NULL argument");
    B200_REQUIRE(B > 0 && c > 0 && P > 0 && B <= 65535, "b200_shuffle_cat_add_fwd: bad shape");
    B200_REQUIRE(lx_dtype == B200_F32 || lx_dtype == B200_BF16, "b200_shuffle_cat_add_fwd: left / x dtype must be f32 or bf16");
    const dim3 grid((P + 31) / 32, (c + 31) / 32, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (left_channels_last) {
        const int64_t rows = (int64_t)B * P;
        const int64_t want = (rows * c + 255) / 256;
        const unsigned g1 = (unsigned)(want < 148 * 16 ? want : 148 * 16);
        if (lx_dtype == B200_F32) shuffle_cat_add_cl_fwd_kernel<float><<<g1, 256, 0, st>>>((const float*)left, (const float*)x, input, out, c, rows);
        else shuffle_cat_add_cl_fwd_kernel<__nv_bfloat16><<<g1, 256, 0, st>>>((const __nv_bfloat16*)left, (const __nv_bfloat16*)x, input, out, c, rows);
        return check_launch("shuffle_cat_add_cl_fwd_kernel");
    }
    if (lx_dtype == B200_F32) shuffle_cat_add_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)left, (const float*)x, input, out, c, P);
    else shuffle_cat_add_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)left, (const __nv_bfloat16*)x, input, out, c, P);
    return check_launch("shuffle_cat_add_fwd_kernel");
}

extern "C" int b200_shuffle_cat_add_bwd(const float* dout, void* dleft, int32_t left_channels_last, void* dx, int32_t lx_dtype, int32_t B,
                                        int32_t c, int32_t P, b200_stream_t stream) {
    B200_REQUIRE(dout && dleft && dx, "b200_shuffle_cat_add_bwd: NULL argument");
    B200_REQUIRE(B > 0 && c > 0 && P > 0 && B <= 65535, "b200_shuffle_cat_add_bwd: bad shape");
    B200_REQUIRE(lx_dtype == B200_F32 || lx_dtype == B200_BF16, "b200_shuffle_cat_add_bwd: left / x dtype must be f32 or bf16");
    const dim3 grid((P + 31) / 32, (c + 31) / 32, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (left_channels_last) {
        const int64_t rows = (int64_t)B * P;
        const int64_t want = (rows * c + 255) / 256;
        const unsigned g1 = (unsigned)(want < 148 * 16 ? want : 148 * 16);
        if (lx_dtype == B200_F32) shuffle_cat_add_cl_bwd_kernel<float><<<g1, 256, 0, st>>>(dout, (float*)dleft, (float*)dx, c, rows);
        else shuffle_cat_add_cl_bwd_kernel<__nv_bfloat16><<<g1, 256, 0, st>>>(dout, (__nv_bfloat16*)dleft, (__nv_bfloat16*)dx, c, rows);
        return check_launch("shuffle_cat_add_cl_bwd_kernel");
    }
    if (lx_dtype == B200_F32) shuffle_cat_add_bwd_kernel<float><<<grid, 256, 0, st>>>(dout, (float*)dleft, (float*)dx, c, P);
    else shuffle_cat_add_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dout, (__nv_bfloat16*)dleft, (__nv_bfloat16*)dx, c, P);
    return check_launch("shuffle_cat_add_bwd_kernel");
}
