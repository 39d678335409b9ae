// glue.cu -- SS_Conv_SSM block tail: out = channel_shuffle(cat(left, x), groups = 2) + input, forward and backward.
//
// Replaces `left.permute(0, 2, 3, 1)`, `x.to(left.dtype)`, `torch.cat`, the channel shuffle's transpose-copy and the
// residual add of the reference block (MedMamba.py:486-499, 533-538): SURVEY.md section 8(f) rank 2.  With c = C / 2
//   out[b, p, 2 j]     = left[b, j, p] + input[b, p, 2 j]        (left: the conv branch's (B, c, H, W) planes)
//   out[b, p, 2 j + 1] = x[b, p, j]    + input[b, p, 2 j + 1]    (x: the SS2D branch, channels-last)
// i.e. a 32 x 32 plane-to-channels-last transpose fused with an interleave and an add: every global access is a
// full coalesced row (lanes along pixels for the planes, along channels for the channels-last tensors).
// input / out are fp32 or -- with bf16 left / x, the residual stream of stages 1-3 under autocast -- bf16 (sum in fp32, one rounding).
#include "common.cuh"

namespace b200 {

// element pair idx of a channels-last row buffer (fp32: one float2, bf16: one 32-bit word), streaming
template <typename T> __device__ __forceinline__ float2 ld_pair(const T* base, int64_t idx);
template <> __device__ __forceinline__ float2 ld_pair<float>(const float* base, int64_t idx) { return __ldcs(reinterpret_cast<const float2*>(base) + idx); }
template <> __device__ __forceinline__ float2 ld_pair<__nv_bfloat16>(const __nv_bfloat16* base, int64_t idx) {
    const unsigned w = __ldcs(reinterpret_cast<const unsigned*>(base) + idx);
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ void st_pair(T* base, int64_t idx, float a, float b);
template <> __device__ __forceinline__ void st_pair<float>(float* base, int64_t idx, float a, float b) {
    __stcs(reinterpret_cast<float2*>(base) + idx, make_float2(a, b));
}
template <> __device__ __forceinline__ void st_pair<__nv_bfloat16>(__nv_bfloat16* base, int64_t idx, float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    __stcs(reinterpret_cast<unsigned*>(base) + idx, *reinterpret_cast<const unsigned*>(&v));
}

template <typename TL, typename TIO>
__global__ void __launch_bounds__(256) shuffle_cat_add_fwd_kernel(const TL* __restrict__ left, const TL* __restrict__ x,
                                                                  const TIO* __restrict__ input, TIO* __restrict__ out, int c, int P) {
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
            const float2 in2 = ld_pair<TIO>(input, (int64_t)row * c + j);
            const float xv = ldg_stream(x + row * c + j);
            st_pair<TIO>(out, (int64_t)row * c + j, tile[tx][r] + in2.x, xv + in2.y);
        }
    }
}

template <typename TL, typename TIO>
__global__ void __launch_bounds__(256) shuffle_cat_add_bwd_kernel(const TIO* __restrict__ dout, TL* __restrict__ dleft, TL* __restrict__ dx, int c,
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
            g = ld_pair<TIO>(dout, (int64_t)row * c + j);
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
template <typename TL, typename TIO>
__global__ void __launch_bounds__(256) shuffle_cat_add_cl_fwd_kernel(const TL* __restrict__ left, const TL* __restrict__ x,
                                                                     const TIO* __restrict__ input, TIO* __restrict__ out, int c, int64_t rows) {
    const int64_t total = rows * c;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const float2 in2 = ld_pair<TIO>(input, idx);   // (row, 2 j), (row, 2 j + 1)
        st_pair<TIO>(out, idx, ldg_stream(left + idx) + in2.x, ldg_stream(x + idx) + in2.y);
    }
}
template <typename TL, typename TIO>
__global__ void __launch_bounds__(256) shuffle_cat_add_cl_bwd_kernel(const TIO* __restrict__ dout, TL* __restrict__ dleft, TL* __restrict__ dx,
                                                                     int c, int64_t rows) {
    const int64_t total = rows * c;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const float2 g = ld_pair<TIO>(dout, idx);
        stg_stream(dleft + idx, g.x);
        stg_stream(dx + idx, g.y);
    }
}

static unsigned cl_grid(int64_t rows, int c) {
    const int64_t want = (rows * c + 255) / 256;
    return (unsigned)(want < 148 * 16 ? want : 148 * 16);
}
template <typename TL, typename TIO>
static int glue_fwd(const void* left, int cl, const void* x, const void* input, void* out, int B, int c, int P, cudaStream_t st) {
    if (cl) {
        const int64_t rows = (int64_t)B * P;
        shuffle_cat_add_cl_fwd_kernel<TL, TIO><<<cl_grid(rows, c), 256, 0, st>>>((const TL*)left, (const TL*)x, (const TIO*)input, (TIO*)out, c, rows);
        return check_launch("shuffle_cat_add_cl_fwd_kernel");
    }
    const dim3 grid((P + 31) / 32, (c + 31) / 32, B);
    shuffle_cat_add_fwd_kernel<TL, TIO><<<grid, 256, 0, st>>>((const TL*)left, (const TL*)x, (const TIO*)input, (TIO*)out, c, P);
    return check_launch("shuffle_cat_add_fwd_kernel");
}
template <typename TL, typename TIO>
static int glue_bwd(const void* dout, void* dleft, int cl, void* dx, int B, int c, int P, cudaStream_t st) {
    if (cl) {
        const int64_t rows = (int64_t)B * P;
        shuffle_cat_add_cl_bwd_kernel<TL, TIO><<<cl_grid(rows, c), 256, 0, st>>>((const TIO*)dout, (TL*)dleft, (TL*)dx, c, rows);
        return check_launch("shuffle_cat_add_cl_bwd_kernel");
    }
    const dim3 grid((P + 31) / 32, (c + 31) / 32, B);
    shuffle_cat_add_bwd_kernel<TL, TIO><<<grid, 256, 0, st>>>((const TIO*)dout, (TL*)dleft, (TL*)dx, c, P);
    return check_launch("shuffle_cat_add_bwd_kernel");
}

// ---- PatchMerging2D gather (reference MedMamba.py:186-204): x (B, H, W, C) -> out (B, H/2, W/2, 4 C) with
//      out[b, h2, w2, k C + c] = x[b, 2 h2 + (k & 1), 2 w2 + (k >> 1), c]   (x0 | x1 | x2 | x3 = (0,0) (1,0) (0,1) (1,1)),
//      rows / columns beyond 2 (H/2), 2 (W/2) dropped as the reference does.  A pure permutation of 16-byte units (C elements of
//      any dtype per pixel, C * size % 16 == 0): one pass instead of four strided slice copies + cat; the backward is the inverse
//      permutation (every input pixel feeds exactly one output slot; dropped rows / columns get zero). ----
__global__ void __launch_bounds__(256) patch_merge_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H, int W, int upp, int64_t total,
                                                          int inverse) {
    // upp: 16-byte units per input pixel.  forward: one thread per OUTPUT unit; inverse: one thread per INPUT unit (of dx)
    const int H2 = H >> 1, W2 = W >> 1;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (!inverse) {
            const int u = (int)(idx % upp);
            int64_t r = idx / upp;
            const int k = (int)(r & 3); r >>= 2;
            const int w2 = (int)(r % W2); r /= W2;
            const int h2 = (int)(r % H2);
            const int64_t b = r / H2;
            out[idx] = __ldcs(x + ((b * H + 2 * h2 + (k & 1)) * W + 2 * w2 + (k >> 1)) * upp + u);
        } else {   // x = d out (B, H2, W2, 4 C), out = d x (B, H, W, C)
            const int u = (int)(idx % upp);
            int64_t r = idx / upp;
            const int w = (int)(r % W); r /= W;
            const int h = (int)(r % H);
            const int64_t b = r / H;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if ((h >> 1) < H2 && (w >> 1) < W2) {
                const int k = (h & 1) + 2 * (w & 1);
                v = __ldcs(x + (((b * H2 + (h >> 1)) * W2 + (w >> 1)) * 4 + k) * upp + u);
            }
            out[idx] = v;
        }
    }
}

}  // namespace b200

extern "C" int b200_patch_merge(const void* x, void* out, int32_t batch, int32_t H, int32_t W, int32_t pixel_bytes, int32_t inverse,
                                b200_stream_t stream) {
    using namespace b200;
    B200_REQUIRE(x && out && batch > 0 && H >= 2 && W >= 2 && pixel_bytes > 0 && pixel_bytes % 16 == 0,
                 "b200_patch_merge: need H, W >= 2 and a pixel size that is a multiple of 16 bytes (got %d)", pixel_bytes);
    B200_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "b200_patch_merge: 16-byte aligned tensors");
    const int upp = pixel_bytes / 16;
    const int64_t total = inverse ? (int64_t)batch * H * W * upp : (int64_t)batch * (H / 2) * (W / 2) * 4 * upp;
    const int64_t want = (total + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
    patch_merge_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, H, W, upp, total, inverse);
    return check_launch("patch_merge_kernel");
}

namespace b200 {
}  // namespace b200

using namespace b200;

// (left / x dtype, input / out dtype): f32 x f32; bf16 x f32 (stage 0 of an autocast model: fp32 residual stream); bf16 x bf16
// (after PatchMerging's Linear the residual stream of an autocast model is bf16)
extern "C" int b200_shuffle_cat_add_fwd(const void* left, int32_t left_channels_last, const void* x, int32_t lx_dtype, const void* input,
                                        void* out, int32_t io_dtype, int32_t B, int32_t c, int32_t P, b200_stream_t stream) {
    B200_REQUIRE(left && x && input && out, "b200_shuffle_cat_add_fwd: NULL argument");
    B200_REQUIRE(B > 0 && c > 0 && P > 0 && B <= 65535, "b200_shuffle_cat_add_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (lx_dtype == B200_F32 && io_dtype == B200_F32) return glue_fwd<float, float>(left, left_channels_last, x, input, out, B, c, P, st);
    if (lx_dtype == B200_BF16 && io_dtype == B200_F32) return glue_fwd<__nv_bfloat16, float>(left, left_channels_last, x, input, out, B, c, P, st);
    if (lx_dtype == B200_BF16 && io_dtype == B200_BF16)
        return glue_fwd<__nv_bfloat16, __nv_bfloat16>(left, left_channels_last, x, input, out, B, c, P, st);
    B200_REQUIRE(false, "b200_shuffle_cat_add_fwd: unsupported dtypes left/x = %d, input/out = %d", lx_dtype, io_dtype);
}

extern "C" int b200_shuffle_cat_add_bwd(const void* dout, int32_t io_dtype, void* dleft, int32_t left_channels_last, void* dx, int32_t lx_dtype,
                                        int32_t B, int32_t c, int32_t P, b200_stream_t stream) {
    B200_REQUIRE(dout && dleft && dx, "b200_shuffle_cat_add_bwd: NULL argument");
    B200_REQUIRE(B > 0 && c > 0 && P > 0 && B <= 65535, "b200_shuffle_cat_add_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (lx_dtype == B200_F32 && io_dtype == B200_F32) return glue_bwd<float, float>(dout, dleft, left_channels_last, dx, B, c, P, st);
    if (lx_dtype == B200_BF16 && io_dtype == B200_F32) return glue_bwd<__nv_bfloat16, float>(dout, dleft, left_channels_last, dx, B, c, P, st);
    if (lx_dtype == B200_BF16 && io_dtype == B200_BF16)
        return glue_bwd<__nv_bfloat16, __nv_bfloat16>(dout, dleft, left_channels_last, dx, B, c, P, st);
    B200_REQUIRE(false, "b200_shuffle_cat_add_bwd: unsupported dtypes left/x = %d, dout = %d", lx_dtype, io_dtype);
}
