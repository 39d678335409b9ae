// dwconv.cu -- SS2D producer stage: x = SiLU(depthwise conv3x3(x_in) + bias), channels-last in, channel planes out.
//
// Replaces `x = x.permute(0, 3, 1, 2).contiguous(); x = self.act(self.conv2d(x))` (reference MedMamba.py:470-473:
// nn.Conv2d(d_inner, d_inner, 3, padding=1, groups=d_inner) + SiLU) and the `.float()` that follows it
// (MedMamba.py:403): one kernel reads the x half of in_proj's (B, H, W, 2 D) output IN PLACE (bf16 under autocast,
// channels contiguous), accumulates the 9 taps in fp32 and writes the (B, D, H, W) fp32 planes the cross-scan reads.
// SURVEY.md section 8(f) rank 1 (producer side).  Backward = one kernel: recompute the pre-activation, SiLU', input
// gradient (the transposed 3x3 stencil), and per-CTA partial sums of d(weight) / d(bias) added with fp32 atomics.
//
// Tiling: CTA = (batch item, 16 channels, band of TH rows); thread = (channel = tid % 16, pixel slot = tid / 16).
// Shared tiles are [pixel][16 channels] (channel fastest: the two half-warps of a warp touch different pixels, no bank
// conflicts; global reads of 16 channels = 32 / 64 contiguous bytes per pixel), outputs go through a [channel][pixel]
// tile so that every global store is a run of consecutive pixels of one plane.
#include "common.cuh"

namespace b200 {

constexpr int DW_C = 16;     // channels per CTA
constexpr int DW_TH = 4;     // rows per CTA
constexpr int DW_THREADS = 256;
constexpr int DW_MAXW = 64;  // widest image row supported by the shared tiles

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(DW_THREADS) dwconv_silu_fwd_kernel(const TI* __restrict__ xin, int64_t pix_stride, const float* __restrict__ wgt,
                                                                     const float* __restrict__ bias, float* __restrict__ out, int B, int D, int H,
                                                                     int W) {
    extern __shared__ __align__(16) float dsm[];
    const int WP = W + 2;                              // padded row length
    float* in_s = dsm;                                 // [(TH + 2) * WP][DW_C]
    float* out_s = dsm + (DW_TH + 2) * WP * DW_C;      // [DW_C][TH * W + 1]
    const int OP = DW_TH * W + 1;
    const int nbands = (H + DW_TH - 1) / DW_TH, ncg = (D + DW_C - 1) / DW_C;
    int bid = blockIdx.x;
    const int band = bid % nbands; bid /= nbands;
    const int cg = bid % ncg;
    const int b = bid / ncg;
    const int h0 = band * DW_TH, c0 = cg * DW_C;
    const int tid = threadIdx.x, c = tid % DW_C, ps = tid / DW_C;
    const bool cok = c0 + c < D;
    // stage the input band with its zero halo
    const int npix = (DW_TH + 2) * WP;
    for (int p = ps; p < npix; p += DW_THREADS / DW_C) {
        const int r = p / WP, wc = p % WP;
        const int h = h0 - 1 + r, w = wc - 1;
        float v = 0.f;
        if (cok && h >= 0 && h < H && w >= 0 && w < W) v = ldg_stream(xin + ((int64_t)(b * H + h) * W + w) * pix_stride + c0 + c);
        in_s[p * DW_C + c] = v;
    }
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = cok ? __ldg(wgt + (size_t)(c0 + c) * 9 + t) : 0.f;
    const float bv = (cok && bias) ? __ldg(bias + c0 + c) : 0.f;
    __syncthreads();
    const int rows = min(DW_TH, H - h0);
    for (int p = ps; p < rows * W; p += DW_THREADS / DW_C) {
        const int r = p / W, w = p % W;
        float acc = bv;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) acc = fmaf(k[dr * 3 + dc], in_s[((r + dr) * WP + w + dc) * DW_C + c], acc);
        out_s[c * OP + p] = acc * sigmoidf_(acc);
    }
    __syncthreads();
    // planes: the band's rows are contiguous in (B, D, H, W)
    const int n = rows * W;
    for (int cc = tid / 32; cc < DW_C; cc += DW_THREADS / 32) {
        if (c0 + cc >= D) continue;
        float* dst = out + ((size_t)(b * D + c0 + cc) * H + h0) * W;
        for (int p = tid % 32; p < n; p += 32) __stcs(dst + p, out_s[cc * OP + p]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: g = d loss / d out (B, D, H, W) fp32  ->  dxin (B, H, W, D) TI (contiguous), dweight (D, 9), dbias (D)
// ---------------------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(DW_THREADS) dwconv_silu_bwd_kernel(const float* __restrict__ g, const TI* __restrict__ xin, int64_t pix_stride,
                                                                     const float* __restrict__ wgt, const float* __restrict__ bias,
                                                                     TI* __restrict__ dxin, float* __restrict__ dwgt, float* __restrict__ dbias, int B,
                                                                     int D, int H, int W) {
    extern __shared__ __align__(16) float dsm[];
    const int WI = W + 4, WG = W + 2;
    float* in_s = dsm;                                  // input, halo 2: [(TH + 4) * WI][DW_C]
    float* dp_s = in_s + (DW_TH + 4) * WI * DW_C;       // d pre-activation, halo 1: [(TH + 2) * WG][DW_C]
    float* red_s = dp_s + (DW_TH + 2) * WG * DW_C;      // [10][DW_THREADS]
    const int nbands = (H + DW_TH - 1) / DW_TH, ncg = (D + DW_C - 1) / DW_C;
    int bid = blockIdx.x;
    const int band = bid % nbands; bid /= nbands;
    const int cg = bid % ncg;
    const int b = bid / ncg;
    const int h0 = band * DW_TH, c0 = cg * DW_C;
    const int tid = threadIdx.x, c = tid % DW_C, ps = tid / DW_C;
    const bool cok = c0 + c < D;
    constexpr int NPS = DW_THREADS / DW_C;
    for (int p = ps; p < (DW_TH + 4) * WI; p += NPS) {
        const int r = p / WI, wc = p % WI;
        const int h = h0 - 2 + r, w = wc - 2;
        float v = 0.f;
        if (cok && h >= 0 && h < H && w >= 0 && w < W) v = ldg_stream(xin + ((int64_t)(b * H + h) * W + w) * pix_stride + c0 + c);
        in_s[p * DW_C + c] = v;
    }
    // upstream gradient, halo 1, from the planes (consecutive lanes = consecutive pixels of one plane)
    for (int cc = tid / 32; cc < DW_C; cc += DW_THREADS / 32) {
        const bool ok = c0 + cc < D;
        const float* src = g + (size_t)(b * D + c0 + cc) * H * W;
        for (int p = tid % 32; p < (DW_TH + 2) * WG; p += 32) {
            const int r = p / WG, wc = p % WG;
            const int h = h0 - 1 + r, w = wc - 1;
            dp_s[p * DW_C + cc] = (ok && h >= 0 && h < H && w >= 0 && w < W) ? __ldcs(src + (size_t)h * W + w) : 0.f;
        }
    }
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = cok ? __ldg(wgt + (size_t)(c0 + c) * 9 + t) : 0.f;
    const float bv = (cok && bias) ? __ldg(bias + c0 + c) : 0.f;
    __syncthreads();
    // g -> d pre = g * silu'(pre) on the halo-1 region (pre recomputed from the halo-2 input)
    for (int p = ps; p < (DW_TH + 2) * WG; p += NPS) {
        const int r = p / WG, wc = p % WG;     // image position (h0 - 1 + r, wc - 1)
        float pre = bv;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) pre = fmaf(k[dr * 3 + dc], in_s[((r + dr) * WI + wc + dc) * DW_C + c], pre);
        const float s = sigmoidf_(pre);
        dp_s[p * DW_C + c] *= s * (1.f + pre * (1.f - s));   // zero outside the image: g was zero-filled there
    }
    __syncthreads();
    // input gradient of the band's own pixels + this CTA's share of d(weight), d(bias)
    float dk[9], db = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) dk[t] = 0.f;
    const int rows = min(DW_TH, H - h0);
    for (int p = ps; p < rows * W; p += NPS) {
        const int r = p / W, w = p % W;        // image position (h0 + r, w); dp_s index (r + 1, w + 1); in_s index (r + 2, w + 2)
        float acc = 0.f;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                // out(h + 1 - dr, w + 1 - dc) used in(h, w) with tap (dr, dc)
                acc = fmaf(k[dr * 3 + dc], dp_s[((r + 2 - dr) * WG + w + 2 - dc) * DW_C + c], acc);
            }
        if (cok) stg_stream(dxin + ((int64_t)(b * H + h0 + r) * W + w) * D + c0 + c, acc);
        const float d0 = dp_s[((r + 1) * WG + w + 1) * DW_C + c];
        db += d0;
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) dk[dr * 3 + dc] = fmaf(d0, in_s[((r + 1 + dr) * WI + w + 1 + dc) * DW_C + c], dk[dr * 3 + dc]);
    }
    // combine the 16 pixel slots of each channel, one atomic per (channel, tap) and CTA
#pragma unroll
    for (int t = 0; t < 9; ++t) red_s[t * DW_THREADS + tid] = dk[t];
    red_s[9 * DW_THREADS + tid] = db;
    __syncthreads();
    if (tid < DW_C * 10) {
        const int cc = tid % DW_C, t = tid / DW_C;
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < NPS; ++q) s += red_s[t * DW_THREADS + q * DW_C + cc];
        if (c0 + cc < D) {
            if (t < 9) atomicAdd(dwgt + (size_t)(c0 + cc) * 9 + t, s);
            else if (dbias) atomicAdd(dbias + c0 + cc, s);
        }
    }
}

static size_t dw_fwd_smem(int W) { return ((size_t)(DW_TH + 2) * (W + 2) * DW_C + (size_t)DW_C * (DW_TH * W + 1)) * sizeof(float); }
static size_t dw_bwd_smem(int W) {
    return ((size_t)(DW_TH + 4) * (W + 4) * DW_C + (size_t)(DW_TH + 2) * (W + 2) * DW_C + 10 * DW_THREADS) * sizeof(float);
}
template <class K> static int dw_set_smem(K kernel, size_t bytes) {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

template <typename TI>
static int dw_fwd(const void* xin, int64_t ps, const float* w, const float* b, float* out, int B, int D, int H, int W, cudaStream_t st) {
    const size_t smem = dw_fwd_smem(DW_MAXW);
    static const int rc0 = dw_set_smem(dwconv_silu_fwd_kernel<TI>, smem);
    if (rc0) return rc0;
    const long long grid = (long long)B * ((D + DW_C - 1) / DW_C) * ((H + DW_TH - 1) / DW_TH);
    dwconv_silu_fwd_kernel<TI><<<(unsigned)grid, DW_THREADS, dw_fwd_smem(W), st>>>((const TI*)xin, ps, w, b, out, B, D, H, W);
    return check_launch("dwconv_silu_fwd_kernel");
}
template <typename TI>
static int dw_bwd(const float* g, const void* xin, int64_t ps, const float* w, const float* b, void* dxin, float* dw, float* db, int B, int D,
                  int H, int W, cudaStream_t st) {
    const size_t smem = dw_bwd_smem(DW_MAXW);
    static const int rc0 = dw_set_smem(dwconv_silu_bwd_kernel<TI>, smem);
    if (rc0) return rc0;
    const long long grid = (long long)B * ((D + DW_C - 1) / DW_C) * ((H + DW_TH - 1) / DW_TH);
    dwconv_silu_bwd_kernel<TI><<<(unsigned)grid, DW_THREADS, dw_bwd_smem(W), st>>>(g, (const TI*)xin, ps, w, b, (TI*)dxin, dw, db, B, D, H, W);
    return check_launch("dwconv_silu_bwd_kernel");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_dwconv_silu_fwd(const void* xin, int64_t pix_stride, int32_t in_dtype, const float* weight, const float* bias, float* out,
                                    int32_t B, int32_t D, int32_t H, int32_t W, b200_stream_t stream) {
    B200_REQUIRE(xin && weight && out, "b200_dwconv_silu_fwd: NULL argument");
    B200_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && W <= DW_MAXW, "b200_dwconv_silu_fwd: bad shape (W must be <= %d)", DW_MAXW);
    B200_REQUIRE(in_dtype == B200_F32 || in_dtype == B200_BF16, "b200_dwconv_silu_fwd: input dtype must be f32 or bf16");
    cudaStream_t st = (cudaStream_t)stream;
    return in_dtype == B200_F32 ? dw_fwd<float>(xin, pix_stride, weight, bias, out, B, D, H, W, st)
                                : dw_fwd<__nv_bfloat16>(xin, pix_stride, weight, bias, out, B, D, H, W, st);
}

extern "C" int b200_dwconv_silu_bwd(const float* gout, const void* xin, int64_t pix_stride, int32_t in_dtype, const float* weight,
                                    const float* bias, void* dxin, float* dweight, float* dbias, int32_t B, int32_t D, int32_t H, int32_t W,
                                    b200_stream_t stream) {
    B200_REQUIRE(gout && xin && weight && dxin && dweight, "b200_dwconv_silu_bwd: NULL argument");
    B200_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && W <= DW_MAXW, "b200_dwconv_silu_bwd: bad shape (W must be <= %d)", DW_MAXW);
    B200_REQUIRE(in_dtype == B200_F32 || in_dtype == B200_BF16, "b200_dwconv_silu_bwd: input dtype must be f32 or bf16");
    cudaStream_t st = (cudaStream_t)stream;
    return in_dtype == B200_F32 ? dw_bwd<float>(gout, xin, pix_stride, weight, bias, dxin, dweight, dbias, B, D, H, W, st)
                                : dw_bwd<__nv_bfloat16>(gout, xin, pix_stride, weight, bias, dxin, dweight, dbias, B, D, H, W, st);
}
