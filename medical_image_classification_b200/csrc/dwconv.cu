// dwconv.cu -- SS2D producer stage: x = SiLU(depthwise conv3x3(x_in) + bias), channels-last in, channel planes out.
//
// Replaces `x = x.permute(0, 3, 1, 2).contiguous(); x = self.act(self.conv2d(x))` (reference MedMamba.py:470-473:
// nn.Conv2d(d_inner, d_inner, 3, padding=1, groups=d_inner) + SiLU) and the `.float()` that follows it
// (MedMamba.py:403): one kernel reads the x half of in_proj's (B, H, W, 2 D) output IN PLACE (bf16 under autocast,
// channels contiguous), accumulates the 9 taps in fp32 and writes the (B, D, H, W) fp32 planes the cross-scan reads.
// SURVEY.md section 8(f) rank 1 (producer side).  Backward = one kernel: recompute the pre-activation, SiLU', input
// gradient (the transposed 3x3 stencil), and per-CTA partial sums of d(weight) / d(bias) added with fp32 atomics.
//
// Tiling: CTA = (batch item, 16 channels, band of 7 (forward) / 6 (backward) rows); thread = (channel = tid % 16, strip slot =
// tid / 16).  Shared input tile [pixel][16 channels] (channel fastest), staged with one 16-byte load per 8 channels of a
// pixel; every strip of 7 consecutive pixels is walked with 3 x 3 register windows (3 new shared loads per pixel, running
// pointers); results / gradients go through [channel][pixel] tiles (pitch == 2 mod 32: conflict free for 16 channels x 2
// strips) so that every global access to the (B, D, H, W) planes is a run of consecutive pixels of one plane.
// Measured on the way (ncu, stage 0, batch 64): the first version was latency-bound on its staging loops (~60 dependent 2- and
// 4-byte loads per thread) and spent half of its issue slots on index arithmetic; vector staging + running pointers took the
// backward from 1.53 to 1.16 ms and the forward from 0.63 to 0.34 ms per MedMamba-T step.
#include "common.cuh"

namespace b200 {

constexpr int DW_C = 16;     // channels per CTA

constexpr int DW_THREADS = 256;
constexpr int DW_MAXW = 64;  // widest image row supported by the shared tiles

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
constexpr int DW_TF = 7;      // rows per CTA in the forward: MedMamba's 56 / 28 / 14 / 7-row images split evenly
constexpr int DW_SEG = 7;     // pixels per sliding-window strip; odd, so the two strips a warp works on fall in different bank halves
// pitch of one channel's plane in a [channel][pixel] shared tile, == 2 (mod 32): a warp touches 16 channels x 2 pixels an odd
// distance apart, the channels then fall on the 16 even banks and the second pixel on the odd ones
__host__ __device__ __forceinline__ int dw_pitch2(int npix) { return ((npix + 29) / 32) * 32 + 2; }

// Stage rows [h_lo, h_lo + nrows) x columns [-halo, W + halo) of 16 channels into in_s[pixel][16] (zero outside the image).
// Vector path (all model shapes): 8 channels of one pixel per 16 / 32-byte load -- a few independent loads per thread instead of
// dozens of dependent 2-byte ones (the kernels were latency-bound on exactly this loop).
template <typename TI>
__device__ __forceinline__ void dw_stage_input(float* in_s, const TI* __restrict__ xin, int64_t pix_stride, int b, int c0, int D, int H, int W,
                                               int h_lo, int nrows, int halo) {
    const int WI = W + 2 * halo, tid = threadIdx.x;
    const bool vec = (D % DW_C) == 0 && (pix_stride % 8) == 0 && (reinterpret_cast<uintptr_t>(xin) & 15) == 0;
    if (vec) {
        const int npix = nrows * WI;
        for (int it = tid; it < npix * 2; it += DW_THREADS) {
            const int p = it >> 1, half = it & 1;
            const int r = p / WI, wc = p - r * WI;
            const int h = h_lo + r, w = wc - halo;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                const TI* src = xin + ((int64_t)(b * H + h) * W + w) * pix_stride + c0 + half * 8;
                if constexpr (sizeof(TI) == 2) {
                    const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(src));
                    const unsigned u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        v[2 * e] = __uint_as_float(u[e] << 16);
                        v[2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u);
                    }
                } else {
                    const float4 lo = __ldcs(reinterpret_cast<const float4*>(src)), hi = __ldcs(reinterpret_cast<const float4*>(src) + 1);
                    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
                }
            }
            float4* dst = reinterpret_cast<float4*>(in_s + p * DW_C + half * 8);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
        constexpr int NPS = DW_THREADS / DW_C;
        const int c = tid % DW_C, ps = tid / DW_C;
        const bool cok = c0 + c < D;
        for (int r = 0; r < nrows; ++r) {
            const int h = h_lo + r;
            const bool hok = cok && h >= 0 && h < H;
            const TI* src = xin + ((int64_t)(b * H + (hok ? h : 0)) * W + (ps - halo)) * pix_stride + c0 + c;
            float* dst = in_s + (r * WI + ps) * DW_C + c;
            for (int wc = ps; wc < WI; wc += NPS) {
                const int w = wc - halo;
                float v = 0.f;
                if (hok && w >= 0 && w < W) v = ldg_stream(src);
                *dst = v;
                src += NPS * pix_stride;
                dst += NPS * DW_C;
            }
        }
    }
}

// Forward.  Thread = (channel, strip slot); a strip is DW_SEG consecutive pixels of one row walked with a 3 x 3 register window
// (3 new shared loads per pixel, running pointers with compile-time offsets).
template <typename TI>
__global__ void __launch_bounds__(DW_THREADS, 3) dwconv_silu_fwd_kernel(const TI* __restrict__ xin, int64_t pix_stride, const float* __restrict__ wgt,
                                                                        const float* __restrict__ bias, float* __restrict__ out, int B, int D,
                                                                        int H, int W) {
    extern __shared__ __align__(16) float dsm[];
    const int WP = W + 2;                              // padded row length
    const int OP = dw_pitch2(DW_TF * W);
    float* in_s = dsm;                                 // [(TF + 2) * WP][DW_C]
    float* out_s = dsm + (DW_TF + 2) * WP * DW_C;      // [DW_C][OP]
    const int nbands = (H + DW_TF - 1) / DW_TF, ncg = (D + DW_C - 1) / DW_C;
    int bid = blockIdx.x;
    const int band = bid % nbands; bid /= nbands;
    const int cg = bid % ncg;
    const int b = bid / ncg;
    const int h0 = band * DW_TF, c0 = cg * DW_C;
    const int tid = threadIdx.x, c = tid % DW_C, ps = tid / DW_C;
    const bool cok = c0 + c < D;
    constexpr int NPS = DW_THREADS / DW_C;
    dw_stage_input<TI>(in_s, xin, pix_stride, b, c0, D, H, W, h0 - 1, DW_TF + 2, 1);
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = cok ? __ldg(wgt + (size_t)(c0 + c) * 9 + t) : 0.f;
    const float bv = (cok && bias) ? __ldg(bias + c0 + c) : 0.f;
    __syncthreads();
    const int rows = min(DW_TF, H - h0);
    {
        const float* inc = in_s + c;
        float* oc = out_s + c * OP;
        const int rowI = WP * DW_C;
        const int nseg = (W + DW_SEG - 1) / DW_SEG;
        for (int t = ps; t < rows * nseg; t += NPS) {
            const int r = t / nseg, w0 = (t - r * nseg) * DW_SEG;
            const int n = min(DW_SEG, W - w0);
            const float* i0 = inc + (r * WP + w0) * DW_C;   // output (r, w) reads in_s rows r..r+2, columns w..w+2
            const float* i1 = i0 + rowI;
            const float* i2 = i1 + rowI;
            float* o = oc + r * W + w0;
            float a0 = i0[0], a1 = i0[DW_C], b0 = i1[0], b1 = i1[DW_C], e0 = i2[0], e1 = i2[DW_C];
#define DW_FWD_STEP(q)                                                                                                    \
    {                                                                                                                     \
        const float a2 = i0[((q) + 2) * DW_C], b2 = i1[((q) + 2) * DW_C], e2 = i2[((q) + 2) * DW_C];                      \
        float acc = bv;                                                                                                   \
        acc = fmaf(k[0], a0, acc); acc = fmaf(k[1], a1, acc); acc = fmaf(k[2], a2, acc);                                  \
        acc = fmaf(k[3], b0, acc); acc = fmaf(k[4], b1, acc); acc = fmaf(k[5], b2, acc);                                  \
        acc = fmaf(k[6], e0, acc); acc = fmaf(k[7], e1, acc); acc = fmaf(k[8], e2, acc);                                  \
        o[q] = acc * sigmoidf_(acc);                                                                                      \
        a0 = a1; a1 = a2; b0 = b1; b1 = b2; e0 = e1; e1 = e2;                                                             \
    }
            // every model shape (W = 56 / 28 / 14 / 7) consists of full strips: the unrolled body turns the window rotation into
            // register renaming and the addresses into immediates (the rolled loop spent ~a third of its instructions on both)
            if (n == DW_SEG) {
#pragma unroll
                for (int q = 0; q < DW_SEG; ++q) DW_FWD_STEP(q)
            } else {
                for (int q = 0; q < n; ++q) DW_FWD_STEP(q)
            }
#undef DW_FWD_STEP
        }
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
// sigmoid from one MUFU.EX2 and one MUFU.RCP (relative error ~5e-7; the forward keeps the libm form)
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.f + ex2(-kLog2e * x)); }

constexpr int DW_TB = 6;      // rows per CTA in the backward
__host__ __device__ __forceinline__ int dw_gp(int W) { return dw_pitch2((DW_TB + 2) * (W + 2)); }

// Backward.  Thread = (channel, strip slot); every strip is DW_SEG consecutive pixels of one row walked with 3 x 3 register
// windows (3 new shared loads per window and pixel instead of 9): phase A recomputes the pre-activation on the band + halo 1 and
// turns g into d(pre); phase B produces the input gradient of the band and the CTA's share of d(weight), d(bias).
template <typename TI>
__global__ void __launch_bounds__(DW_THREADS, 3) dwconv_silu_bwd_kernel(const float* __restrict__ g, const TI* __restrict__ xin, int64_t pix_stride,
                                                                        const float* __restrict__ wgt, const float* __restrict__ bias,
                                                                        TI* __restrict__ dxin, float* __restrict__ dwgt, float* __restrict__ dbias,
                                                                        int B, int D, int H, int W) {
    extern __shared__ __align__(16) float dsm[];
    const int WI = W + 4, WG = W + 2;
    const int GP = dw_gp(W);
    float* in_s = dsm;                                  // input, halo 2: [(TB + 4) * WI][DW_C]   (channel fastest)
    float* dp_s = in_s + (DW_TB + 4) * WI * DW_C;       // d pre-activation, halo 1: [DW_C][GP]    (pixel fastest)
    float* red_s = in_s;                                // [10][DW_THREADS], reuses the input tile after the last phase
    const int nbands = (H + DW_TB - 1) / DW_TB, ncg = (D + DW_C - 1) / DW_C;
    int bid = blockIdx.x;
    const int band = bid % nbands; bid /= nbands;
    const int cg = bid % ncg;
    const int b = bid / ncg;
    const int h0 = band * DW_TB, c0 = cg * DW_C;
    const int tid = threadIdx.x, c = tid % DW_C, ps = tid / DW_C;
    const bool cok = c0 + c < D;
    constexpr int NPS = DW_THREADS / DW_C;
    dw_stage_input<TI>(in_s, xin, pix_stride, b, c0, D, H, W, h0 - 2, DW_TB + 4, 2);
    // upstream gradient, halo 1, from the planes
    if ((W & 3) == 0 && (D % DW_C) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int W4 = W >> 2, per_c = (DW_TB + 2) * W4;
        for (int it = tid; it < DW_C * per_c; it += DW_THREADS) {
            const int cc = it / per_c, rem = it - cc * per_c;
            const int r = rem / W4, w4 = rem - r * W4;
            const int h = h0 - 1 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h >= 0 && h < H) v = __ldcs(reinterpret_cast<const float4*>(g + ((size_t)(b * D + c0 + cc) * H + h) * W) + w4);
            float* dst = dp_s + cc * GP + r * WG + 4 * w4 + 1;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
        for (int it = tid; it < DW_C * (DW_TB + 2) * 2; it += DW_THREADS) {   // the two halo columns
            const int cc = it / ((DW_TB + 2) * 2), rem = it - cc * (DW_TB + 2) * 2;
            dp_s[cc * GP + (rem >> 1) * WG + ((rem & 1) ? WG - 1 : 0)] = 0.f;
        }
    } else {
        for (int cc = tid / 32; cc < DW_C; cc += DW_THREADS / 32) {
            const bool ok = c0 + cc < D;
            const float* src = g + (size_t)(b * D + (ok ? c0 + cc : 0)) * H * W;
            float* dst = dp_s + cc * GP;
            for (int r = 0; r < DW_TB + 2; ++r) {
                const int h = h0 - 1 + r;
                const bool hok = ok && h >= 0 && h < H;
                for (int wc = tid % 32; wc < WG; wc += 32) {
                    const int w = wc - 1;
                    dst[r * WG + wc] = (hok && w >= 0 && w < W) ? __ldcs(src + (size_t)h * W + w) : 0.f;
                }
            }
        }
    }
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = cok ? __ldg(wgt + (size_t)(c0 + c) * 9 + t) : 0.f;
    const float bv = (cok && bias) ? __ldg(bias + c0 + c) : 0.f;
    __syncthreads();
    float* dpc = dp_s + c * GP;
    const float* inc = in_s + c;
    // ---- phase A: g -> d pre = g * silu'(pre) on the halo-1 region; dp (r, wc) <-> in_s rows r..r+2, columns wc..wc+2.
    // All shared addresses are running pointers with compile-time offsets (index arithmetic was half of the issued instructions).
    const int rowI = WI * DW_C;                          // words between two in_s rows
    {
        const int nseg = (WG + DW_SEG - 1) / DW_SEG;
        for (int t = ps; t < (DW_TB + 2) * nseg; t += NPS) {
            const int r = t / nseg, wc0 = (t - r * nseg) * DW_SEG;
            const int n = min(DW_SEG, WG - wc0);
            const float* i0 = inc + (r * WI + wc0) * DW_C;
            const float* i1 = i0 + rowI;
            const float* i2 = i1 + rowI;
            float* dq = dpc + r * WG + wc0;
            float a0 = i0[0], a1 = i0[DW_C], b0 = i1[0], b1 = i1[DW_C], e0 = i2[0], e1 = i2[DW_C];
#define DW_BWD_A_STEP(q)                                                                                                  \
    {                                                                                                                     \
        const float a2 = i0[((q) + 2) * DW_C], b2 = i1[((q) + 2) * DW_C], e2 = i2[((q) + 2) * DW_C];                      \
        float pre = bv;                                                                                                   \
        pre = fmaf(k[0], a0, pre); pre = fmaf(k[1], a1, pre); pre = fmaf(k[2], a2, pre);                                  \
        pre = fmaf(k[3], b0, pre); pre = fmaf(k[4], b1, pre); pre = fmaf(k[5], b2, pre);                                  \
        pre = fmaf(k[6], e0, pre); pre = fmaf(k[7], e1, pre); pre = fmaf(k[8], e2, pre);                                  \
        const float sg = sigmoid_fast(pre);                                                                               \
        dq[q] *= sg * (1.f + pre * (1.f - sg)); /* zero outside the image: g was zero-filled there */                     \
        a0 = a1; a1 = a2; b0 = b1; b1 = b2; e0 = e1; e1 = e2;                                                             \
    }
            if (n == DW_SEG) {
#pragma unroll
                for (int q = 0; q < DW_SEG; ++q) DW_BWD_A_STEP(q)
            } else {
                for (int q = 0; q < n; ++q) DW_BWD_A_STEP(q)
            }
#undef DW_BWD_A_STEP
        }
    }
    __syncthreads();
    // ---- phase B: input gradient of the band's own pixels + this CTA's share of d(weight), d(bias)
    float dk[9], db = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) dk[t] = 0.f;
    {
        const int rows = min(DW_TB, H - h0);
        const int nseg = (W + DW_SEG - 1) / DW_SEG;
        for (int t = ps; t < rows * nseg; t += NPS) {
            const int r = t / nseg, w0 = (t - r * nseg) * DW_SEG;
            const int n = min(DW_SEG, W - w0);
            // image pixel (h0 + r, w): d(pre) window rows r..r+2, columns w..w+2 of dp; input window rows r+1..r+3, columns w+1..w+3 of in_s
            const float* d0p = dpc + r * WG + w0;
            const float* d1p = d0p + WG;
            const float* d2p = d1p + WG;
            const float* i0 = inc + ((r + 1) * WI + w0 + 1) * DW_C;
            const float* i1 = i0 + rowI;
            const float* i2 = i1 + rowI;
            float x00 = d0p[0], x01 = d0p[1], x10 = d1p[0], x11 = d1p[1], x20 = d2p[0], x21 = d2p[1];
            float y00 = i0[0], y01 = i0[DW_C], y10 = i1[0], y11 = i1[DW_C], y20 = i2[0], y21 = i2[DW_C];
            TI* o = dxin + ((int64_t)(b * H + h0 + r) * W + w0) * D + c0 + c;
            // out(h + 1 - dr, w + 1 - dc) used in(h, w) with tap (dr, dc): acc = sum k[dr][dc] * dp[r + 2 - dr][w + 2 - dc]
#define DW_BWD_B_STEP(q)                                                                                                  \
    {                                                                                                                     \
        const float x02 = d0p[(q) + 2], x12 = d1p[(q) + 2], x22 = d2p[(q) + 2];                                           \
        const float y02 = i0[((q) + 2) * DW_C], y12 = i1[((q) + 2) * DW_C], y22 = i2[((q) + 2) * DW_C];                   \
        float acc = k[0] * x22;                                                                                           \
        acc = fmaf(k[1], x21, acc); acc = fmaf(k[2], x20, acc);                                                           \
        acc = fmaf(k[3], x12, acc); acc = fmaf(k[4], x11, acc); acc = fmaf(k[5], x10, acc);                               \
        acc = fmaf(k[6], x02, acc); acc = fmaf(k[7], x01, acc); acc = fmaf(k[8], x00, acc);                               \
        const float dd = x11;                                                                                             \
        dk[0] = fmaf(dd, y00, dk[0]); dk[1] = fmaf(dd, y01, dk[1]); dk[2] = fmaf(dd, y02, dk[2]);                         \
        dk[3] = fmaf(dd, y10, dk[3]); dk[4] = fmaf(dd, y11, dk[4]); dk[5] = fmaf(dd, y12, dk[5]);                         \
        dk[6] = fmaf(dd, y20, dk[6]); dk[7] = fmaf(dd, y21, dk[7]); dk[8] = fmaf(dd, y22, dk[8]);                         \
        db += dd;                                                                                                         \
        if (cok) stg_stream(o + (int64_t)(q) * D, acc);                                                                   \
        x00 = x01; x01 = x02; x10 = x11; x11 = x12; x20 = x21; x21 = x22;                                                 \
        y00 = y01; y01 = y02; y10 = y11; y11 = y12; y20 = y21; y21 = y22;                                                 \
    }
            if (n == DW_SEG) {
#pragma unroll
                for (int q = 0; q < DW_SEG; ++q) DW_BWD_B_STEP(q)
            } else {
                for (int q = 0; q < n; ++q) DW_BWD_B_STEP(q)
            }
#undef DW_BWD_B_STEP
        }
    }
    __syncthreads();   // red_s aliases in_s
    // combine the 16 strip slots of each channel, one atomic per (channel, tap) and CTA
#pragma unroll
    for (int t = 0; t < 9; ++t) red_s[t * DW_THREADS + tid] = dk[t];
    red_s[9 * DW_THREADS + tid] = db;
    __syncthreads();
    if (tid < DW_C * 10) {
        const int cc = tid % DW_C, t = tid / DW_C;
        float sacc = 0.f;
#pragma unroll
        for (int q = 0; q < NPS; ++q) sacc += red_s[t * DW_THREADS + q * DW_C + cc];
        if (c0 + cc < D) {
            if (t < 9) atomicAdd(dwgt + (size_t)(c0 + cc) * 9 + t, sacc);
            else if (dbias) atomicAdd(dbias + c0 + cc, sacc);
        }
    }
}

static size_t dw_fwd_smem(int W) { return ((size_t)(DW_TF + 2) * (W + 2) * DW_C + (size_t)DW_C * dw_pitch2(DW_TF * W)) * sizeof(float); }
static size_t dw_bwd_smem(int W) {
    size_t in_words = (size_t)(DW_TB + 4) * (W + 4) * DW_C;
    if (in_words < 10 * DW_THREADS) in_words = 10 * DW_THREADS;   // the reduction tile reuses the input tile
    return (in_words + (size_t)DW_C * dw_gp(W)) * sizeof(float);
}
template <class K> static int dw_set_smem(K kernel, size_t bytes) {
    const cudaError_t e = func_attr_per_device((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

template <typename TI>
static int dw_fwd(const void* xin, int64_t ps, const float* w, const float* b, float* out, int B, int D, int H, int W, cudaStream_t st) {
    const size_t smem = dw_fwd_smem(DW_MAXW);
    const int rc0 = dw_set_smem(dwconv_silu_fwd_kernel<TI>, smem);   // per (kernel, device), see common.cuh
    if (rc0) return rc0;
    const long long grid = (long long)B * ((D + DW_C - 1) / DW_C) * ((H + DW_TF - 1) / DW_TF);
    dwconv_silu_fwd_kernel<TI><<<(unsigned)grid, DW_THREADS, dw_fwd_smem(W), st>>>((const TI*)xin, ps, w, b, out, B, D, H, W);
    return check_launch("dwconv_silu_fwd_kernel");
}
template <typename TI>
static int dw_bwd(const float* g, const void* xin, int64_t ps, const float* w, const float* b, void* dxin, float* dw, float* db, int B, int D,
                  int H, int W, cudaStream_t st) {
    const size_t smem = dw_bwd_smem(DW_MAXW);
    const int rc0 = dw_set_smem(dwconv_silu_bwd_kernel<TI>, smem);
    if (rc0) return rc0;
    const long long grid = (long long)B * ((D + DW_C - 1) / DW_C) * ((H + DW_TB - 1) / DW_TB);
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
