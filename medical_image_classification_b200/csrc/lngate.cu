// lngate.cu -- SS2D output stage: out = LayerNorm(y) * silu(z), forward and backward, sm_100a.
//
// Replaces the three element-wise passes the reference runs after the cross-merge (reference MedMamba.py:478-479:
// `y = self.out_norm(y); y = y * F.silu(z)`, LayerNorm over the d_inner channels, eps 1e-5) and their four autograd
// kernels (layer-norm input grad, the gamma/beta reduction, silu backward, mul backward).  SURVEY.md section 8(f) rank 1.
// y may be bf16 when z is absent (block pre-norm / PatchMerging norm on a bf16 residual stream; dy returns in bf16).
// HBM-bound streaming kernels: one warp per row, the row lives in registers (d_inner <= 1024), statistics by
// warp shuffles, one read of y / z / dout and one write of each output per element; d(weight), d(bias) are
// accumulated per lane over the rows of a warp, combined per CTA in shared memory and written as per-CTA partials.
//
//   xhat = (y - mean) * rstd      n = xhat * w + b      out = n * silu(z)
//   dn = dout * silu(z)           dz = dout * n * sigmoid(z) * (1 + z * (1 - sigmoid(z)))
//   dw = sum_rows dn * xhat       db = sum_rows dn
//   dy = rstd * (dn*w - mean_D(dn*w) - xhat * mean_D(dn*w * xhat))
#include <type_traits>

#include "common.cuh"

namespace b200 {

constexpr int LG_MAXD = 1024;          // columns held in registers: 32 per lane
constexpr int LG_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename TY, typename TZ, typename TO, int NPL, bool HAS_Z>
__global__ void __launch_bounds__(LG_WARPS * 32) ln_gate_fwd_kernel(const TY* __restrict__ y, int64_t y_row_stride, const TZ* __restrict__ z,
                                                                    int64_t z_row_stride,
                                                                    const float* __restrict__ w, const float* __restrict__ b,
                                                                    TO* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd,
                                                                    int64_t rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * LG_WARPS + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * LG_WARPS;
    const float invD = 1.f / (float)D;   // w / b are re-read per row (L1-resident): the row itself needs the registers
    for (int64_t r = warp; r < rows; r += nwarps) {
        float v[NPL], zv[HAS_Z ? NPL : 1];   // every global load of the row is issued before the first use (latency-bound otherwise)
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int c = lane + 32 * k;
            v[k] = c < D ? ldg_stream(y + r * y_row_stride + c) : 0.f;
            if constexpr (HAS_Z) zv[k] = c < D ? ldg_stream(z + r * z_row_stride + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) s += v[k];
        const float mu = warp_sum(s) * invD;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int c = lane + 32 * k;
            const float dlt = c < D ? v[k] - mu : 0.f;
            q += dlt * dlt;
        }
        const float rs = rsqrtf(warp_sum(q) * invD + eps);
        if (lane == 0) {
            mean[r] = mu;
            rstd[r] = rs;
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int c = lane + 32 * k;
            if (c < D) {
                const float n = (v[k] - mu) * rs * __ldg(w + c) + __ldg(b + c);
                if constexpr (HAS_Z) {
                    const float zz = zv[k];
                    stg_stream(out + r * D + c, n * zz * sigmoidf_(zz));
                } else {
                    stg_stream(out + r * D + c, n);
                }
            }
        }
    }
}

template <typename TY, typename TZ, typename TO, int NPL, bool HAS_Z>
__global__ void __launch_bounds__(LG_WARPS * 32) ln_gate_bwd_kernel(const TO* __restrict__ dout, const TY* __restrict__ y, int64_t y_row_stride,
                                                                    const TZ* __restrict__ z, int64_t z_row_stride, const float* __restrict__ w,
                                                                    const float* __restrict__ b, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, TY* __restrict__ dy, TZ* __restrict__ dz,
                                                                    float* __restrict__ dw_part, float* __restrict__ db_part, int64_t rows, int D) {
    __shared__ float red[LG_WARPS][32 * NPL];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * LG_WARPS + wid, nwarps = (int64_t)gridDim.x * LG_WARPS;
    float aw[NPL], ab[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        aw[k] = 0.f;
        ab[k] = 0.f;
    }
    const float invD = 1.f / (float)D;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
        float xh[NPL], g[NPL];   // y -> xhat, dout -> dn * w
        float zv[HAS_Z ? NPL : 1];
        float s1 = 0.f, s2 = 0.f;
        // every global load of the row is issued before the first use (one dependent round trip per row instead of NPL)
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int c = lane + 32 * k;
            xh[k] = c < D ? ldg_stream(y + r * y_row_stride + c) : 0.f;
            g[k] = c < D ? ldg_stream(dout + r * D + c) : 0.f;
            if constexpr (HAS_Z) zv[k] = c < D ? ldg_stream(z + r * z_row_stride + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int c = lane + 32 * k;
            const float yy = xh[k], go = g[k];
            xh[k] = 0.f;
            g[k] = 0.f;
            if (c < D) {
                const float wc = __ldg(w + c);
                xh[k] = (yy - mu) * rs;
                float dn = go;
                if constexpr (HAS_Z) {
                    const float zz = zv[k];
                    const float sg = sigmoidf_(zz);
                    const float n = xh[k] * wc + __ldg(b + c);
                    dn = go * zz * sg;
                    stg_stream(dz + r * D + c, go * n * sg * (1.f + zz * (1.f - sg)));
                }
                aw[k] += dn * xh[k];
                ab[k] += dn;
                g[k] = dn * wc;
                s1 += g[k];
                s2 += g[k] * xh[k];
            }
        }
        const float m1 = warp_sum(s1) * invD, m2 = warp_sum(s2) * invD;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int c = lane + 32 * k;
            if (c < D) stg_stream(dy + r * D + c, rs * (g[k] - m1 - xh[k] * m2));
        }
    }
    // d(weight), d(bias): combine the CTA's warps, write one partial row per CTA
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int k = 0; k < NPL; ++k) red[wid][lane + 32 * k] = pass ? ab[k] : aw[k];
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += LG_WARPS * 32) {
            float sacc = 0.f;
#pragma unroll
            for (int ww = 0; ww < LG_WARPS; ++ww) sacc += red[ww][c];
            (pass ? db_part : dw_part)[(size_t)blockIdx.x * D + c] = sacc;
        }
        __syncthreads();
    }
}

static int lg_grid(int64_t rows) {
    const int64_t want = (rows + LG_WARPS - 1) / LG_WARPS;
    return (int)(want < 148 * 4 ? (want < 1 ? 1 : want) : 148 * 4);
}

template <typename TY, typename TZ, typename TO, int NPL>
static int lg_fwd_launch(const void* y, int64_t ys, const void* z, int64_t zs, const float* w, const float* b, void* out, float* mean, float* rstd,
                         int64_t rows, int D, float eps, cudaStream_t st) {
    if constexpr (std::is_same<TY, float>::value) {
        if (z) {
            ln_gate_fwd_kernel<TY, TZ, TO, NPL, true><<<lg_grid(rows), LG_WARPS * 32, 0, st>>>((const TY*)y, ys, (const TZ*)z, zs, w, b, (TO*)out,
                                                                                               mean, rstd, rows, D, eps);
            return check_launch("ln_gate_fwd_kernel");
        }
    }
    ln_gate_fwd_kernel<TY, TZ, TO, NPL, false><<<lg_grid(rows), LG_WARPS * 32, 0, st>>>((const TY*)y, ys, (const TZ*)z, zs, w, b, (TO*)out, mean,
                                                                                        rstd, rows, D, eps);
    return check_launch("ln_gate_fwd_kernel");
}
template <typename TY, typename TZ, typename TO, int NPL>
static int lg_bwd_launch(const void* dout, const void* y, int64_t ys, const void* z, int64_t zs, const float* w, const float* b, const float* mean,
                         const float* rstd, void* dy, void* dz, float* dwp, float* dbp, int64_t rows, int D, cudaStream_t st) {
    if constexpr (std::is_same<TY, float>::value) {
        if (z) {
            ln_gate_bwd_kernel<TY, TZ, TO, NPL, true><<<lg_grid(rows), LG_WARPS * 32, 0, st>>>((const TO*)dout, (const TY*)y, ys, (const TZ*)z, zs, w,
                                                                                               b, mean, rstd, (TY*)dy, (TZ*)dz, dwp, dbp, rows, D);
            return check_launch("ln_gate_bwd_kernel");
        }
    }
    ln_gate_bwd_kernel<TY, TZ, TO, NPL, false><<<lg_grid(rows), LG_WARPS * 32, 0, st>>>((const TO*)dout, (const TY*)y, ys, (const TZ*)z, zs, w, b,
                                                                                        mean, rstd, (TY*)dy, (TZ*)dz, dwp, dbp, rows, D);
    return check_launch("ln_gate_bwd_kernel");
}

// dispatch on (y dtype, z dtype, out dtype, columns per lane); a bf16 y (the bf16 residual stream of an autocast model) is the
// plain-LayerNorm case only (z == NULL)
#define LG_NPL(FN, TY, TZ, TO, ...)                                    \
    do {                                                               \
        const int npl = (D + 31) / 32;                                 \
        if (npl <= 4) return FN<TY, TZ, TO, 4>(__VA_ARGS__);           \
        if (npl <= 8) return FN<TY, TZ, TO, 8>(__VA_ARGS__);           \
        if (npl <= 16) return FN<TY, TZ, TO, 16>(__VA_ARGS__);         \
        if (npl <= 24) return FN<TY, TZ, TO, 24>(__VA_ARGS__);         \
        return FN<TY, TZ, TO, 32>(__VA_ARGS__);                        \
    } while (0)
#define LG_DISPATCH(FN, ...)                                                                                       \
    do {                                                                                                           \
        if (y_dtype == B200_BF16) {                                                                                \
            B200_REQUIRE(z == nullptr, "b200_ln_gate: a bf16 y is supported without z only");                      \
            if (out_dtype == B200_BF16) LG_NPL(FN, __nv_bfloat16, __nv_bfloat16, __nv_bfloat16, __VA_ARGS__);      \
            if (out_dtype == B200_F32) LG_NPL(FN, __nv_bfloat16, float, float, __VA_ARGS__);                       \
        } else if (y_dtype == B200_F32) {                                                                          \
            if (z_dtype == B200_F32 && out_dtype == B200_F32) LG_NPL(FN, float, float, float, __VA_ARGS__);        \
            if (z_dtype == B200_BF16 && out_dtype == B200_BF16) LG_NPL(FN, float, __nv_bfloat16, __nv_bfloat16, __VA_ARGS__); \
            if (z_dtype == B200_BF16 && out_dtype == B200_F32) LG_NPL(FN, float, __nv_bfloat16, float, __VA_ARGS__); \
        }                                                                                                          \
        B200_REQUIRE(false, "b200_ln_gate: unsupported dtype combination y=%d z=%d out=%d", y_dtype, z_dtype, out_dtype); \
    } while (0)

}  // namespace b200

using namespace b200;

extern "C" int b200_ln_gate_grid(int64_t rows) { return lg_grid(rows); }

extern "C" int b200_ln_gate_fwd(const void* y, int32_t y_dtype, int64_t y_row_stride, const void* z, int64_t z_row_stride, int32_t z_dtype, const float* w,
                                const float* b, void* out, int32_t out_dtype, float* mean, float* rstd, int64_t rows, int32_t D, float eps,
                                b200_stream_t stream) {
    B200_REQUIRE(y && w && b && out && mean && rstd, "b200_ln_gate_fwd: NULL argument");
    if (!z) z_dtype = out_dtype == B200_BF16 ? B200_BF16 : B200_F32;   // plain LayerNorm: pick an instantiated pair
    B200_REQUIRE(rows >= 0 && D >= 1 && D <= LG_MAXD, "b200_ln_gate_fwd: D = %d outside [1, %d]", D, LG_MAXD);
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    LG_DISPATCH(lg_fwd_launch, y, y_row_stride, z, z_row_stride, w, b, out, mean, rstd, rows, D, eps, st);
}

extern "C" int b200_ln_gate_bwd(const void* dout, const void* y, int32_t y_dtype, int64_t y_row_stride, const void* z, int64_t z_row_stride,
                                int32_t z_dtype, const float* w, const float* b, int32_t out_dtype, const float* mean, const float* rstd, void* dy, void* dz,
                                float* dw_partial, float* db_partial, int64_t rows, int32_t D, b200_stream_t stream) {
    B200_REQUIRE(dout && y && w && b && mean && rstd && dy && dw_partial && db_partial, "b200_ln_gate_bwd: NULL argument");
    B200_REQUIRE((z == nullptr) == (dz == nullptr), "b200_ln_gate_bwd: dz must be given exactly when z is");
    if (!z) z_dtype = out_dtype == B200_BF16 ? B200_BF16 : B200_F32;
    B200_REQUIRE(rows >= 1 && D >= 1 && D <= LG_MAXD, "b200_ln_gate_bwd: D = %d outside [1, %d]", D, LG_MAXD);
    cudaStream_t st = (cudaStream_t)stream;
    LG_DISPATCH(lg_bwd_launch, dout, y, y_row_stride, z, z_row_stride, w, b, mean, rstd, dy, dz, dw_partial, db_partial, rows, D, st);
}
