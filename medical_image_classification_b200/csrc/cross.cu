// cross.cu -- SS2D cross-scan / cross-merge data movement (reference MedMamba.py:393-395,
// 420-424, 476-477; SSD twin SSD/MedSSD.py:332-336, 376-391; atrous twin
// CrossMamba/FusionMamba/models/cross.py:34-92, 139-190).
//
// The reference materialises four permuted copies of every channel plane (transpose, stack, flip,
// cat: ~18 plane-sized passes) and later un-permutes four outputs (~19 passes).  Here the two
// flipped directions never exist in memory (the scan kernel walks them backwards, sscan.cu
// rev_mask), so cross-scan is ONE pass that writes the row-major and the column-major image, and
// cross-merge is ONE pass that reads the four direction outputs and writes their sum in the
// (B, H, W, D) layout the following LayerNorm wants.  All kernels are pure HBM streams; tiles are
// turned through shared memory so that both the reads and the writes are contiguous runs.
//
// Kernel families in this file:
//   cross_scan_pack[_bwd], cross_merge[_bwd]      32 x 32 / 8 x 8-patch tiles, contiguous (B, 2, D, L) layout (standalone entry points)
//   cross_scan_pack_{v4,warp,plane,strided}       the pack for the fused SS2D core: x2 in a caller-given strided layout; 128-bit
//   cross_scan_unpack4_{v4,warp,plane,tile}       whole-plane, warp-per-plane (small planes), scalar whole-plane and tile variants;
//                                                 unpack4 = the whole adjoint (4 direction gradients + the projection's gradient)
//   cross_scan4 / ssd_merge4 (+ adjoints)         SSD twins with four materialised directions
//   atrous_scan / atrous_merge                    EfficientVMamba's four sub-lattice "directions" (each the other's adjoint)
#include "common.cuh"

namespace b200 {

// ---- x (B, D, H, W) -> x2 (B, 2, D, L): [0] = x, [1] = x^T -----------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cross_scan_pack_kernel(const T* __restrict__ x, T* __restrict__ x2, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int plane = blockIdx.x;  // b * D + d
    const int b = plane / D, d = plane % D;
    const size_t L = (size_t)H * W;
    const T* src = x + (size_t)plane * L;
    T* dst0 = x2 + ((size_t)b * 2 * D + d) * L;
    T* dst1 = x2 + ((size_t)b * 2 * D + D + d) * L;
    const int w0 = blockIdx.y * 32, h0 = blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int h = h0 + j, w = w0 + tx;
        if (h < H && w < W) {
            const T v = src[(size_t)h * W + w];
            dst0[(size_t)h * W + w] = v;
            tile[j][tx] = to_f32<T>(v);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int w = w0 + j, h = h0 + tx;
        if (h < H && w < W) dst1[(size_t)w * H + h] = from_f32<T>(tile[tx][j]);
    }
}

// ---- dx2 (B, 2, D, L) -> dx (B, D, H, W) = dx2[0] + dx2[1]^T ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cross_scan_pack_bwd_kernel(const T* __restrict__ dx2, T* __restrict__ dx, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int plane = blockIdx.x;
    const int b = plane / D, d = plane % D;
    const size_t L = (size_t)H * W;
    const T* s0 = dx2 + ((size_t)b * 2 * D + d) * L;
    const T* s1 = dx2 + ((size_t)b * 2 * D + D + d) * L;
    T* dst = dx + (size_t)plane * L;
    const int w0 = blockIdx.y * 32, h0 = blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int w = w0 + j, h = h0 + tx;
        if (h < H && w < W) tile[tx][j] = to_f32<T>(s1[(size_t)w * H + h]);
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int h = h0 + j, w = w0 + tx;
        if (h < H && w < W) dst[(size_t)h * W + w] = from_f32<T>(to_f32<T>(s0[(size_t)h * W + w]) + tile[j][tx]);
    }
}

// ---- ys (B, 4, D, L) -> y (B, L, D) ------------------------------------------------------------
// direction order of ys: 0 = row-major, 1 = row-major scanned backwards, 2 = column-major,
// 3 = column-major scanned backwards (all stored at their memory position, see sscan.cu).
// One CTA: 32 channels x an 8x8 spatial patch.
template <typename T>
__global__ void __launch_bounds__(256) cross_merge_kernel(const T* __restrict__ ys, T* __restrict__ y, int D, int H, int W) {
    __shared__ float acc[64][33];
    const int b = blockIdx.z;
    const int d0 = blockIdx.y * 32;
    const int tiles_w = (W + 7) >> 3;
    const int h0 = (blockIdx.x / tiles_w) * 8, w0 = (blockIdx.x % tiles_w) * 8;
    const size_t L = (size_t)H * W;
    const T* base = ys + (size_t)b * 4 * D * L;
    const int tid = threadIdx.x;
    // row-major directions: runs of 8 contiguous w
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = tid + k * 256;
        const int dd = idx >> 6, pos = idx & 63;
        const int hh = pos >> 3, ww = pos & 7;
        const int d = d0 + dd, h = h0 + hh, w = w0 + ww;
        float v = 0.f;
        if (d < D && h < H && w < W) {
            const size_t o = (size_t)d * L + (size_t)h * W + w;
            v = to_f32<T>(base[o]) + to_f32<T>(base[(size_t)D * L + o]);
        }
        acc[pos][dd] = v;
    }
    __syncthreads();
    // column-major directions: runs of 8 contiguous h
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = tid + k * 256;
        const int dd = idx >> 6, pos = idx & 63;
        const int ww = pos >> 3, hh = pos & 7;
        const int d = d0 + dd, h = h0 + hh, w = w0 + ww;
        if (d < D && h < H && w < W) {
            const size_t o = (size_t)d * L + (size_t)w * H + h;
            acc[hh * 8 + ww][dd] += to_f32<T>(base[(size_t)2 * D * L + o]) + to_f32<T>(base[(size_t)3 * D * L + o]);
        }
    }
    __syncthreads();
    // y[b][l][d]: 32 contiguous channels per position
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = tid + k * 256;
        const int pos = idx >> 5, dd = idx & 31;
        const int hh = pos >> 3, ww = pos & 7;
        const int d = d0 + dd, h = h0 + hh, w = w0 + ww;
        if (d < D && h < H && w < W) y[((size_t)b * L + (size_t)h * W + w) * D + d] = from_f32<T>(acc[pos][dd]);
    }
}

// ---- dy (B, L, D) -> dys2 (B, 2, D, L): [0] row-major planes, [1] column-major planes ----------
// (directions 0/1 share dys2[0], directions 2/3 share dys2[1]: the scan backward reads them with
//  dout_group_div = 2.)
template <typename T>
__global__ void __launch_bounds__(256) cross_merge_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dys, int D, int H, int W) {
    __shared__ float buf[64][33];
    const int b = blockIdx.z;
    const int d0 = blockIdx.y * 32;
    const int tiles_w = (W + 7) >> 3;
    const int h0 = (blockIdx.x / tiles_w) * 8, w0 = (blockIdx.x % tiles_w) * 8;
    const size_t L = (size_t)H * W;
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = tid + k * 256;
        const int pos = idx >> 5, dd = idx & 31;
        const int hh = pos >> 3, ww = pos & 7;
        const int d = d0 + dd, h = h0 + hh, w = w0 + ww;
        buf[pos][dd] = (d < D && h < H && w < W) ? to_f32<T>(dy[((size_t)b * L + (size_t)h * W + w) * D + d]) : 0.f;
    }
    __syncthreads();
    T* o0 = dys + (size_t)b * 2 * D * L;
    T* o1 = o0 + (size_t)D * L;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = tid + k * 256;
        const int dd = idx >> 6, pos = idx & 63;
        const int d = d0 + dd;
        {
            const int hh = pos >> 3, ww = pos & 7;
            const int h = h0 + hh, w = w0 + ww;
            if (d < D && h < H && w < W) o0[(size_t)d * L + (size_t)h * W + w] = from_f32<T>(buf[pos][dd]);
        }
        {
            const int ww = pos >> 3, hh = pos & 7;
            const int h = h0 + hh, w = w0 + ww;
            if (d < D && h < H && w < W) o1[(size_t)d * L + (size_t)w * H + h] = from_f32<T>(buf[hh * 8 + ww][dd]);
        }
    }
}

// ---- SSD twin (reference SSD/MedSSD.py:332-336, 376-391): the Mamba-2 operator has no reversed / shared-row mode and
//      mixes the four directions' B / C inside one state group, so all four orderings are materialised -- in ONE pass:
//      x (B, C, H, W) fp32 planes (batch stride given: a channel slice of a wider tensor is read in place)
//        -> x4 (B, 4, C, L):  k=0 row-major, k=1 column-major, k=2 / k=3 their time reversals (the reference's order)
template <bool BWD>
__global__ void __launch_bounds__(256) cross_scan4_kernel(const float* __restrict__ src, int64_t src_batch_stride, float* __restrict__ dst,
                                                          int C, int H, int W) {
    __shared__ float tile[32][33];
    const int plane = blockIdx.x;  // b * C + c
    const int b = plane / C, c = plane % C;
    const size_t L = (size_t)H * W;
    // forward: src = x planes, dst = x4;  backward: src = dx4, dst = dx planes (batch stride applies to the planes)
    const float* xin = BWD ? nullptr : src + (size_t)b * src_batch_stride + (size_t)c * L;
    float* xout = BWD ? dst + (size_t)b * src_batch_stride + (size_t)c * L : nullptr;
    const size_t q = ((size_t)b * 4 * C + c) * L;   // direction 0 plane; directions are C * L apart
    const float* g4 = BWD ? src + q : nullptr;
    float* o4 = BWD ? nullptr : dst + q;
    const size_t dstep = (size_t)C * L;
    const int w0 = blockIdx.y * 32, h0 = blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (!BWD) {
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int h = h0 + j, w = w0 + tx;
            if (h < H && w < W) {
                const size_t l = (size_t)h * W + w;
                const float v = __ldcs(xin + l);
                o4[l] = v;
                o4[2 * dstep + (L - 1 - l)] = v;
                tile[j][tx] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int w = w0 + j, h = h0 + tx;
            if (h < H && w < W) {
                const size_t l = (size_t)w * H + h;
                const float v = tile[tx][j];
                o4[dstep + l] = v;
                o4[3 * dstep + (L - 1 - l)] = v;
            }
        }
    } else {
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int w = w0 + j, h = h0 + tx;
            if (h < H && w < W) {
                const size_t l = (size_t)w * H + h;
                tile[tx][j] = __ldcs(g4 + dstep + l) + __ldcs(g4 + 3 * dstep + (L - 1 - l));
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int h = h0 + j, w = w0 + tx;
            if (h < H && w < W) {
                const size_t l = (size_t)h * W + w;
                xout[l] = __ldcs(g4 + l) + __ldcs(g4 + 2 * dstep + (L - 1 - l)) + tile[j][tx];
            }
        }
    }
}

// ---- y (B, L, 4, d) fp32 (the SSD output, direction-major heads) -> out (B, L, d):
//      out[b, hW+w] = y[b, hW+w, 0] + y[b, L-1-(hW+w), 2] + y[b, wH+h, 1] + y[b, L-1-(wH+h), 3]   (SSD/MedSSD.py:380-391)
//      and its adjoint (every y row receives exactly one out row): rows of d contiguous floats, gathers only
template <bool BWD>
__global__ void __launch_bounds__(256) ssd_merge4_kernel(const float* __restrict__ src, float* __restrict__ dst, int d, int H, int W, int64_t total) {
    const int L = H * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (!BWD) {   // idx over (b, l, j)
            const int j = (int)(idx % d);
            const int64_t bl = idx / d;
            const int l = (int)(bl % L);
            const int64_t b = bl / L;
            const int h = l / W, w = l % W, lt = w * H + h;
            const float* yb = src + (size_t)b * L * 4 * d + j;
            dst[idx] = __ldcs(yb + ((size_t)l * 4 + 0) * d) + __ldcs(yb + ((size_t)(L - 1 - l) * 4 + 2) * d) +
                       __ldcs(yb + ((size_t)lt * 4 + 1) * d) + __ldcs(yb + ((size_t)(L - 1 - lt) * 4 + 3) * d);
        } else {      // idx over (b, l', k, j): dy[b, l', k] = dout[b, l(l', k)]
            const int j = (int)(idx % d);
            int64_t r = idx / d;
            const int k = (int)(r & 3);
            r >>= 2;
            const int lp = (int)(r % L);
            const int64_t b = r / L;
            int l = (k & 2) ? L - 1 - lp : lp;            // undo the reversal
            if (k & 1) l = (l % H) * W + l / H;           // column-major index w H + h -> row-major h W + w
            dst[idx] = __ldcs(src + ((size_t)b * L + l) * d + j);
        }
    }
}

static int check_dims(const void* a, const void* b, int batch, int D, int H, int W, int dtype, const char* who) {
    B200_REQUIRE(a && b, "%s: NULL tensor", who);
    B200_REQUIRE(batch > 0 && D > 0 && H > 0 && W > 0, "%s: non-positive size", who);
    B200_REQUIRE((long long)batch * D < (1ll << 31) && batch < 65536 && (D + 31) / 32 < 65536, "%s: batch/D too large for the launch grid", who);
    B200_REQUIRE(dtype >= B200_F32 && dtype <= B200_F16, "%s: bad dtype %d", who, dtype);
    return 0;
}

template <typename T>
static int run_pack(const void* x, void* x2, int batch, int D, int H, int W, bool bwd, cudaStream_t st) {
    dim3 grid(batch * D, (W + 31) / 32, (H + 31) / 32);
    if (!bwd) cross_scan_pack_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (T*)x2, D, H, W);
    else cross_scan_pack_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (T*)x2, D, H, W);
    return check_launch(bwd ? "cross_scan_pack_bwd_kernel" : "cross_scan_pack_kernel");
}

template <typename T>
static int run_merge(const void* a, void* o, int batch, int D, int H, int W, bool bwd, cudaStream_t st) {
    dim3 grid(((H + 7) / 8) * ((W + 7) / 8), (D + 31) / 32, batch);
    if (!bwd) cross_merge_kernel<T><<<grid, 256, 0, st>>>((const T*)a, (T*)o, D, H, W);
    else cross_merge_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)a, (T*)o, D, H, W);
    return check_launch(bwd ? "cross_merge_bwd_kernel" : "cross_merge_kernel");
}

#define DISPATCH(fn, ...)                                        \
    switch (dtype) {                                             \
        case B200_F32: return fn<float>(__VA_ARGS__);            \
        case B200_BF16: return fn<__nv_bfloat16>(__VA_ARGS__);   \
        default: return fn<__half>(__VA_ARGS__);                 \
    }


// ---- strided twins for the fused SS2D core (cross.py::SS2DCoreFn): x2 lives in the layout (2, D, B, L) -- or any other
//      given by (batch, layout, row) element strides -- so that the x_proj / dt_proj contraction over it is ONE GEMM with
//      N = B * L columns.  x (B, D, H, W) f32 -> x2[b, 0, d] = plane, x2[b, 1, d] = plane^T.
__global__ void __launch_bounds__(256) cross_scan_pack_strided_kernel(const float* __restrict__ x, float* __restrict__ x2, int64_t sB, int64_t sI,
                                                                      int64_t sD, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int plane = blockIdx.x;  // b * D + d
    const int b = plane / D, d = plane % D;
    const size_t L = (size_t)H * W;
    const float* src = x + (size_t)plane * L;
    float* dst0 = x2 + (size_t)b * sB + (size_t)d * sD;
    float* dst1 = dst0 + sI;
    const int w0 = blockIdx.y * 32, h0 = blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int h = h0 + j, w = w0 + tx;
        if (h < H && w < W) {
            const float v = __ldcs(src + (size_t)h * W + w);
            dst0[(size_t)h * W + w] = v;
            tile[j][tx] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int w = w0 + j, h = h0 + tx;
        if (h < H && w < W) dst1[(size_t)w * H + h] = tile[tx][j];
    }
}

// ---- adjoint of the cross-scan for the fused core: dx (B, D, H, W) = du[b,0,d] + du[b,1,d] + gx2[b,0,d]
//      + (du[b,2,d] + du[b,3,d] + gx2[b,1,d])^T with du (B, 4, D, L) the scan's input gradient per direction (internal order, at
//      memory positions) and gx2 the projection's input gradient in x2's strided layout: one pass instead of a 4 -> 2 sum,
//      a gradient accumulation and an un-pack.
__global__ void __launch_bounds__(256) cross_scan_unpack4_kernel(const float* __restrict__ du, const float* __restrict__ gx2, int64_t sB,
                                                                 int64_t sI, int64_t sD, float* __restrict__ dx, int D, int H, int W) {
    __shared__ float tile[32][33];
    const int plane = blockIdx.x;
    const int b = plane / D, d = plane % D;
    const size_t L = (size_t)H * W;
    const float* u0 = du + ((size_t)b * 4 * D + d) * L;   // direction k at u0 + k * D * L
    const float* g0 = gx2 + (size_t)b * sB + (size_t)d * sD;
    const float* g1 = g0 + sI;
    float* dst = dx + (size_t)plane * L;
    const int w0 = blockIdx.y * 32, h0 = blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t DL = (size_t)D * L;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int w = w0 + j, h = h0 + tx;
        if (h < H && w < W) {
            const size_t o = (size_t)w * H + h;
            tile[tx][j] = __ldcs(u0 + 2 * DL + o) + __ldcs(u0 + 3 * DL + o) + __ldcs(g1 + o);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int h = h0 + j, w = w0 + tx;
        if (h < H && w < W) {
            const size_t o = (size_t)h * W + w;
            __stcs(dst + o, __ldcs(u0 + o) + __ldcs(u0 + DL + o) + __ldcs(g0 + o) + tile[j][tx]);
        }
    }
}


// ---- whole-plane variants (H * W <= CS_PLANE_MAX): one CTA turns an entire channel plane through shared memory, so every
//      thread has a dozen independent loads in flight and no tile is partially filled (the 32 x 32-tile kernels above run
//      at ~40 % of the HBM roofline on 56 x 56 planes: 76 % tile occupancy, two short dependent phases per CTA).
constexpr int CS_PLANE_MAX = 4096;
constexpr int CS_THREADS = 256;

__global__ void __launch_bounds__(CS_THREADS) cross_scan_pack_plane_kernel(const float* __restrict__ x, float* __restrict__ x2, int64_t sB, int64_t sI,
                                                                           int64_t sD, int D, int H, int W, int nplanes) {
    extern __shared__ float plane_s[];                 // [h][W + 1 | 1]: pitch odd
    const int L = H * W, PW = W | 1;
    for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
        const int b = plane / D, d = plane % D;
        const float* src = x + (size_t)plane * L;
        float* dst0 = x2 + (size_t)b * sB + (size_t)d * sD;
        float* dst1 = dst0 + sI;
#pragma unroll 4
        for (int p = threadIdx.x; p < L; p += CS_THREADS) {      // p = h * W + w
            const float v = __ldcs(src + p);
            const int h = p / W, w = p - h * W;
            dst0[p] = v;
            plane_s[h * PW + w] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int o = threadIdx.x; o < L; o += CS_THREADS) {      // o = w * H + h
            const int w = o / H, h = o - w * H;
            dst1[o] = plane_s[h * PW + w];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CS_THREADS) cross_scan_unpack4_plane_kernel(const float* __restrict__ du, const float* __restrict__ gx2, int64_t sB,
                                                                              int64_t sI, int64_t sD, float* __restrict__ dx, int D, int H, int W,
                                                                              int nplanes) {
    extern __shared__ float plane_s[];                 // [w][H | 1]
    const int L = H * W, PH = H | 1;
    const size_t DL = (size_t)D * L;
    for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
        const int b = plane / D, d = plane % D;
        const float* u0 = du + ((size_t)b * 4 * D + d) * L;
        const float* g0 = gx2 + (size_t)b * sB + (size_t)d * sD;
        const float* g1 = g0 + sI;
        float* dst = dx + (size_t)plane * L;
#pragma unroll 4
        for (int o = threadIdx.x; o < L; o += CS_THREADS) {      // o = w * H + h: the column-major directions
            const float v = __ldcs(u0 + 2 * DL + o) + __ldcs(u0 + 3 * DL + o) + __ldcs(g1 + o);
            const int w = o / H, h = o - w * H;
            plane_s[w * PH + h] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = threadIdx.x; p < L; p += CS_THREADS) {      // p = h * W + w
            const float v = __ldcs(u0 + p) + __ldcs(u0 + DL + p) + __ldcs(g0 + p);
            const int h = p / W, w = p - h * W;
            __stcs(dst + p, v + plane_s[w * PH + h]);
        }
        __syncthreads();
    }
}


// ---- 128-bit whole-plane variants (H, W multiples of 4, both <= 64; all MedMamba stages but the 7 x 7 one).
// Phase 1 moves float4 along the source's contiguous axis (16 lanes per row) into an aligned shared tile; phase 2 builds the
// float4 of the transposed side from 4 shared rows with lanes arranged 8 (contiguous axis of the tile) x 4 (float4 slots), so
// that shared reads are at most 2-way conflicted and every group of 4 lanes touches 64 contiguous bytes of global memory.
constexpr int CS_V_MAX = 64;
constexpr int CS_V_PITCH = CS_V_MAX + 4;

__global__ void __launch_bounds__(CS_THREADS) cross_scan_pack_v4_kernel(const float* __restrict__ x, float* __restrict__ x2, int64_t sB, int64_t sI,
                                                                        int64_t sD, int D, int H, int W, int nplanes) {
    __shared__ __align__(16) float S[CS_V_MAX * CS_V_PITCH];     // S[h][w]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lane = tid & 31, warp = tid >> 5, a_l = lane & 7, q_l = lane >> 3;
    const int W4 = W >> 2, H4 = H >> 2, L = H * W;
    for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
        const int b = plane / D, d = plane % D;
        const float4* src = reinterpret_cast<const float4*>(x + (size_t)plane * L);
        float* dst0 = x2 + (size_t)b * sB + (size_t)d * sD;
        float* dst1 = dst0 + sI;
        if (tx < W4) {
#pragma unroll 4
            for (int h = ty; h < H; h += 16) {
                const float4 v = __ldcs(src + h * W4 + tx);
                reinterpret_cast<float4*>(dst0)[h * W4 + tx] = v;
                *reinterpret_cast<float4*>(&S[h * CS_V_PITCH + 4 * tx]) = v;
            }
        }
        __syncthreads();
        const int w = warp * 8 + a_l;                            // 8 warps x 8 columns cover W <= 64
        if (w < W) {
#pragma unroll 4
            for (int h4 = q_l; h4 < H4; h4 += 4) {
                const float* col = &S[(4 * h4) * CS_V_PITCH + w];
                const float4 v = make_float4(col[0], col[CS_V_PITCH], col[2 * CS_V_PITCH], col[3 * CS_V_PITCH]);
                reinterpret_cast<float4*>(dst1 + (size_t)w * H)[h4] = v;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__global__ void __launch_bounds__(CS_THREADS) cross_scan_unpack4_v4_kernel(const float* __restrict__ du, const float* __restrict__ gx2, int64_t sB,
                                                                           int64_t sI, int64_t sD, float* __restrict__ dx, int D, int H, int W,
                                                                           int nplanes) {
    __shared__ __align__(16) float S[CS_V_MAX * CS_V_PITCH];     // S[w][h]: the column-major directions, summed
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lane = tid & 31, warp = tid >> 5, a_l = lane & 7, q_l = lane >> 3;
    const int W4 = W >> 2, H4 = H >> 2, L = H * W;
    const size_t DL = (size_t)D * L;
    for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
        const int b = plane / D, d = plane % D;
        const float* u0 = du + ((size_t)b * 4 * D + d) * L;
        const float* g0 = gx2 + (size_t)b * sB + (size_t)d * sD;
        const float* g1 = g0 + sI;
        if (tx < H4) {
            const float4* a = reinterpret_cast<const float4*>(u0 + 2 * DL);
            const float4* c = reinterpret_cast<const float4*>(u0 + 3 * DL);
            const float4* e = reinterpret_cast<const float4*>(g1);
#pragma unroll 4
            for (int w = ty; w < W; w += 16) {
                const int o = w * H4 + tx;
                *reinterpret_cast<float4*>(&S[w * CS_V_PITCH + 4 * tx]) = add4(add4(__ldcs(a + o), __ldcs(c + o)), __ldcs(e + o));
            }
        }
        __syncthreads();
        const int h = warp * 8 + a_l;                            // 8 warps x 8 rows cover H <= 64
        if (h < H) {
            const float4* a = reinterpret_cast<const float4*>(u0) + h * W4;
            const float4* c = reinterpret_cast<const float4*>(u0 + DL) + h * W4;
            const float4* e = reinterpret_cast<const float4*>(g0) + h * W4;
            float4* dst = reinterpret_cast<float4*>(dx + (size_t)plane * L) + h * W4;
#pragma unroll 4
            for (int w4 = q_l; w4 < W4; w4 += 4) {
                const float* col = &S[(4 * w4) * CS_V_PITCH + h];
                const float4 t = make_float4(col[0], col[CS_V_PITCH], col[2 * CS_V_PITCH], col[3 * CS_V_PITCH]);
                __stcs(dst + w4, add4(add4(add4(__ldcs(a + w4), __ldcs(c + w4)), __ldcs(e + w4)), t));
            }
        }
        __syncthreads();
    }
}


// ---- small planes (H * W <= CS_WARP_MAX, e.g. 14 x 14 and 7 x 7): one WARP per plane, eight planes in flight per CTA, private
//      shared tile per warp, no block barrier.
constexpr int CS_WARP_MAX = 1024;

__global__ void __launch_bounds__(CS_THREADS) cross_scan_pack_warp_kernel(const float* __restrict__ x, float* __restrict__ x2, int64_t sB, int64_t sI,
                                                                          int64_t sD, int D, int H, int W, int nplanes) {
    extern __shared__ float plane_s[];
    const int L = H * W, PW = W | 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* T = plane_s + warp * (H * PW);
    for (int plane = blockIdx.x * (CS_THREADS / 32) + warp; plane < nplanes; plane += gridDim.x * (CS_THREADS / 32)) {
        const int b = plane / D, d = plane % D;
        const float* src = x + (size_t)plane * L;
        float* dst0 = x2 + (size_t)b * sB + (size_t)d * sD;
        float* dst1 = dst0 + sI;
#pragma unroll 4
        for (int p = lane; p < L; p += 32) {
            const float v = __ldcs(src + p);
            const int h = p / W, w = p - h * W;
            dst0[p] = v;
            T[h * PW + w] = v;
        }
        __syncwarp();
#pragma unroll 4
        for (int o = lane; o < L; o += 32) {
            const int w = o / H, h = o - w * H;
            dst1[o] = T[h * PW + w];
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(CS_THREADS) cross_scan_unpack4_warp_kernel(const float* __restrict__ du, const float* __restrict__ gx2, int64_t sB,
                                                                             int64_t sI, int64_t sD, float* __restrict__ dx, int D, int H, int W,
                                                                             int nplanes) {
    extern __shared__ float plane_s[];
    const int L = H * W, PH = H | 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t DL = (size_t)D * L;
    float* T = plane_s + warp * (W * PH);
    for (int plane = blockIdx.x * (CS_THREADS / 32) + warp; plane < nplanes; plane += gridDim.x * (CS_THREADS / 32)) {
        const int b = plane / D, d = plane % D;
        const float* u0 = du + ((size_t)b * 4 * D + d) * L;
        const float* g0 = gx2 + (size_t)b * sB + (size_t)d * sD;
        const float* g1 = g0 + sI;
        float* dst = dx + (size_t)plane * L;
#pragma unroll 4
        for (int o = lane; o < L; o += 32) {
            const float v = __ldcs(u0 + 2 * DL + o) + __ldcs(u0 + 3 * DL + o) + __ldcs(g1 + o);
            const int w = o / H, h = o - w * H;
            T[w * PH + h] = v;
        }
        __syncwarp();
#pragma unroll 4
        for (int p = lane; p < L; p += 32) {
            const float v = __ldcs(u0 + p) + __ldcs(u0 + DL + p) + __ldcs(g0 + p);
            const int h = p / W, w = p - h * W;
            __stcs(dst + p, v + T[w * PH + h]);
        }
        __syncwarp();
    }
}


// ---- EfficientVMamba atrous scan / merge (reference CrossMamba/FusionMamba/models/cross.py:34-92, 139-190),
//      step 2: the four "directions" are the four sub-lattices of the image, (row parity, column parity) = (k & 1, k >> 1), k even
//      flattened row-major (i * W2 + j), k odd column-major (j * H2 + i), zero padded to even sizes.
//      scan:  xs[b, k, c, idx] = x[b, c, 2 i + (k & 1), 2 j + (k >> 1)];   merge is the inverse scatter (and each is the other's adjoint).
//      One CTA turns one plane through shared memory; both global sides are contiguous runs.
constexpr int AT_PLANE_MAX = 8192;   // pixels of the full-resolution plane held in shared memory (32 KB)

template <typename T>
__global__ void __launch_bounds__(CS_THREADS) atrous_scan_kernel(const T* __restrict__ x, T* __restrict__ xs, int C, int H, int W, int nplanes) {
    extern __shared__ float plane_s[];                 // [h][W | 1]
    const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1, L2 = H2 * W2, L = H * W, PW = W | 1;
    for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
        const int b = plane / C, c = plane % C;
        const T* src = x + (size_t)plane * L;
#pragma unroll 4
        for (int p = threadIdx.x; p < L; p += CS_THREADS) {
            const int h = p / W, w = p - h * W;
            plane_s[h * PW + w] = to_f32<T>(src[p]);
        }
        __syncthreads();
        for (int k = 0; k < 4; ++k) {
            T* dst = xs + (((size_t)b * 4 + k) * C + c) * L2;
            const int hr = k & 1, wr = k >> 1;
#pragma unroll 4
            for (int o = threadIdx.x; o < L2; o += CS_THREADS) {
                int i, j;
                if (hr == 0) { i = o / W2; j = o - i * W2; } else { j = o / H2; i = o - j * H2; }
                const int h = 2 * i + hr, w = 2 * j + wr;
                dst[o] = from_f32<T>((h < H && w < W) ? plane_s[h * PW + w] : 0.f);
            }
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(CS_THREADS) atrous_merge_kernel(const T* __restrict__ ys, T* __restrict__ y, int C, int H, int W, int nplanes) {
    extern __shared__ float plane_s[];                 // [4][L2]
    const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1, L2 = H2 * W2, L = H * W;
    for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
        const int b = plane / C, c = plane % C;
        for (int k = 0; k < 4; ++k) {
            const T* src = ys + (((size_t)b * 4 + k) * C + c) * L2;
#pragma unroll 4
            for (int o = threadIdx.x; o < L2; o += CS_THREADS) plane_s[k * L2 + o] = to_f32<T>(src[o]);
        }
        __syncthreads();
        T* dst = y + (size_t)plane * L;
#pragma unroll 4
        for (int p = threadIdx.x; p < L; p += CS_THREADS) {
            const int h = p / W, w = p - h * W;
            const int hr = h & 1, wr = w & 1, i = h >> 1, j = w >> 1;
            dst[p] = from_f32<T>(plane_s[(hr + 2 * wr) * L2 + (hr == 0 ? i * W2 + j : j * H2 + i)]);
        }
        __syncthreads();
    }
}

template <typename T>
static int run_atrous(const void* src, void* dst, int batch, int C, int H, int W, bool merge, cudaStream_t st) {
    const int nplanes = batch * C;
    const int grid1 = nplanes < 148 * 8 ? nplanes : 148 * 8;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    if (!merge) {
        atrous_scan_kernel<T><<<grid1, CS_THREADS, (size_t)H * (W | 1) * sizeof(float), st>>>((const T*)src, (T*)dst, C, H, W, nplanes);
        return check_launch("atrous_scan_kernel");
    }
    atrous_merge_kernel<T><<<grid1, CS_THREADS, (size_t)4 * H2 * W2 * sizeof(float), st>>>((const T*)src, (T*)dst, C, H, W, nplanes);
    return check_launch("atrous_merge_kernel");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_cross_scan_pack(const void* x, void* x2, int32_t batch, int32_t D, int32_t H, int32_t W, int32_t dtype,
                                    b200_stream_t stream) {
    if (int rc = check_dims(x, x2, batch, D, H, W, dtype, "b200_cross_scan_pack")) return rc;
    DISPATCH(run_pack, x, x2, batch, D, H, W, false, (cudaStream_t)stream)
}
extern "C" int b200_cross_scan_pack_bwd(const void* dx2, void* dx, int32_t batch, int32_t D, int32_t H, int32_t W, int32_t dtype,
                                        b200_stream_t stream) {
    if (int rc = check_dims(dx2, dx, batch, D, H, W, dtype, "b200_cross_scan_pack_bwd")) return rc;
    DISPATCH(run_pack, dx2, dx, batch, D, H, W, true, (cudaStream_t)stream)
}
extern "C" int b200_cross_merge(const void* ys, void* y, int32_t batch, int32_t D, int32_t H, int32_t W, int32_t dtype,
                                b200_stream_t stream) {
    if (int rc = check_dims(ys, y, batch, D, H, W, dtype, "b200_cross_merge")) return rc;
    DISPATCH(run_merge, ys, y, batch, D, H, W, false, (cudaStream_t)stream)
}
extern "C" int b200_cross_merge_bwd(const void* dy, void* dys, int32_t batch, int32_t D, int32_t H, int32_t W, int32_t dtype,
                                    b200_stream_t stream) {
    if (int rc = check_dims(dy, dys, batch, D, H, W, dtype, "b200_cross_merge_bwd")) return rc;
    DISPATCH(run_merge, dy, dys, batch, D, H, W, true, (cudaStream_t)stream)
}

extern "C" int b200_atrous_scan(const void* x, void* xs, int32_t batch, int32_t C, int32_t H, int32_t W, int32_t dtype, b200_stream_t stream) {
    if (int rc = check_dims(x, xs, batch, C, H, W, dtype, "b200_atrous_scan")) return rc;
    B200_REQUIRE((long long)(H + 1) * (W + 2) <= AT_PLANE_MAX, "b200_atrous_scan: plane %d x %d exceeds %d pixels", H, W, AT_PLANE_MAX);
    DISPATCH(run_atrous, x, xs, batch, C, H, W, false, (cudaStream_t)stream)
}
extern "C" int b200_atrous_merge(const void* ys, void* y, int32_t batch, int32_t C, int32_t H, int32_t W, int32_t dtype, b200_stream_t stream) {
    if (int rc = check_dims(ys, y, batch, C, H, W, dtype, "b200_atrous_merge")) return rc;
    B200_REQUIRE((long long)(H + 1) * (W + 2) <= AT_PLANE_MAX, "b200_atrous_merge: plane %d x %d exceeds %d pixels", H, W, AT_PLANE_MAX);
    DISPATCH(run_atrous, ys, y, batch, C, H, W, true, (cudaStream_t)stream)
}

extern "C" int b200_cross_scan_pack_strided(const float* x, float* x2, int64_t x2_batch_stride, int64_t x2_layout_stride, int64_t x2_row_stride,
                                            int32_t batch, int32_t D, int32_t H, int32_t W, b200_stream_t stream) {
    if (int rc = check_dims(x, x2, batch, D, H, W, B200_F32, "b200_cross_scan_pack_strided")) return rc;
    B200_REQUIRE(x2_row_stride >= (int64_t)H * W || x2_batch_stride >= (int64_t)H * W, "b200_cross_scan_pack_strided: overlapping rows");
    const bool v4 = H <= CS_V_MAX && W <= CS_V_MAX && (H & 3) == 0 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(x2) & 15) == 0 && (x2_batch_stride & 3) == 0 && (x2_layout_stride & 3) == 0 && (x2_row_stride & 3) == 0;
    if (v4) {
        const int nplanes = batch * D;
        const int grid1 = nplanes < 148 * 8 ? nplanes : 148 * 8;
        cross_scan_pack_v4_kernel<<<grid1, CS_THREADS, 0, (cudaStream_t)stream>>>(x, x2, x2_batch_stride, x2_layout_stride, x2_row_stride, D, H, W,
                                                                                 nplanes);
        return check_launch("cross_scan_pack_v4_kernel");
    }
    if (H * W <= CS_WARP_MAX && (size_t)(CS_THREADS / 32) * H * (W | 1) * sizeof(float) <= 48 * 1024) {
        const int nplanes = batch * D;
        const size_t smem = (size_t)(CS_THREADS / 32) * H * (W | 1) * sizeof(float);
        const int want = (nplanes + CS_THREADS / 32 - 1) / (CS_THREADS / 32);
        const int grid1 = want < 148 * 8 ? want : 148 * 8;
        cross_scan_pack_warp_kernel<<<grid1, CS_THREADS, smem, (cudaStream_t)stream>>>(x, x2, x2_batch_stride, x2_layout_stride, x2_row_stride, D, H,
                                                                                      W, nplanes);
        return check_launch("cross_scan_pack_warp_kernel");
    }
    if (H * W <= CS_PLANE_MAX) {
        const int nplanes = batch * D;
        const size_t smem = (size_t)H * (W | 1) * sizeof(float);
        const int grid1 = nplanes < 148 * 8 ? nplanes : 148 * 8;
        cross_scan_pack_plane_kernel<<<grid1, CS_THREADS, smem, (cudaStream_t)stream>>>(x, x2, x2_batch_stride, x2_layout_stride, x2_row_stride, D, H,
                                                                                       W, nplanes);
        return check_launch("cross_scan_pack_plane_kernel");
    }
    dim3 grid(batch * D, (W + 31) / 32, (H + 31) / 32);
    cross_scan_pack_strided_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x2, x2_batch_stride, x2_layout_stride, x2_row_stride, D, H, W);
    return check_launch("cross_scan_pack_strided_kernel");
}
extern "C" int b200_cross_scan_unpack4(const float* du, const float* gx2, int64_t x2_batch_stride, int64_t x2_layout_stride, int64_t x2_row_stride,
                                       float* dx, int32_t batch, int32_t D, int32_t H, int32_t W, b200_stream_t stream) {
    if (int rc = check_dims(du, dx, batch, D, H, W, B200_F32, "b200_cross_scan_unpack4")) return rc;
    B200_REQUIRE(gx2 != nullptr, "b200_cross_scan_unpack4: gx2 is NULL");
    const bool v4 = H <= CS_V_MAX && W <= CS_V_MAX && (H & 3) == 0 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(du) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(gx2) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (x2_batch_stride & 3) == 0 &&
                    (x2_layout_stride & 3) == 0 && (x2_row_stride & 3) == 0;
    if (v4) {
        const int nplanes = batch * D;
        const int grid1 = nplanes < 148 * 8 ? nplanes : 148 * 8;
        cross_scan_unpack4_v4_kernel<<<grid1, CS_THREADS, 0, (cudaStream_t)stream>>>(du, gx2, x2_batch_stride, x2_layout_stride, x2_row_stride, dx, D,
                                                                                    H, W, nplanes);
        return check_launch("cross_scan_unpack4_v4_kernel");
    }
    if (H * W <= CS_WARP_MAX && (size_t)(CS_THREADS / 32) * W * (H | 1) * sizeof(float) <= 48 * 1024) {
        const int nplanes = batch * D;
        const size_t smem = (size_t)(CS_THREADS / 32) * W * (H | 1) * sizeof(float);
        const int want = (nplanes + CS_THREADS / 32 - 1) / (CS_THREADS / 32);
        const int grid1 = want < 148 * 8 ? want : 148 * 8;
        cross_scan_unpack4_warp_kernel<<<grid1, CS_THREADS, smem, (cudaStream_t)stream>>>(du, gx2, x2_batch_stride, x2_layout_stride, x2_row_stride,
                                                                                         dx, D, H, W, nplanes);
        return check_launch("cross_scan_unpack4_warp_kernel");
    }
    if (H * W <= CS_PLANE_MAX) {
        const int nplanes = batch * D;
        const size_t smem = (size_t)W * (H | 1) * sizeof(float);
        const int grid1 = nplanes < 148 * 8 ? nplanes : 148 * 8;
        cross_scan_unpack4_plane_kernel<<<grid1, CS_THREADS, smem, (cudaStream_t)stream>>>(du, gx2, x2_batch_stride, x2_layout_stride, x2_row_stride,
                                                                                          dx, D, H, W, nplanes);
        return check_launch("cross_scan_unpack4_plane_kernel");
    }
    dim3 grid(batch * D, (W + 31) / 32, (H + 31) / 32);
    cross_scan_unpack4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(du, gx2, x2_batch_stride, x2_layout_stride, x2_row_stride, dx, D, H, W);
    return check_launch("cross_scan_unpack4_kernel");
}

extern "C" int b200_cross_scan4(const float* x, int64_t x_batch_stride, float* x4, int32_t batch, int32_t C, int32_t H, int32_t W,
                                b200_stream_t stream) {
    if (int rc = check_dims(x, x4, batch, C, H, W, B200_F32, "b200_cross_scan4")) return rc;
    dim3 grid(batch * C, (W + 31) / 32, (H + 31) / 32);
    cross_scan4_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_batch_stride, x4, C, H, W);
    return check_launch("cross_scan4_kernel");
}
extern "C" int b200_cross_scan4_bwd(const float* dx4, float* dx, int64_t dx_batch_stride, int32_t batch, int32_t C, int32_t H, int32_t W,
                                    b200_stream_t stream) {
    if (int rc = check_dims(dx4, dx, batch, C, H, W, B200_F32, "b200_cross_scan4_bwd")) return rc;
    dim3 grid(batch * C, (W + 31) / 32, (H + 31) / 32);
    cross_scan4_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(dx4, dx_batch_stride, dx, C, H, W);
    return check_launch("cross_scan4_bwd_kernel");
}
static unsigned merge4_grid(int64_t total) {
    const int64_t want = (total + 255) / 256;
    return (unsigned)(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
}
extern "C" int b200_ssd_merge4(const float* y, float* out, int32_t batch, int32_t d, int32_t H, int32_t W, b200_stream_t stream) {
    if (int rc = check_dims(y, out, batch, d, H, W, B200_F32, "b200_ssd_merge4")) return rc;
    const int64_t total = (int64_t)batch * H * W * d;
    ssd_merge4_kernel<false><<<merge4_grid(total), 256, 0, (cudaStream_t)stream>>>(y, out, d, H, W, total);
    return check_launch("ssd_merge4_kernel");
}
extern "C" int b200_ssd_merge4_bwd(const float* dout, float* dy, int32_t batch, int32_t d, int32_t H, int32_t W, b200_stream_t stream) {
    if (int rc = check_dims(dout, dy, batch, d, H, W, B200_F32, "b200_ssd_merge4_bwd")) return rc;
    const int64_t total = (int64_t)batch * H * W * 4 * d;
    ssd_merge4_kernel<true><<<merge4_grid(total), 256, 0, (cudaStream_t)stream>>>(dout, dy, d, H, W, total);
    return check_launch("ssd_merge4_bwd_kernel");
}
