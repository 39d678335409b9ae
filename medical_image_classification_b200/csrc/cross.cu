// cross.cu -- SS2D cross-scan / cross-merge data movement (reference MedMamba.py:393-395,
// 420-424, 476-477; SSD twin SSD/MedSSD.py:332-336, 376-391).
//
// The reference materialises four permuted copies of every channel plane (transpose, stack, flip,
// cat: ~18 plane-sized passes) and later un-permutes four outputs (~19 passes).  Here the two
// flipped directions never exist in memory (the scan kernel walks them backwards, sscan.cu
// rev_mask), so cross-scan is ONE pass that writes the row-major and the column-major image, and
// cross-merge is ONE pass that reads the four direction outputs and writes their sum in the
// (B, H, W, D) layout the following LayerNorm wants.  All kernels are pure HBM streams; tiles are
// turned through shared memory so that both the reads and the writes are contiguous runs.
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
