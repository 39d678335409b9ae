// sscan.cu -- Mamba-1 selective scan, forward and backward, for sm_100a.
//
// Replaces selective_scan_cuda.fwd/.bwd (reference CrossMamba/FusionMamba/selective_scan/
// selective_scan_fwd_kernel.cuh:67-303, selective_scan_bwd_kernel.cuh:75-489).  Not a port: the
// reference maps one CTA to one (batch, channel) row and runs a CUB block scan per state; here
//
//   * a CTA ("task") owns 32 consecutive channels of one (batch, group).  Lane = channel row,
//     warp = a quad of states: warp w keeps states 4w..4w+3 of its 32 rows in registers, so the
//     time recurrence is a plain in-register FFMA chain -- no block scan, no shuffles, and every
//     exp(delta*A) is evaluated exactly once;
//   * the group's B/C tile is staged in shared memory once per 32 rows and read back as
//     warp-wide broadcasts (the reference re-reads B/C from L2 for every row);
//   * u/delta tiles stream in through a double-buffered cp.async pipeline (16-byte copies when
//     the tensors allow it) and are read back by the owning lane with conflict-free 128-bit loads
//     (row pitch TT+4 floats);
//   * softplus(delta + bias) is evaluated once per element by a cooperative pre-pass, not once
//     per state;
//   * a group can be scanned in reverse time (rev_mask) and groups can share u / dout rows
//     (u_group_div, dout_group_div): that is all the SS2D cross-scan / cross-merge needs
//     (MedMamba.py:393-395, 420-424), the flipped copies never exist;
//   * backward = recompute: the forward stores the 16-float state every `ckpt_every` steps, the
//     backward walks chunks last->first, re-derives the forward states of one chunk in registers,
//     runs the adjoint recurrence, and reduces dB/dC over the 32 rows of the warp with a
//     reduce-scatter butterfly before issuing one fp32 atomic per (state, step).
//
// Binding pipes: forward = MUFU.EX2 (16 per element), backward = FP32 issue.
#include "common.cuh"

namespace b200 {

constexpr int NS = 16;   // states per task pass (4 warps x 4)
constexpr int SPW = 4;   // states per warp
constexpr int PB = 20;   // pitch of the transposed B/C tiles [t][n] (16 + 4: 16B-aligned rows, spreads banks)
constexpr int CKPT_EVERY = 8;  // steps between state checkpoints (== the backward chunk)

// ---- cp.async (LDGSTS) -----------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(float* smem, const float* gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = pred ? 4 : 0;  // src-size 0 => the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem, const float* gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Task {
    int b, g, r0, nrows, rpg, d0;  // d0 = first channel of the task
    bool rev;
};

__device__ __forceinline__ Task decode_task(const b200_sscan_fwd_params& p, int task) {
    Task t;
    t.rpg = p.dim / p.n_groups;
    const int tiles = (t.rpg + 31) >> 5;
    const int rt = task % tiles;
    const int bg = task / tiles;
    t.g = bg % p.n_groups;
    t.b = bg / p.n_groups;
    t.r0 = rt * 32;
    t.nrows = min(32, t.rpg - t.r0);
    t.d0 = t.g * t.rpg + t.r0;
    t.rev = (p.rev_mask >> t.g) & 1u;
    return t;
}

// Stage a [32 rows][TT] tile of a row-major activation tensor into shared memory in MEMORY order:
// column c <-> sequence position l_lo + c.  fp32 goes through cp.async (asynchronous; 16-byte copies
// when `vec`), 16-bit types are converted on the fly (synchronous).  Out-of-range elements become 0.
// NTHR is a compile-time constant so the index arithmetic folds away.
template <typename T, int TT, int NTHR>
__device__ __forceinline__ void stage_rows(float* tile, const T* base, int64_t row_stride, int nrows, int l_lo, int L,
                                           bool vec, int tid) {
    constexpr int TP = TT + 4;
    if constexpr (sizeof(T) == 4) {
        if (vec) {
#pragma unroll
            for (int idx0 = 0; idx0 < 32 * (TT / 4); idx0 += NTHR) {
                const int idx = idx0 + tid;
                if ((32 * (TT / 4)) % NTHR != 0 && idx >= 32 * (TT / 4)) break;
                const int rr = idx / (TT / 4), c = (idx % (TT / 4)) * 4;
                const int l = l_lo + c;
                const bool ok = rr < nrows && l >= 0 && l < L;  // L % 4 == 0 and l_lo % 4 == 0: all-or-nothing
                cp_async16(tile + rr * TP + c, ok ? (const float*)base + (size_t)rr * row_stride + l : (const float*)base, ok);
            }
        } else {
#pragma unroll
            for (int idx0 = 0; idx0 < 32 * TT; idx0 += NTHR) {
                const int idx = idx0 + tid;
                const int rr = idx / TT, c = idx % TT;
                const int l = l_lo + c;
                const bool ok = rr < nrows && l >= 0 && l < L;
                cp_async4(tile + rr * TP + c, ok ? (const float*)base + (size_t)rr * row_stride + l : (const float*)base, ok);
            }
        }
    } else {
#pragma unroll
        for (int idx0 = 0; idx0 < 32 * TT; idx0 += NTHR) {
            const int idx = idx0 + tid;
            const int rr = idx / TT, c = idx % TT;
            const int l = l_lo + c;
            const bool ok = rr < nrows && l >= 0 && l < L;
            tile[rr * TP + c] = ok ? ldg_stream(base + (size_t)rr * row_stride + l) : 0.f;
        }
    }
}

// Stage the group's [N][TT] slice of B or C transposed to [TT][PB] (state index fastest).
template <typename T, int TT, int NTHR>
__device__ __forceinline__ void stage_bc(float* tile, const T* base, int64_t state_stride, int N, int l_lo, int L, int tid) {
#pragma unroll
    for (int idx0 = 0; idx0 < NS * TT; idx0 += NTHR) {
        const int idx = idx0 + tid;
        const int n = idx / TT, c = idx % TT;
        const int l = l_lo + c;
        const bool ok = n < N && l >= 0 && l < L;
        if constexpr (sizeof(T) == 4) {
            cp_async4(tile + c * PB + n, ok ? (const float*)base + (size_t)n * state_stride + l : (const float*)base, ok);
        } else {
            tile[c * PB + n] = ok ? to_f32<T>(__ldg(base + (size_t)n * state_stride + l)) : 0.f;
        }
    }
}

template <bool REV> __device__ __forceinline__ float4 ld4(const float* p) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    return REV ? make_float4(v.w, v.z, v.y, v.x) : v;
}
template <bool REV> __device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = REV ? make_float4(d, c, b, a) : make_float4(a, b, c, d);
}
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }

template <typename T>
__device__ __forceinline__ bool can_vectorize(const void* p, int64_t row_stride, int64_t batch_stride, int64_t group_stride,
                                              int L) {
    return sizeof(T) == 4 && (L & 3) == 0 && (row_stride & 3) == 0 && (batch_stride & 3) == 0 && (group_stride & 3) == 0 &&
           (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
template <int TT>
struct FwdSmem {
    static constexpr int TP = TT + 4;
    float u[2][32 * TP];     // raw u tiles (double buffered)
    float d[2][32 * TP];     // raw delta tiles -> softplus(delta + bias) in place
    float B[2][TT * PB];     // [c][n]
    float C[2][TT * PB];
    float du[32 * TP];       // delta' * u
    float y[4][32 * TP];     // per-warp partial outputs
    float bias[32], D[32];
};

template <typename T, int TT, int NW, bool HAS_Z, bool REV>
__device__ __forceinline__ void sscan_fwd_body(const b200_sscan_fwd_params& p, const Task& t, FwdSmem<TT>& sm) {
    constexpr int TP = FwdSmem<TT>::TP;
    constexpr int NTHR = NW * 32;
    const int tid = threadIdx.x;
    const int lane = tid & 31, w = tid >> 5;
    const int L = p.seqlen, N = p.dstate;
    const bool row_ok = lane < t.nrows;
    const int d_lane = t.d0 + lane;

    // this warp's 4 states as two packed pairs
    float2 A2[2], x[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int n = w * SPW + 2 * j;
        A2[j].x = (row_ok && n < N) ? __ldg(p.A + (size_t)d_lane * N + n) * kLog2e : 0.f;
        A2[j].y = (row_ok && n + 1 < N) ? __ldg(p.A + (size_t)d_lane * N + n + 1) * kLog2e : 0.f;
        x[j] = make_float2(0.f, 0.f);
    }
    if (tid < 32) {
        sm.D[tid] = (p.D && row_ok) ? __ldg(p.D + d_lane) : 0.f;
        sm.bias[tid] = (p.delta_bias && row_ok) ? __ldg(p.delta_bias + d_lane) : 0.f;
    }
    const bool softplus = p.delta_softplus != 0;

    const T* u_base = (const T*)p.u + (size_t)t.b * p.u_batch_stride + (size_t)(t.g / p.u_group_div) * p.u_group_stride +
                      (size_t)t.r0 * p.u_row_stride;
    const T* d_base = (const T*)p.delta + (size_t)t.b * p.delta_batch_stride + (size_t)t.d0 * p.delta_row_stride;
    T* o_base = (T*)p.out + (size_t)t.b * p.out_batch_stride + (size_t)t.d0 * p.out_row_stride;
    const T* z_base = HAS_Z ? (const T*)p.z + (size_t)t.b * p.z_batch_stride + (size_t)t.d0 * p.z_row_stride : nullptr;
    const T* B_base = (const T*)p.B + (size_t)t.b * p.B_batch_stride + (size_t)t.g * p.B_group_stride;
    const T* C_base = (const T*)p.C + (size_t)t.b * p.C_batch_stride + (size_t)t.g * p.C_group_stride;
    const bool vec_u = can_vectorize<T>(p.u, p.u_row_stride, p.u_batch_stride, p.u_group_stride, L);
    const bool vec_d = can_vectorize<T>(p.delta, p.delta_row_stride, p.delta_batch_stride, 0, L);
    const bool vec_o = can_vectorize<T>(p.out, p.out_row_stride, p.out_batch_stride, 0, L) && !HAS_Z;

    const int nck = p.ckpt ? (L + CKPT_EVERY - 1) / CKPT_EVERY : 0;
    float* ck = p.ckpt ? p.ckpt + (size_t)blockIdx.x * (size_t)(nck - 1) * NS * 32 : nullptr;

    const int ntiles = (L + TT - 1) / TT;
    auto l_lo_of = [&](int tile) { return REV ? L - (tile + 1) * TT : tile * TT; };
    auto prefetch = [&](int tile) {
        if (tile < ntiles) {
            const int buf = tile & 1, l_lo = l_lo_of(tile);
            stage_rows<T, TT, NTHR>(sm.u[buf], u_base, p.u_row_stride, t.nrows, l_lo, L, vec_u, tid);
            stage_rows<T, TT, NTHR>(sm.d[buf], d_base, p.delta_row_stride, t.nrows, l_lo, L, vec_d, tid);
            stage_bc<T, TT, NTHR>(sm.B[buf], B_base, p.B_state_stride, N, l_lo, L, tid);
            stage_bc<T, TT, NTHR>(sm.C[buf], C_base, p.C_state_stride, N, l_lo, L, tid);
        }
        cp_async_commit();  // (possibly empty) group: keeps the wait_group arithmetic uniform
    };
    prefetch(0);
    prefetch(1);

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        const int s0 = tile * TT;
        const int nv = min(TT, L - s0);
        const int l_lo = l_lo_of(tile);
        cp_async_wait<1>();
        __syncthreads();  // (1) tile landed; previous tile's output store finished reading sm.y
        // ---- pre-pass: delta' = softplus(delta + bias) once per element; du = delta' * u ----
#pragma unroll
        for (int idx0 = 0; idx0 < 32 * (TT / 4); idx0 += NTHR) {
            const int idx = idx0 + tid;
            const int rr = idx / (TT / 4), c = (idx % (TT / 4)) * 4;
            const float4 dr = *reinterpret_cast<const float4*>(&sm.d[buf][rr * TP + c]);
            const float4 ur = *reinterpret_cast<const float4*>(&sm.u[buf][rr * TP + c]);
            const float bias = sm.bias[rr];
            const float draw[4] = {dr.x, dr.y, dr.z, dr.w};
            const float uraw[4] = {ur.x, ur.y, ur.z, ur.w};
            float dl[4], dq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int l = l_lo + c + k;
                float v = draw[k] + bias;
                if (softplus) v = softplus_sigmoid(v).sp;
                if (l < 0 || l >= L) v = 0.f;  // out-of-range steps become the identity: a = 1, b = 0
                dl[k] = v;
                dq[k] = v * uraw[k];
            }
            *reinterpret_cast<float4*>(&sm.d[buf][rr * TP + c]) = make_float4(dl[0], dl[1], dl[2], dl[3]);
            *reinterpret_cast<float4*>(&sm.du[rr * TP + c]) = make_float4(dq[0], dq[1], dq[2], dq[3]);
        }
        __syncthreads();  // (2)
        // ---- recurrence: lane = row, this warp's 4 states (2 packed pairs), 4 steps per iteration ----
        const float Dv = (w == 0) ? sm.D[lane] : 0.f;
#pragma unroll
        for (int i4 = 0; i4 < TT / 4; ++i4) {
            const int cb = REV ? TT - 4 - 4 * i4 : 4 * i4;  // memory-order column of this step quad
            if (i4 * 4 < nv) {                              // warp-uniform
                const int sb = s0 + i4 * 4;
                if ((i4 & 1) == 0 && ck != nullptr && sb > 0) {  // sb % CKPT_EVERY == 0
                    float* dst = ck + ((size_t)(sb / CKPT_EVERY - 1) * NS + w * SPW) * 32 + lane;
                    dst[0] = x[0].x; dst[32] = x[0].y; dst[64] = x[1].x; dst[96] = x[1].y;
                }
                const float4 d4 = ld4<REV>(&sm.d[buf][lane * TP + cb]);
                const float4 q4 = ld4<REV>(&sm.du[lane * TP + cb]);
                float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w == 0) u4 = ld4<REV>(&sm.u[buf][lane * TP + cb]);
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
                const float dq[4] = {q4.x, q4.y, q4.z, q4.w};
                const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
                float yy[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = REV ? TT - 1 - (i4 * 4 + k) : i4 * 4 + k;
                    const float4 b4 = *reinterpret_cast<const float4*>(&sm.B[buf][c * PB + w * SPW]);
                    const float4 c4 = *reinterpret_cast<const float4*>(&sm.C[buf][c * PB + w * SPW]);
                    const float2 d2 = dup2(dd[k]), q2 = dup2(dq[k]);
                    const float2 e0 = __fmul2_rn(d2, A2[0]), e1 = __fmul2_rn(d2, A2[1]);
                    const float2 a0 = make_float2(ex2(e0.x), ex2(e0.y)), a1 = make_float2(ex2(e1.x), ex2(e1.y));
                    x[0] = __ffma2_rn(a0, x[0], __fmul2_rn(q2, make_float2(b4.x, b4.y)));
                    x[1] = __ffma2_rn(a1, x[1], __fmul2_rn(q2, make_float2(b4.z, b4.w)));
                    float2 y2 = __fmul2_rn(make_float2(c4.x, c4.y), x[0]);
                    y2 = __ffma2_rn(make_float2(c4.z, c4.w), x[1], y2);
                    yy[k] = fmaf(Dv, uu[k], y2.x + y2.y);
                }
                st4<REV>(&sm.y[w][lane * TP + cb], yy[0], yy[1], yy[2], yy[3]);
            }
        }
        __syncthreads();  // (3) partial outputs ready; raw buffers of this tile are dead
        prefetch(tile + 2);
        // ---- out = sum of the warps' partials (* silu(z)), written at its memory position ----
        if (vec_o) {
#pragma unroll
            for (int idx0 = 0; idx0 < 32 * (TT / 4); idx0 += NTHR) {
                const int idx = idx0 + tid;
                const int rr = idx / (TT / 4), c = (idx % (TT / 4)) * 4;
                const int l = l_lo + c;
                if (rr < t.nrows && l >= 0 && l < L) {
                    float4 v = *reinterpret_cast<const float4*>(&sm.y[0][rr * TP + c]);
#pragma unroll
                    for (int ww = 1; ww < NW; ++ww) {
                        const float4 o = *reinterpret_cast<const float4*>(&sm.y[ww][rr * TP + c]);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    __stcs(reinterpret_cast<float4*>((float*)o_base + (size_t)rr * p.out_row_stride + l), v);
                }
            }
        } else {
#pragma unroll
            for (int idx0 = 0; idx0 < 32 * TT; idx0 += NTHR) {
                const int idx = idx0 + tid;
                const int rr = idx / TT, c = idx % TT;
                const int l = l_lo + c;
                if (rr < t.nrows && l >= 0 && l < L) {
                    float v = sm.y[0][rr * TP + c];
#pragma unroll
                    for (int ww = 1; ww < NW; ++ww) v += sm.y[ww][rr * TP + c];
                    if (HAS_Z) {
                        const float zz = ldg_stream(z_base + (size_t)rr * p.z_row_stride + l);
                        v *= zz * sigmoidf_(zz);
                    }
                    stg_stream(o_base + (size_t)rr * p.out_row_stride + l, v);
                }
            }
        }
    }
    cp_async_wait<0>();
    if (p.last_state != nullptr && row_ok) {
        float* ls = p.last_state + ((size_t)t.b * p.dim + d_lane) * N;
        const float xs[4] = {x[0].x, x[0].y, x[1].x, x[1].y};
#pragma unroll
        for (int j = 0; j < SPW; ++j)
            if (w * SPW + j < N) ls[w * SPW + j] = xs[j];
    }
}

template <typename T, int TT, int NW, bool HAS_Z>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 6 : 8) sscan_fwd_kernel(const __grid_constant__ b200_sscan_fwd_params p) {
    __shared__ __align__(16) FwdSmem<TT> sm;
    const Task t = decode_task(p, blockIdx.x);
    if (t.rev) sscan_fwd_body<T, TT, NW, HAS_Z, true>(p, t, sm);
    else sscan_fwd_body<T, TT, NW, HAS_Z, false>(p, t, sm);
}

// ----------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------
// Sum v[0..TC) over the 32 lanes of the warp with a reduce-scatter butterfly: each halving step
// exchanges half of the remaining values, so the whole reduction costs ~TC shuffles instead of
// 5*TC.  On return every lane holds the 32-row total of item `rs_item<TC>(lane)`.
template <int TC> __device__ __forceinline__ int rs_item(int lane);
template <> __device__ __forceinline__ int rs_item<8>(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

template <int M> __device__ __forceinline__ float rs_step(float (&v)[M], int lane, int bit) {
    // M values -> M/2 values across the lane pair (lane ^ bit); lanes with `bit` set keep the upper half
    const bool up = (lane & bit) != 0;
    if constexpr (M == 1) {
        return v[0];
    } else {
        float h[M / 2];
#pragma unroll
        for (int j = 0; j < M / 2; ++j) {
            const float keep = up ? v[j + M / 2] : v[j];
            const float send = up ? v[j] : v[j + M / 2];
            h[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        return rs_step<M / 2>(h, lane, bit >> 1);
    }
}

template <int TC> __device__ __forceinline__ float reduce_scatter(float (&v)[TC], int lane) {
    static_assert(TC == 8, "reduce_scatter is instantiated for 8-step chunks");
    float r = rs_step<TC>(v, lane, 16);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

template <int TC, bool HAS_Z>
struct BwdSmem {
    static constexpr int TP = TC + 4;
    static constexpr int SP = (HAS_Z ? 3 : 2) * TC + 4;  // pitch of the per-warp partial sums
    float u[2][32 * TP];    // u tile
    float d[2][32 * TP];    // raw delta tile -> delta' in place
    float g[2][32 * TP];    // dout tile -> dout * silu(z) in place
    float z[HAS_Z ? 2 : 1][HAS_Z ? 32 * TP : 4];  // z tile -> dout * d silu(z)/dz in place
    float B[2][TC * PB];
    float C[2][TC * PB];
    float ck[2][NS * 32];   // state entering the chunk, [n][lane]
    float sg[32 * TP];      // sigmoid(delta + bias) = d softplus: chain-rule factor of ddelta
    float part[4][32 * SP]; // per-warp partial s1 | s2 (| y) in memory-order columns
    float bias[32], D[32];
};

template <typename T, int TC, int NW, bool HAS_Z, bool REV>
__device__ __forceinline__ void sscan_bwd_body(const b200_sscan_bwd_params& q, const Task& t, BwdSmem<TC, HAS_Z>& sm) {
    const b200_sscan_fwd_params& p = q.f;
    constexpr int TP = BwdSmem<TC, HAS_Z>::TP;
    constexpr int SP = BwdSmem<TC, HAS_Z>::SP;
    constexpr int NTHR = NW * 32;
    constexpr int EP = (32 * (TC / 4) + NTHR - 1) / NTHR;  // passes of the (4 steps per thread) epilogue
    const int tid = threadIdx.x;
    const int lane = tid & 31, w = tid >> 5;
    const int L = p.seqlen, N = p.dstate;
    const bool row_ok = lane < t.nrows;
    const int d_lane = t.d0 + lane;

    float2 An[2], A2[2], h[2], dA[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int n = w * SPW + 2 * j;
        An[j].x = (row_ok && n < N) ? __ldg(p.A + (size_t)d_lane * N + n) : 0.f;
        An[j].y = (row_ok && n + 1 < N) ? __ldg(p.A + (size_t)d_lane * N + n + 1) : 0.f;
        A2[j] = make_float2(An[j].x * kLog2e, An[j].y * kLog2e);
        h[j] = make_float2(0.f, 0.f);
        dA[j] = make_float2(0.f, 0.f);
    }
    if (tid < 32) {
        sm.D[tid] = (p.D && row_ok) ? __ldg(p.D + d_lane) : 0.f;
        sm.bias[tid] = (p.delta_bias && row_ok) ? __ldg(p.delta_bias + d_lane) : 0.f;
    }
    const bool softplus = p.delta_softplus != 0;
    float dD_acc[EP], dbias_acc[EP];  // this thread's share of the per-row sums (its rows are fixed across chunks)
#pragma unroll
    for (int k = 0; k < EP; ++k) { dD_acc[k] = 0.f; dbias_acc[k] = 0.f; }

    const T* u_base = (const T*)p.u + (size_t)t.b * p.u_batch_stride + (size_t)(t.g / p.u_group_div) * p.u_group_stride +
                      (size_t)t.r0 * p.u_row_stride;
    const T* d_base = (const T*)p.delta + (size_t)t.b * p.delta_batch_stride + (size_t)t.d0 * p.delta_row_stride;
    const T* g_base = (const T*)q.dout + (size_t)t.b * q.dout_batch_stride +
                      (size_t)(t.g / (int)q.dout_group_div) * q.dout_group_stride + (size_t)t.r0 * q.dout_row_stride;
    const T* z_base = HAS_Z ? (const T*)p.z + (size_t)t.b * p.z_batch_stride + (size_t)t.d0 * p.z_row_stride : nullptr;
    T* du_base = (T*)q.du + (size_t)t.b * q.du_batch_stride + (size_t)t.d0 * q.du_row_stride;
    T* dd_base = (T*)q.ddelta + (size_t)t.b * q.ddelta_batch_stride + (size_t)t.d0 * q.ddelta_row_stride;
    T* dz_base = HAS_Z ? (T*)q.dz + (size_t)t.b * q.dz_batch_stride + (size_t)t.d0 * q.dz_row_stride : nullptr;
    const T* B_base = (const T*)p.B + (size_t)t.b * p.B_batch_stride + (size_t)t.g * p.B_group_stride;
    const T* C_base = (const T*)p.C + (size_t)t.b * p.C_batch_stride + (size_t)t.g * p.C_group_stride;
    float* dB_base = q.dB + ((size_t)t.b * p.n_groups + t.g) * (size_t)N * L;
    float* dC_base = q.dC + ((size_t)t.b * p.n_groups + t.g) * (size_t)N * L;
    const bool vec_u = can_vectorize<T>(p.u, p.u_row_stride, p.u_batch_stride, p.u_group_stride, L);
    const bool vec_d = can_vectorize<T>(p.delta, p.delta_row_stride, p.delta_batch_stride, 0, L);
    const bool vec_g = can_vectorize<T>(q.dout, q.dout_row_stride, q.dout_batch_stride, q.dout_group_stride, L);
    const bool vec_z = HAS_Z && can_vectorize<T>(p.z, p.z_row_stride, p.z_batch_stride, 0, L);
    const bool vec_out = !HAS_Z && can_vectorize<T>(q.du, q.du_row_stride, q.du_batch_stride, 0, L) &&
                         can_vectorize<T>(q.ddelta, q.ddelta_row_stride, q.ddelta_batch_stride, 0, L);

    const int nck = (L + TC - 1) / TC;
    const float* ck = p.ckpt + (size_t)blockIdx.x * (size_t)(nck - 1) * NS * 32;
    const int item = rs_item<TC>(lane);

    auto l_lo_of = [&](int c) { return REV ? L - (c + 1) * TC : c * TC; };
    auto prefetch = [&](int c) {  // chunk c (scan order); chunks are visited last -> first
        if (c >= 0) {
            const int buf = c & 1, l_lo = l_lo_of(c);
            stage_rows<T, TC, NTHR>(sm.u[buf], u_base, p.u_row_stride, t.nrows, l_lo, L, vec_u, tid);
            stage_rows<T, TC, NTHR>(sm.d[buf], d_base, p.delta_row_stride, t.nrows, l_lo, L, vec_d, tid);
            stage_rows<T, TC, NTHR>(sm.g[buf], g_base, q.dout_row_stride, t.nrows, l_lo, L, vec_g, tid);
            if (HAS_Z) stage_rows<T, TC, NTHR>(sm.z[buf], z_base, p.z_row_stride, t.nrows, l_lo, L, vec_z, tid);
            stage_bc<T, TC, NTHR>(sm.B[buf], B_base, p.B_state_stride, N, l_lo, L, tid);
            stage_bc<T, TC, NTHR>(sm.C[buf], C_base, p.C_state_stride, N, l_lo, L, tid);
            if (c > 0) {  // entry state of the chunk: 2 KB contiguous
                const float* src = ck + (size_t)(c - 1) * NS * 32;
#pragma unroll
                for (int idx0 = 0; idx0 < NS * 32 / 4; idx0 += NTHR) cp_async16(&sm.ck[buf][(idx0 + tid) * 4], src + (idx0 + tid) * 4, true);
            } else {
#pragma unroll
                for (int idx0 = 0; idx0 < NS * 32; idx0 += NTHR) sm.ck[buf][idx0 + tid] = 0.f;
            }
        }
        cp_async_commit();
    };
    prefetch(nck - 1);
    prefetch(nck - 2);

    for (int c = nck - 1; c >= 0; --c) {
        const int buf = c & 1;
        const int s0 = c * TC;
        const int nv = min(TC, L - s0);
        const int l_lo = l_lo_of(c);
        cp_async_wait<1>();
        __syncthreads();  // (1)
        // ---- pre-pass (once per element): delta', its sigmoid, the gated upstream gradient ----
#pragma unroll
        for (int idx0 = 0; idx0 < 32 * (TC / 4); idx0 += NTHR) {
            const int idx = idx0 + tid;
            if ((32 * (TC / 4)) % NTHR != 0 && idx >= 32 * (TC / 4)) break;  // warp-uniform (TC/4*32 is a multiple of 32)
            const int rr = idx / (TC / 4), cq = (idx % (TC / 4)) * 4;
            const int o = rr * TP + cq;
            const float4 dr = *reinterpret_cast<const float4*>(&sm.d[buf][o]);
            const float bias = sm.bias[rr];
            const float draw[4] = {dr.x, dr.y, dr.z, dr.w};
            float dl[4], sg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int l = l_lo + cq + k;
                float v = draw[k] + bias, sgv = 1.f;
                if (softplus) {
                    const SoftplusSig r = softplus_sigmoid(v);
                    v = r.sp;
                    sgv = r.sig;
                }
                if (l < 0 || l >= L) { v = 0.f; sgv = 0.f; }  // identity steps (u and dout are already zero-filled)
                dl[k] = v;
                sg[k] = sgv;
            }
            *reinterpret_cast<float4*>(&sm.d[buf][o]) = make_float4(dl[0], dl[1], dl[2], dl[3]);
            *reinterpret_cast<float4*>(&sm.sg[o]) = make_float4(sg[0], sg[1], sg[2], sg[3]);
            if (HAS_Z) {
                const float4 zr = *reinterpret_cast<const float4*>(&sm.z[buf][o]);
                const float4 gr = *reinterpret_cast<const float4*>(&sm.g[buf][o]);
                const float zz[4] = {zr.x, zr.y, zr.z, zr.w};
                const float go[4] = {gr.x, gr.y, gr.z, gr.w};
                float dzc[4], gg[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float s = sigmoidf_(zz[k]);
                    dzc[k] = go[k] * s * (1.f + zz[k] * (1.f - s));  // dout * d silu(z)/dz
                    gg[k] = go[k] * zz[k] * s;                       // dout * silu(z)
                }
                *reinterpret_cast<float4*>(&sm.z[buf][o]) = make_float4(dzc[0], dzc[1], dzc[2], dzc[3]);
                *reinterpret_cast<float4*>(&sm.g[buf][o]) = make_float4(gg[0], gg[1], gg[2], gg[3]);
            }
        }
        __syncthreads();  // (2)

        // ---- this lane's row of the chunk in registers, in scan order ----
        float dl[TC], duu[TC], go[TC];
        float2 s1[TC], s2[TC];
        float2 yacc[HAS_Z ? TC : 1];
#pragma unroll
        for (int i4 = 0; i4 < TC / 4; ++i4) {
            const int cb = REV ? TC - 4 - 4 * i4 : 4 * i4;
            const float4 u4 = ld4<REV>(&sm.u[buf][lane * TP + cb]);
            const float4 d4 = ld4<REV>(&sm.d[buf][lane * TP + cb]);
            const float4 g4 = ld4<REV>(&sm.g[buf][lane * TP + cb]);
            dl[i4 * 4] = d4.x; dl[i4 * 4 + 1] = d4.y; dl[i4 * 4 + 2] = d4.z; dl[i4 * 4 + 3] = d4.w;
            go[i4 * 4] = g4.x; go[i4 * 4 + 1] = g4.y; go[i4 * 4 + 2] = g4.z; go[i4 * 4 + 3] = g4.w;
            duu[i4 * 4] = d4.x * u4.x; duu[i4 * 4 + 1] = d4.y * u4.y; duu[i4 * 4 + 2] = d4.z * u4.z; duu[i4 * 4 + 3] = d4.w * u4.w;
        }
#pragma unroll
        for (int i = 0; i < TC; ++i) {
            s1[i] = make_float2(0.f, 0.f);
            s2[i] = make_float2(0.f, 0.f);
            if (HAS_Z) yacc[i] = make_float2(0.f, 0.f);
        }

        // ---- this warp's states, one packed pair at a time ----
#pragma unroll
        for (int pr = 0; pr < SPW / 2; ++pr) {
            const int n = w * SPW + 2 * pr;
            if (n < N) {  // warp-uniform
                float2 a[TC], ax[TC];
                float2 xp = make_float2(sm.ck[buf][n * 32 + lane], sm.ck[buf][(n + 1) * 32 + lane]);
                // forward recompute of the chunk from its checkpoint
#pragma unroll
                for (int i = 0; i < TC; ++i) {
                    const int cc = REV ? TC - 1 - i : i;
                    const float2 Bv = *reinterpret_cast<const float2*>(&sm.B[buf][cc * PB + n]);
                    const float2 e = __fmul2_rn(dup2(dl[i]), A2[pr]);
                    a[i] = make_float2(ex2(e.x), ex2(e.y));
                    ax[i] = __fmul2_rn(a[i], xp);
                    xp = __ffma2_rn(dup2(duu[i]), Bv, ax[i]);
                }
                // adjoint recurrence, last step first.  gn = a_{i+1} * (adjoint of x_{i+1})
                float2 gn = h[pr];
                float2 dAp = make_float2(0.f, 0.f);
                float vB0[TC], vB1[TC], vC0[TC], vC1[TC];
#pragma unroll
                for (int i = TC - 1; i >= 0; --i) {
                    const int cc = REV ? TC - 1 - i : i;
                    const float2 Bv = *reinterpret_cast<const float2*>(&sm.B[buf][cc * PB + n]);
                    const float2 Cv = *reinterpret_cast<const float2*>(&sm.C[buf][cc * PB + n]);
                    const float2 go2 = dup2(go[i]), du2 = dup2(duu[i]);
                    const float2 g = __ffma2_rn(go2, Cv, gn);
                    const float2 xv = __ffma2_rn(du2, Bv, ax[i]);
                    const float2 vC = __fmul2_rn(go2, xv);
                    const float2 vB = __fmul2_rn(g, du2);
                    vC0[i] = vC.x; vC1[i] = vC.y; vB0[i] = vB.x; vB1[i] = vB.y;
                    s1[i] = __ffma2_rn(g, Bv, s1[i]);
                    const float2 wv = __fmul2_rn(g, ax[i]);
                    s2[i] = __ffma2_rn(An[pr], wv, s2[i]);
                    dAp = __ffma2_rn(wv, dup2(dl[i]), dAp);
                    if (HAS_Z) yacc[i] = __ffma2_rn(Cv, xv, yacc[i]);
                    gn = __fmul2_rn(a[i], g);
                }
                h[pr] = gn;
                dA[pr] = __fadd2_rn(dA[pr], dAp);
                // dB/dC: sum over the 32 rows of this warp, then one atomic per (state, step)
                const float rB0 = reduce_scatter<TC>(vB0, lane);
                const float rC0 = reduce_scatter<TC>(vC0, lane);
                const float rB1 = reduce_scatter<TC>(vB1, lane);
                const float rC1 = reduce_scatter<TC>(vC1, lane);
                if ((lane & 3) == 0 && item < nv) {
                    const int sl = s0 + item;
                    const int ll = REV ? (L - 1 - sl) : sl;
                    atomicAdd(dB_base + (size_t)n * L + ll, rB0);
                    atomicAdd(dC_base + (size_t)n * L + ll, rC0);
                    if (n + 1 < N) {
                        atomicAdd(dB_base + (size_t)(n + 1) * L + ll, rB1);
                        atomicAdd(dC_base + (size_t)(n + 1) * L + ll, rC1);
                    }
                }
            }
        }
        // this warp's partial sums over its states, stored in memory-order columns
#pragma unroll
        for (int i4 = 0; i4 < TC / 4; ++i4) {
            const int cb = REV ? TC - 4 - 4 * i4 : 4 * i4;
            const int i = 4 * i4;
            st4<REV>(&sm.part[w][lane * SP + cb], s1[i].x + s1[i].y, s1[i + 1].x + s1[i + 1].y, s1[i + 2].x + s1[i + 2].y,
                     s1[i + 3].x + s1[i + 3].y);
            st4<REV>(&sm.part[w][lane * SP + TC + cb], s2[i].x + s2[i].y, s2[i + 1].x + s2[i + 1].y, s2[i + 2].x + s2[i + 2].y,
                     s2[i + 3].x + s2[i + 3].y);
            if (HAS_Z)
                st4<REV>(&sm.part[w][lane * SP + 2 * TC + cb], yacc[i].x + yacc[i].y, yacc[i + 1].x + yacc[i + 1].y,
                         yacc[i + 2].x + yacc[i + 2].y, yacc[i + 3].x + yacc[i + 3].y);
        }
        __syncthreads();  // (3)

        // ---- per-element gradients, four steps per thread ----
#pragma unroll
        for (int k = 0; k < EP; ++k) {
            const int idx = tid + k * NTHR;
            if (idx < 32 * (TC / 4)) {
                const int rr = idx / (TC / 4), cq = (idx % (TC / 4)) * 4;
                float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, vy = v1;
#pragma unroll
                for (int ww = 0; ww < NW; ++ww) {
                    const float4 a1 = *reinterpret_cast<const float4*>(&sm.part[ww][rr * SP + cq]);
                    const float4 a2 = *reinterpret_cast<const float4*>(&sm.part[ww][rr * SP + TC + cq]);
                    v1.x += a1.x; v1.y += a1.y; v1.z += a1.z; v1.w += a1.w;
                    v2.x += a2.x; v2.y += a2.y; v2.z += a2.z; v2.w += a2.w;
                    if (HAS_Z) {
                        const float4 a3 = *reinterpret_cast<const float4*>(&sm.part[ww][rr * SP + 2 * TC + cq]);
                        vy.x += a3.x; vy.y += a3.y; vy.z += a3.z; vy.w += a3.w;
                    }
                }
                const float4 d4 = *reinterpret_cast<const float4*>(&sm.d[buf][rr * TP + cq]);
                const float4 u4 = *reinterpret_cast<const float4*>(&sm.u[buf][rr * TP + cq]);
                const float4 g4 = *reinterpret_cast<const float4*>(&sm.g[buf][rr * TP + cq]);
                const float4 s4 = *reinterpret_cast<const float4*>(&sm.sg[rr * TP + cq]);
                const float Dv = sm.D[rr];
                const float s1v[4] = {v1.x, v1.y, v1.z, v1.w}, s2v[4] = {v2.x, v2.y, v2.z, v2.w};
                const float dlv[4] = {d4.x, d4.y, d4.z, d4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
                const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, sgv[4] = {s4.x, s4.y, s4.z, s4.w};
                float du[4], dd[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    du[e] = fmaf(dlv[e], s1v[e], Dv * gv[e]);
                    // chain rule through softplus: sg = sigmoid(delta + bias) (1 without softplus, 0 off-range)
                    dd[e] = fmaf(uv[e], s1v[e], s2v[e]) * sgv[e];
                    dbias_acc[k] += dd[e];                  // off-range / padded rows contribute exact zeros
                    dD_acc[k] = fmaf(gv[e], uv[e], dD_acc[k]);
                }
                if (rr < t.nrows) {
                    const int l = l_lo + cq;
                    if (vec_out) {
                        if (l >= 0 && l < L) {
                            __stcs(reinterpret_cast<float4*>((float*)du_base + (size_t)rr * q.du_row_stride + l),
                                   make_float4(du[0], du[1], du[2], du[3]));
                            __stcs(reinterpret_cast<float4*>((float*)dd_base + (size_t)rr * q.ddelta_row_stride + l),
                                   make_float4(dd[0], dd[1], dd[2], dd[3]));
                        }
                    } else {
                        const float4 z4 = HAS_Z ? *reinterpret_cast<const float4*>(&sm.z[buf][rr * TP + cq]) : make_float4(0, 0, 0, 0);
                        const float dzc[4] = {z4.x, z4.y, z4.z, z4.w}, yv[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (l + e >= 0 && l + e < L) {
                                stg_stream(du_base + (size_t)rr * q.du_row_stride + l + e, du[e]);
                                stg_stream(dd_base + (size_t)rr * q.ddelta_row_stride + l + e, dd[e]);
                                if (HAS_Z) stg_stream(dz_base + (size_t)rr * q.dz_row_stride + l + e, dzc[e] * fmaf(Dv, uv[e], yv[e]));
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // (4) every thread is done with the raw tiles of this chunk
        prefetch(c - 2);
    }
    cp_async_wait<0>();

    if (row_ok) {
        const float dAs[4] = {dA[0].x, dA[0].y, dA[1].x, dA[1].y};
#pragma unroll
        for (int j = 0; j < SPW; ++j)
            if (w * SPW + j < N) atomicAdd(q.dA + (size_t)d_lane * N + w * SPW + j, dAs[j]);
    }
    // per-row sums: TC/4 consecutive threads share a row in every epilogue pass
#pragma unroll
    for (int k = 0; k < EP; ++k) {
        const int idx = tid + k * NTHR;
        float sb = dbias_acc[k], sd = dD_acc[k];
#pragma unroll
        for (int m = TC / 8; m >= 1; m >>= 1) {
            sb += __shfl_xor_sync(0xffffffffu, sb, m);
            sd += __shfl_xor_sync(0xffffffffu, sd, m);
        }
        if (idx < 32 * (TC / 4) && (idx % (TC / 4)) == 0) {
            const int rr = idx / (TC / 4);
            if (rr < t.nrows) {
                if (q.ddelta_bias) atomicAdd(q.ddelta_bias + t.d0 + rr, sb);
                if (q.dD) atomicAdd(q.dD + t.d0 + rr, sd);
            }
        }
    }
}

template <typename T, int TC, int NW, bool HAS_Z>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 3 : 4) sscan_bwd_kernel(const __grid_constant__ b200_sscan_bwd_params q) {
    __shared__ __align__(16) BwdSmem<TC, HAS_Z> sm;
    const Task t = decode_task(q.f, blockIdx.x);
    if (t.rev) sscan_bwd_body<T, TC, NW, HAS_Z, true>(q, t, sm);
    else sscan_bwd_body<T, TC, NW, HAS_Z, false>(q, t, sm);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int validate(const b200_sscan_fwd_params* p) {
    B200_REQUIRE(p != nullptr, "b200_sscan: params is NULL");
    B200_REQUIRE(p->batch > 0 && p->dim > 0 && p->seqlen > 0, "b200_sscan: batch/dim/seqlen must be positive (got %d/%d/%d)",
                 p->batch, p->dim, p->seqlen);
    B200_REQUIRE(p->dstate >= 1 && p->dstate <= B200_SSCAN_MAX_DSTATE, "b200_sscan: dstate %d outside [1, %d]", p->dstate,
                 B200_SSCAN_MAX_DSTATE);
    B200_REQUIRE(p->dstate <= NS, "b200_sscan: dstate %d > %d is not implemented in this build", p->dstate, NS);
    B200_REQUIRE(p->n_groups >= 1 && p->dim % p->n_groups == 0, "b200_sscan: dim %d is not divisible by n_groups %d", p->dim,
                 p->n_groups);
    B200_REQUIRE(p->io_dtype >= B200_F32 && p->io_dtype <= B200_F16, "b200_sscan: bad io_dtype %d", p->io_dtype);
    B200_REQUIRE(p->rev_mask == 0 || p->n_groups <= 32, "b200_sscan: rev_mask needs n_groups <= 32");
    B200_REQUIRE(p->u_group_div >= 1, "b200_sscan: u_group_div must be >= 1");
    B200_REQUIRE(p->u && p->delta && p->A && p->B && p->C, "b200_sscan: u/delta/A/B/C must be non-NULL");
    B200_REQUIRE(p->ckpt == nullptr || p->ckpt_every == 8, "b200_sscan: ckpt_every must be 8");
    return 0;
}

static long long n_tasks(const b200_sscan_fwd_params* p) {
    const int rpg = p->dim / p->n_groups;
    return (long long)p->batch * p->n_groups * ((rpg + 31) / 32);
}

template <typename T, bool HAS_Z>
static int launch_fwd_z(const b200_sscan_fwd_params* p, cudaStream_t st) {
    const unsigned grid = (unsigned)n_tasks(p);
    switch ((p->dstate + SPW - 1) / SPW) {
        case 1: sscan_fwd_kernel<T, 16, 1, HAS_Z><<<grid, 32, 0, st>>>(*p); break;
        case 2: sscan_fwd_kernel<T, 16, 2, HAS_Z><<<grid, 64, 0, st>>>(*p); break;
        case 3: sscan_fwd_kernel<T, 16, 3, HAS_Z><<<grid, 96, 0, st>>>(*p); break;
        default: sscan_fwd_kernel<T, 16, 4, HAS_Z><<<grid, 128, 0, st>>>(*p); break;
    }
    return check_launch("sscan_fwd_kernel");
}
template <typename T>
static int launch_fwd(const b200_sscan_fwd_params* p, cudaStream_t st) {
    return p->z ? launch_fwd_z<T, true>(p, st) : launch_fwd_z<T, false>(p, st);
}

template <typename T, bool HAS_Z>
static int launch_bwd_z(const b200_sscan_bwd_params* q, cudaStream_t st) {
    const unsigned grid = (unsigned)n_tasks(&q->f);
    switch ((q->f.dstate + SPW - 1) / SPW) {
        case 1: sscan_bwd_kernel<T, 8, 1, HAS_Z><<<grid, 32, 0, st>>>(*q); break;
        case 2: sscan_bwd_kernel<T, 8, 2, HAS_Z><<<grid, 64, 0, st>>>(*q); break;
        case 3: sscan_bwd_kernel<T, 8, 3, HAS_Z><<<grid, 96, 0, st>>>(*q); break;
        default: sscan_bwd_kernel<T, 8, 4, HAS_Z><<<grid, 128, 0, st>>>(*q); break;
    }
    return check_launch("sscan_bwd_kernel");
}
template <typename T>
static int launch_bwd(const b200_sscan_bwd_params* q, cudaStream_t st) {
    return q->f.z ? launch_bwd_z<T, true>(q, st) : launch_bwd_z<T, false>(q, st);
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_sscan_ckpt_bytes(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate, int32_t n_groups,
                                        int32_t ckpt_every) {
    (void)dstate;
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || n_groups <= 0 || dim % n_groups || ckpt_every <= 0) return 0;
    const int rpg = dim / n_groups;
    const size_t tasks = (size_t)batch * n_groups * ((rpg + 31) / 32);
    const size_t nck = (seqlen + ckpt_every - 1) / ckpt_every;
    const size_t bytes = tasks * (nck - 1) * NS * 32 * sizeof(float);
    return bytes ? bytes : 16;
}

extern "C" int b200_sscan_fwd(const b200_sscan_fwd_params* p, b200_stream_t stream) {
    if (int rc = validate(p)) return rc;
    B200_REQUIRE(p->out != nullptr, "b200_sscan_fwd: out is NULL");
    B200_REQUIRE(n_tasks(p) < (1ll << 31), "b200_sscan_fwd: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->io_dtype) {
        case B200_F32: return launch_fwd<float>(p, st);
        case B200_BF16: return launch_fwd<__nv_bfloat16>(p, st);
        default: return launch_fwd<__half>(p, st);
    }
}

extern "C" int b200_sscan_bwd(const b200_sscan_bwd_params* q, b200_stream_t stream) {
    B200_REQUIRE(q != nullptr, "b200_sscan_bwd: params is NULL");
    if (int rc = validate(&q->f)) return rc;
    B200_REQUIRE(q->f.ckpt != nullptr, "b200_sscan_bwd: the forward checkpoints (f.ckpt) are required");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(q->f.ckpt) & 15) == 0, "b200_sscan_bwd: ckpt must be 16-byte aligned");
    B200_REQUIRE(q->dout && q->du && q->ddelta && q->dA && q->dB && q->dC, "b200_sscan_bwd: dout/du/ddelta/dA/dB/dC must be non-NULL");
    B200_REQUIRE((q->f.z == nullptr) == (q->dz == nullptr), "b200_sscan_bwd: dz must be given exactly when z is");
    B200_REQUIRE(q->dout_group_div >= 1, "b200_sscan_bwd: dout_group_div must be >= 1");
    B200_REQUIRE(n_tasks(&q->f) < (1ll << 31), "b200_sscan_bwd: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    switch (q->f.io_dtype) {
        case B200_F32: return launch_bwd<float>(q, st);
        case B200_BF16: return launch_bwd<__nv_bfloat16>(q, st);
        default: return launch_bwd<__half>(q, st);
    }
}
