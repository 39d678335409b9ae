// sscan.cu -- Mamba-1 selective scan, forward and backward, for sm_100a.
//
// Replaces selective_scan_cuda.fwd/.bwd (reference CrossMamba/FusionMamba/selective_scan/
// selective_scan_fwd_kernel.cuh:67-303, selective_scan_bwd_kernel.cuh:75-489).  Not a port: the
// reference maps one CTA to one (batch, channel) row and runs a CUB block scan per state; here
//
//   * one LANE owns one (batch, channel) row and keeps all 16 states of the recurrence in
//     registers, so the time recurrence is a plain in-register FFMA chain: no block scan, no
//     shuffles, every exp(delta*A) is evaluated exactly once;
//   * a warp task = 32 consecutive channels of one (batch, group): the group's B/C tile is staged
//     in shared memory once per 32 rows and read back as warp-wide broadcasts (the reference
//     re-reads B/C from L2 for every row);
//   * u/delta tiles are read with coalesced row segments and turned through shared memory
//     (row pitch TT+4 floats => conflict-free 128-bit reads by the owning lane);
//   * a group can be scanned in reverse time (rev_mask) and groups can share u rows (u_group_div),
//     which is all the SS2D cross-scan needs (MedMamba.py:393-395);
//   * backward = recompute: the forward stores the 16-float state every `ckpt_every` steps, the
//     backward walks chunks last->first, re-derives the forward states of one chunk in registers,
//     runs the adjoint recurrence, and reduces dB/dC over the 32 rows of the warp with a
//     reduce-scatter butterfly before issuing one fp32 atomic per (state, step).
//
// Binding pipes: forward = MUFU.EX2 (16 per element), backward = FP32 issue.
#include "common.cuh"

namespace b200 {

constexpr int NS = 16;  // states per lane held in registers
constexpr int PB = 20;  // pitch of the transposed B/C tiles [t][n] (16 + 4: 16B-aligned rows, spreads banks)

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

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
template <int TT>
struct FwdSmem {
    static constexpr int TP = TT + 4;
    float u[32 * TP];   // u tile, overwritten in place by the output tile
    float d[32 * TP];   // raw delta tile
    float B[TT * PB];   // [t][n]
    float C[TT * PB];
};

template <typename T, int TT, bool HAS_Z>
__global__ void __launch_bounds__(32) sscan_fwd_kernel(const __grid_constant__ b200_sscan_fwd_params p) {
    __shared__ __align__(16) FwdSmem<TT> sm;
    constexpr int TP = FwdSmem<TT>::TP;
    constexpr int RPI = 32 / TT;  // rows covered by one warp-wide load instruction
    const int lane = threadIdx.x;
    const Task t = decode_task(p, blockIdx.x);
    const int L = p.seqlen, N = p.dstate;
    const bool row_ok = lane < t.nrows;
    const int d_lane = t.d0 + lane;

    float A2[NS], x[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
        A2[n] = (row_ok && n < N) ? __ldg(p.A + (size_t)d_lane * N + n) * kLog2e : 0.f;
        x[n] = 0.f;
    }
    const float Dv = (p.D && row_ok) ? __ldg(p.D + d_lane) : 0.f;
    const float bias = (p.delta_bias && row_ok) ? __ldg(p.delta_bias + d_lane) : 0.f;
    const bool softplus = p.delta_softplus != 0;

    const T* u_base = (const T*)p.u + (size_t)t.b * p.u_batch_stride + (size_t)(t.g / p.u_group_div) * p.u_group_stride +
                      (size_t)t.r0 * p.u_row_stride;
    const T* d_base = (const T*)p.delta + (size_t)t.b * p.delta_batch_stride + (size_t)t.d0 * p.delta_row_stride;
    T* o_base = (T*)p.out + (size_t)t.b * p.out_batch_stride + (size_t)t.d0 * p.out_row_stride;
    const T* z_base = HAS_Z ? (const T*)p.z + (size_t)t.b * p.z_batch_stride + (size_t)t.d0 * p.z_row_stride : nullptr;
    const T* B_base = (const T*)p.B + (size_t)t.b * p.B_batch_stride + (size_t)t.g * p.B_group_stride;
    const T* C_base = (const T*)p.C + (size_t)t.b * p.C_batch_stride + (size_t)t.g * p.C_group_stride;

    const int ce = p.ckpt_every;
    const int nck = p.ckpt ? (L + ce - 1) / ce : 0;
    float* ck = p.ckpt ? p.ckpt + (size_t)blockIdx.x * (size_t)(nck - 1) * NS * 32 : nullptr;

    const int li = lane % TT, lr = lane / TT;

    for (int s0 = 0; s0 < L; s0 += TT) {
        const int nv = min(TT, L - s0);
        const int s = s0 + li;
        const bool tok = s < L;
        const int l = t.rev ? (L - 1 - s) : s;  // memory position of scan step s
        // ---- stage tiles in scan order ----
#pragma unroll 4
        for (int rr = lr; rr < 32; rr += RPI) {
            float uv = 0.f, dv = 0.f;
            if (tok && rr < t.nrows) {
                uv = ldg_stream(u_base + (size_t)rr * p.u_row_stride + l);
                dv = ldg_stream(d_base + (size_t)rr * p.delta_row_stride + l);
            }
            sm.u[rr * TP + li] = uv;
            sm.d[rr * TP + li] = dv;
        }
#pragma unroll 4
        for (int n = lr; n < NS; n += RPI) {
            float bv = 0.f, cv = 0.f;
            if (tok && n < N) {
                bv = to_f32<T>(__ldg(B_base + (size_t)n * p.B_state_stride + l));
                cv = to_f32<T>(__ldg(C_base + (size_t)n * p.C_state_stride + l));
            }
            sm.B[li * PB + n] = bv;
            sm.C[li * PB + n] = cv;
        }
        __syncwarp();
        // ---- recurrence: this lane's row, 4 steps per iteration ----
#pragma unroll 1
        for (int i4 = 0; i4 < TT / 4; ++i4) {
            if (i4 * 4 >= nv) break;
            const int sb = s0 + i4 * 4;
            if (ck != nullptr && sb > 0 && (sb % ce) == 0) {
                float* dst = ck + ((size_t)(sb / ce - 1) * NS) * 32 + lane;
#pragma unroll
                for (int n = 0; n < NS; ++n) dst[n * 32] = x[n];
            }
            const float4 u4 = *reinterpret_cast<const float4*>(&sm.u[lane * TP + i4 * 4]);
            const float4 d4 = *reinterpret_cast<const float4*>(&sm.d[lane * TP + i4 * 4]);
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            float yy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i4 * 4 + k;
                float dl = dd[k] + bias;
                if (softplus) dl = softplus20(dl);
                float uk = uu[k];
                if (i >= nv) { dl = 0.f; uk = 0.f; }  // identity step: a = 1, b = 0
                const float du = dl * uk;
                float y = Dv * uk;
                float Bv[NS], Cv[NS];
#pragma unroll
                for (int q = 0; q < NS / 4; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&sm.B[i * PB + 4 * q]);
                    const float4 c4 = *reinterpret_cast<const float4*>(&sm.C[i * PB + 4 * q]);
                    Bv[4 * q] = b4.x; Bv[4 * q + 1] = b4.y; Bv[4 * q + 2] = b4.z; Bv[4 * q + 3] = b4.w;
                    Cv[4 * q] = c4.x; Cv[4 * q + 1] = c4.y; Cv[4 * q + 2] = c4.z; Cv[4 * q + 3] = c4.w;
                }
#pragma unroll
                for (int n = 0; n < NS; ++n) {
                    const float a = ex2(dl * A2[n]);
                    x[n] = fmaf(a, x[n], du * Bv[n]);
                    y = fmaf(Cv[n], x[n], y);
                }
                yy[k] = y;
            }
            *reinterpret_cast<float4*>(&sm.u[lane * TP + i4 * 4]) = make_float4(yy[0], yy[1], yy[2], yy[3]);
        }
        __syncwarp();
        // ---- store the output tile at its memory position ----
#pragma unroll 4
        for (int rr = lr; rr < 32; rr += RPI) {
            if (tok && rr < t.nrows) {
                float v = sm.u[rr * TP + li];
                if (HAS_Z) {
                    const float zz = ldg_stream(z_base + (size_t)rr * p.z_row_stride + l);
                    v *= zz * sigmoidf_(zz);
                }
                stg_stream(o_base + (size_t)rr * p.out_row_stride + l, v);
            }
        }
        __syncwarp();
    }
    if (p.last_state != nullptr && row_ok) {
        float* ls = p.last_state + ((size_t)t.b * p.dim + d_lane) * N;
#pragma unroll
        for (int n = 0; n < NS; ++n)
            if (n < N) ls[n] = x[n];
    }
}

// ----------------------------------------------------------------------------------------------
// backward
// ----------------------------------------------------------------------------------------------
// Sum v[0..TC) over the 32 lanes of the warp with a reduce-scatter butterfly: each halving step
// exchanges half of the remaining values, so the whole reduction costs ~TC shuffles instead of
// 5*TC.  On return every lane holds the 32-row total of item `rs_item<TC>(lane)`.
template <int TC> __device__ __forceinline__ int rs_item(int lane);
template <> __device__ __forceinline__ int rs_item<8>(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
template <> __device__ __forceinline__ int rs_item<16>(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

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
    float r = rs_step<TC>(v, lane, 16);
    if constexpr (TC == 8) {
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
    } else {
        r += __shfl_xor_sync(0xffffffffu, r, 1);
    }
    return r;
}

template <int TC, bool HAS_Z>
struct BwdSmem {
    static constexpr int TP = TC + 4;
    float u[32 * TP];    // u tile  -> du tile
    float d[32 * TP];    // raw delta tile -> ddelta tile
    float g[32 * TP];    // dout tile -> dz tile
    float z[HAS_Z ? 32 * TP : 4];
    float B[TC * PB];
    float C[TC * PB];
    float ck[NS * 32];   // state entering the chunk, [n][lane]
    float h[NS * 32];    // adjoint carried from the later chunk, [n][lane]
    float dA[NS * 32];   // per-row dA accumulators
    float A[NS * 32];
};

template <typename T, int TC, bool HAS_Z>
__global__ void __launch_bounds__(32) sscan_bwd_kernel(const __grid_constant__ b200_sscan_bwd_params q) {
    const b200_sscan_fwd_params& p = q.f;
    __shared__ __align__(16) BwdSmem<TC, HAS_Z> sm;
    constexpr int TP = BwdSmem<TC, HAS_Z>::TP;
    constexpr int RPI = 32 / TC;
    const int lane = threadIdx.x;
    const Task t = decode_task(p, blockIdx.x);
    const int L = p.seqlen, N = p.dstate;
    const int Npad = (N + 1) & ~1;
    const bool row_ok = lane < t.nrows;
    const int d_lane = t.d0 + lane;

#pragma unroll
    for (int n = 0; n < NS; ++n) {
        sm.h[n * 32 + lane] = 0.f;
        sm.dA[n * 32 + lane] = 0.f;
        sm.A[n * 32 + lane] = (row_ok && n < N) ? __ldg(p.A + (size_t)d_lane * N + n) : 0.f;
    }
    const float Dv = (p.D && row_ok) ? __ldg(p.D + d_lane) : 0.f;
    const float bias = (p.delta_bias && row_ok) ? __ldg(p.delta_bias + d_lane) : 0.f;
    const bool softplus = p.delta_softplus != 0;
    float dD_acc = 0.f, dbias_acc = 0.f;

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

    const int nck = (L + TC - 1) / TC;
    const float* ck = p.ckpt + (size_t)blockIdx.x * (size_t)(nck - 1) * NS * 32;
    const int li = lane % TC, lr = lane / TC;
    const int item = rs_item<TC>(lane);

    for (int c = nck - 1; c >= 0; --c) {
        const int s0 = c * TC;
        const int nv = min(TC, L - s0);
        const int s = s0 + li;
        const bool tok = s < L;
        const int l = t.rev ? (L - 1 - s) : s;
        // ---- stage tiles (scan order) + the chunk's entry state ----
#pragma unroll 4
        for (int rr = lr; rr < 32; rr += RPI) {
            float uv = 0.f, dv = 0.f, gv = 0.f, zv = 0.f;
            if (tok && rr < t.nrows) {
                uv = ldg_stream(u_base + (size_t)rr * p.u_row_stride + l);
                dv = ldg_stream(d_base + (size_t)rr * p.delta_row_stride + l);
                gv = ldg_stream(g_base + (size_t)rr * q.dout_row_stride + l);
                if (HAS_Z) zv = ldg_stream(z_base + (size_t)rr * p.z_row_stride + l);
            }
            sm.u[rr * TP + li] = uv;
            sm.d[rr * TP + li] = dv;
            sm.g[rr * TP + li] = gv;
            if (HAS_Z) sm.z[rr * TP + li] = zv;
        }
#pragma unroll 4
        for (int n = lr; n < NS; n += RPI) {
            float bv = 0.f, cv = 0.f;
            if (tok && n < N) {
                bv = to_f32<T>(__ldg(B_base + (size_t)n * p.B_state_stride + l));
                cv = to_f32<T>(__ldg(C_base + (size_t)n * p.C_state_stride + l));
            }
            sm.B[li * PB + n] = bv;
            sm.C[li * PB + n] = cv;
        }
        if (c > 0) {
            const float* src = ck + ((size_t)(c - 1) * NS) * 32 + lane;
#pragma unroll
            for (int n = 0; n < NS; ++n) sm.ck[n * 32 + lane] = __ldcs(src + n * 32);
        } else {
#pragma unroll
            for (int n = 0; n < NS; ++n) sm.ck[n * 32 + lane] = 0.f;
        }
        __syncwarp();

        // ---- this lane's row of the chunk, in registers ----
        float dl[TC], uu[TC], go[TC], s1[TC], s2[TC];
        float dzc[HAS_Z ? TC : 1], yacc[HAS_Z ? TC : 1];
#pragma unroll
        for (int i4 = 0; i4 < TC / 4; ++i4) {
            const float4 u4 = *reinterpret_cast<const float4*>(&sm.u[lane * TP + i4 * 4]);
            const float4 d4 = *reinterpret_cast<const float4*>(&sm.d[lane * TP + i4 * 4]);
            const float4 g4 = *reinterpret_cast<const float4*>(&sm.g[lane * TP + i4 * 4]);
            uu[i4 * 4] = u4.x; uu[i4 * 4 + 1] = u4.y; uu[i4 * 4 + 2] = u4.z; uu[i4 * 4 + 3] = u4.w;
            dl[i4 * 4] = d4.x; dl[i4 * 4 + 1] = d4.y; dl[i4 * 4 + 2] = d4.z; dl[i4 * 4 + 3] = d4.w;
            go[i4 * 4] = g4.x; go[i4 * 4 + 1] = g4.y; go[i4 * 4 + 2] = g4.z; go[i4 * 4 + 3] = g4.w;
        }
#pragma unroll
        for (int i = 0; i < TC; ++i) {
            float v = dl[i] + bias;
            if (softplus) v = softplus20(v);
            if (i >= nv) { v = 0.f; uu[i] = 0.f; go[i] = 0.f; }
            dl[i] = v;
            s1[i] = 0.f;
            s2[i] = 0.f;
            if (HAS_Z) {
                const float zz = sm.z[lane * TP + i];
                const float sg = sigmoidf_(zz);
                dzc[i] = go[i] * sg * (1.f + zz * (1.f - sg));  // dout * d silu(z)/dz
                go[i] *= zz * sg;                               // dout * silu(z)
                yacc[i] = 0.f;
            }
        }

        // ---- states, two at a time ----
#pragma unroll 1
        for (int n = 0; n < Npad; n += 2) {
            const float An0 = sm.A[n * 32 + lane], An1 = sm.A[(n + 1) * 32 + lane];
            const float A20 = An0 * kLog2e, A21 = An1 * kLog2e;
            float a0[TC], a1[TC], ax0[TC], ax1[TC];
            float xp0 = sm.ck[n * 32 + lane], xp1 = sm.ck[(n + 1) * 32 + lane];
            // forward recompute of the chunk from its checkpoint
#pragma unroll
            for (int i = 0; i < TC; ++i) {
                const float2 Bv = *reinterpret_cast<const float2*>(&sm.B[i * PB + n]);
                const float duu = dl[i] * uu[i];
                a0[i] = ex2(dl[i] * A20);
                a1[i] = ex2(dl[i] * A21);
                ax0[i] = a0[i] * xp0;
                ax1[i] = a1[i] * xp1;
                xp0 = fmaf(duu, Bv.x, ax0[i]);
                xp1 = fmaf(duu, Bv.y, ax1[i]);
            }
            // adjoint recurrence, last step first.  gn = a_{i+1} * (adjoint of x_{i+1})
            float gn0 = sm.h[n * 32 + lane], gn1 = sm.h[(n + 1) * 32 + lane];
            float dA0 = 0.f, dA1 = 0.f;
            float vB0[TC], vB1[TC], vC0[TC], vC1[TC];
#pragma unroll
            for (int i = TC - 1; i >= 0; --i) {
                const float2 Bv = *reinterpret_cast<const float2*>(&sm.B[i * PB + n]);
                const float2 Cv = *reinterpret_cast<const float2*>(&sm.C[i * PB + n]);
                const float duu = dl[i] * uu[i];
                const float g0 = fmaf(go[i], Cv.x, gn0);
                const float g1 = fmaf(go[i], Cv.y, gn1);
                const float x0 = fmaf(duu, Bv.x, ax0[i]);
                const float x1 = fmaf(duu, Bv.y, ax1[i]);
                vC0[i] = go[i] * x0;
                vC1[i] = go[i] * x1;
                vB0[i] = g0 * duu;
                vB1[i] = g1 * duu;
                s1[i] = fmaf(g0, Bv.x, s1[i]);
                s1[i] = fmaf(g1, Bv.y, s1[i]);
                const float w0 = g0 * ax0[i];
                const float w1 = g1 * ax1[i];
                s2[i] = fmaf(An0, w0, s2[i]);
                s2[i] = fmaf(An1, w1, s2[i]);
                dA0 = fmaf(w0, dl[i], dA0);
                dA1 = fmaf(w1, dl[i], dA1);
                if (HAS_Z) {
                    yacc[i] = fmaf(Cv.x, x0, yacc[i]);
                    yacc[i] = fmaf(Cv.y, x1, yacc[i]);
                }
                gn0 = a0[i] * g0;
                gn1 = a1[i] * g1;
            }
            sm.h[n * 32 + lane] = gn0;
            sm.h[(n + 1) * 32 + lane] = gn1;
            sm.dA[n * 32 + lane] += dA0;
            sm.dA[(n + 1) * 32 + lane] += dA1;
            // dB/dC: sum over the 32 rows of this warp, then one atomic per (state, step)
            const float rB0 = reduce_scatter<TC>(vB0, lane);
            const float rC0 = reduce_scatter<TC>(vC0, lane);
            const float rB1 = reduce_scatter<TC>(vB1, lane);
            const float rC1 = reduce_scatter<TC>(vC1, lane);
            constexpr int REP = (TC == 8) ? 3 : 1;  // lanes holding the same total
            if ((lane & REP) == 0 && item < nv) {
                const int sl = s0 + item;
                const int ll = t.rev ? (L - 1 - sl) : sl;
                atomicAdd(dB_base + (size_t)n * L + ll, rB0);
                atomicAdd(dC_base + (size_t)n * L + ll, rC0);
                if (n + 1 < N) {
                    atomicAdd(dB_base + (size_t)(n + 1) * L + ll, rB1);
                    atomicAdd(dC_base + (size_t)(n + 1) * L + ll, rC1);
                }
            }
        }

        // ---- per-step gradients of this row ----
#pragma unroll
        for (int i = 0; i < TC; ++i) {
            const float du = fmaf(dl[i], s1[i], Dv * go[i]);
            float dd = fmaf(uu[i], s1[i], s2[i]);
            // d softplus(v)/dv = sigmoid(v) = 1 - exp(-softplus(v)); == 1 in fp32 beyond the threshold
            if (softplus) dd *= -expm1f(-dl[i]);
            if (i >= nv) dd = 0.f;
            dbias_acc += dd;
            dD_acc = fmaf(go[i], uu[i], dD_acc);
            sm.u[lane * TP + i] = du;
            sm.d[lane * TP + i] = dd;
            if (HAS_Z) sm.g[lane * TP + i] = dzc[i] * fmaf(Dv, uu[i], yacc[i]);
        }
        __syncwarp();
#pragma unroll 4
        for (int rr = lr; rr < 32; rr += RPI) {
            if (tok && rr < t.nrows) {
                stg_stream(du_base + (size_t)rr * q.du_row_stride + l, sm.u[rr * TP + li]);
                stg_stream(dd_base + (size_t)rr * q.ddelta_row_stride + l, sm.d[rr * TP + li]);
                if (HAS_Z) stg_stream(dz_base + (size_t)rr * q.dz_row_stride + l, sm.g[rr * TP + li]);
            }
        }
        __syncwarp();
    }

    if (row_ok) {
#pragma unroll
        for (int n = 0; n < NS; ++n)
            if (n < N) atomicAdd(q.dA + (size_t)d_lane * N + n, sm.dA[n * 32 + lane]);
        if (q.dD) atomicAdd(q.dD + d_lane, dD_acc);
        if (q.ddelta_bias) atomicAdd(q.ddelta_bias + d_lane, dbias_acc);
    }
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
    B200_REQUIRE(p->ckpt == nullptr || p->ckpt_every == 8 || p->ckpt_every == 16, "b200_sscan: ckpt_every must be 8 or 16");
    return 0;
}

static long long n_tasks(const b200_sscan_fwd_params* p) {
    const int rpg = p->dim / p->n_groups;
    return (long long)p->batch * p->n_groups * ((rpg + 31) / 32);
}

template <typename T>
static int launch_fwd(const b200_sscan_fwd_params* p, cudaStream_t st) {
    const unsigned grid = (unsigned)n_tasks(p);
    if (p->z)
        sscan_fwd_kernel<T, 16, true><<<grid, 32, 0, st>>>(*p);
    else
        sscan_fwd_kernel<T, 16, false><<<grid, 32, 0, st>>>(*p);
    return check_launch("sscan_fwd_kernel");
}

template <typename T>
static int launch_bwd(const b200_sscan_bwd_params* q, cudaStream_t st) {
    const unsigned grid = (unsigned)n_tasks(&q->f);
    const bool z = q->f.z != nullptr;
    if (q->f.ckpt_every == 8) {
        if (z) sscan_bwd_kernel<T, 8, true><<<grid, 32, 0, st>>>(*q);
        else sscan_bwd_kernel<T, 8, false><<<grid, 32, 0, st>>>(*q);
    } else {
        if (z) sscan_bwd_kernel<T, 16, true><<<grid, 32, 0, st>>>(*q);
        else sscan_bwd_kernel<T, 16, false><<<grid, 32, 0, st>>>(*q);
    }
    return check_launch("sscan_bwd_kernel");
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
    return bytes ? bytes : sizeof(float);
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
