// sscan2.cu -- Mamba-1 selective scan, second-generation forward and backward kernels for sm_100a (fp32 I/O, 16-byte aligned
// layouts).
//
// Same operator and C ABI as sscan.cu (reference selective_scan_fwd_kernel.cuh:67-303 and selective_scan_bwd_kernel.cuh:75-489
// are what they replace); this file holds the kernels b200_sscan_fwd / b200_sscan_bwd dispatch to whenever the tensors can be
// described by TMA tensor maps (fp32, L % 4 == 0, 16-byte aligned strides) -- every launch the SS2D modules make.  Other layouts /
// 16-bit I/O stay on sscan.cu.  Measurements, pipe-level analysis and the A/B history: DESIGN.md 3.1, profiles/ncu_sscan_r02.md.
//
// Why a second design: ncu on the first one (profiles/ncu_sscan_r01.md) showed 5.8 warp-instructions and ~1.7 LSU wavefronts
// per (row, step) element against a floor of 1.5 / 0.3 -- the MUFU pipe (16 ex2 per element) should be the only limit.
//
//   * task = one warp = 16 channels of one (batch, group); lane = (row r = lane & 15, state half h = lane >> 4): a lane
//     owns ONE row and 8 states, carried as 4 packed f32x2 registers (two adjacent states per FFMA2/FMUL2).  B and C of two
//     adjacent states come out of shared memory as natural float2 pairs, so every product is packed; the only duplicated
//     operands are delta and delta*u (2 MOV per step).  y needs one add across the two halves -- no shared-memory reduction,
//     no exchange tile.
//   * all input tiles (u, delta: 16 rows x 32 steps; B, C: 16 states x 32 steps) arrive by TMA (cp.async.bulk.tensor.4d,
//     SWIZZLE_128B, zero fill outside the tensor) into a 2-stage ring, one mbarrier per stage: no LSU wavefronts, no address
//     arithmetic and no predicates for the loads; a 128-byte row per request instead of 32 bytes.
//   * per 8-step chunk a lane reads its own 4 steps (one LDS.128 per tensor), evaluates softplus once per element and trades
//     halves with its partner lane by shuffles; the B / C chunk is transposed once into a [step][state] tile (pitch 20 words:
//     conflict-free STS.32, broadcast LDS.128).  Reversed groups (rev_mask) differ only in shared-memory addresses, shuffle
//     source lanes and two register reversals per chunk.
//   * checkpoints keep the layout of sscan.cu ([chunk][state][row pairs]), so either generation's backward can consume them.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"

namespace b200 {
namespace v2 {

constexpr int NS = 16;             // states per task
constexpr int TR = 16;             // rows per task
constexpr int TC = 8;              // steps per chunk == checkpoint interval
constexpr int WIN = 32;            // steps per TMA window (one 128-byte row per tile row)
constexpr int CPW = WIN / TC;      // chunks per window
constexpr int TBP = 20;            // pitch (words) of the transposed B / C chunk tiles
constexpr int NSTAGE = 2;

struct FwdMaps {
    CUtensorMap u, delta, B, C;    // 4-D: (L, rows | states, groups, batch)
};

struct __align__(1024) FwdSmem {
    float win[NSTAGE][4][TR * WIN];   // [stage][u, delta, B, C]: 128-byte rows, 16-byte units XOR-swizzled with (row & 7)
    float tb[2][2][TC * TBP];         // [chunk parity][B, C][scan step][state]
    uint64_t bar[NSTAGE];
};

__device__ __forceinline__ float2 ex2_2(float2 e) { return make_float2(ex2(e.x), ex2(e.y)); }

template <bool HAS_Z, bool SOFTPLUS>
__global__ void __launch_bounds__(32, 11)
sscan_fwd2_kernel(const __grid_constant__ b200_sscan_fwd_params p, const __grid_constant__ FwdMaps tm, const unsigned tx_bytes,
                  const int zero_fill) {
    __shared__ FwdSmem sm;
    const unsigned sbase = smem_addr(&sm);   // one generic -> shared conversion for everything the TMA engine touches
    constexpr unsigned OFF_BAR = (unsigned)offsetof(FwdSmem, bar), STAGE_BYTES = 4u * TR * WIN * 4u, TILE_BYTES = (unsigned)TR * WIN * 4u;
    const int lane = threadIdx.x;
    const int r = lane & 15, h = lane >> 4;
    const int L = p.seqlen, N = p.dstate;
    const int rpg = p.dim / p.n_groups;
    const int tiles = (rpg + TR - 1) / TR;
    const int task = blockIdx.x;
    const int rt = task % tiles, bg = task / tiles;
    const int g = bg % p.n_groups, b = bg / p.n_groups;
    const int r0 = rt * TR;
    const bool row_ok = r0 + r < rpg;
    const int d = g * rpg + r0 + r;
    const bool rev = (p.rev_mask >> g) & 1u;
    const int nwin = (L + WIN - 1) / WIN, nck = (L + TC - 1) / TC;

    if (zero_fill) {   // boxes smaller than the tiles (N < 16 or fewer than 16 rows per group): the rest must read as 0
        float4* w4 = reinterpret_cast<float4*>(&sm.win[0][0][0]);
        for (int k = lane; k < NSTAGE * 4 * TR * WIN / 4; k += 32) w4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async();
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init_a(sbase + OFF_BAR + 8u * s, 1);
        mbar_init_fence();
    }
    __syncwarp();
    const int gu = g / p.u_group_div;
    auto issue = [&](int w) {   // one thread (elect_one): the four tiles of window w
        const unsigned s = (unsigned)w % NSTAGE;
        const int l_lo = rev ? L - (w + 1) * WIN : w * WIN;
        const unsigned bar = sbase + OFF_BAR + 8u * s, dst = sbase + s * STAGE_BYTES;   // win is the first member
        mbar_expect_tx_a(bar, tx_bytes);
        tma_load_4d_a(dst, &tm.u, l_lo, r0, gu, b, bar);
        tma_load_4d_a(dst + TILE_BYTES, &tm.delta, l_lo, r0, g, b, bar);
        tma_load_4d_a(dst + 2 * TILE_BYTES, &tm.B, l_lo, 0, g, b, bar);
        tma_load_4d_a(dst + 3 * TILE_BYTES, &tm.C, l_lo, 0, g, b, bar);
    };
    if (elect_one()) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s)
            if (s < nwin) issue(s);
    }

    // ---- per-lane constants ----
    float2 A2[4], x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int n0 = 8 * h + 2 * q;
        A2[q].x = (row_ok && n0 < N) ? __ldg(p.A + (size_t)d * N + n0) * kLog2e : 0.f;
        A2[q].y = (row_ok && n0 + 1 < N) ? __ldg(p.A + (size_t)d * N + n0 + 1) * kLog2e : 0.f;
        x[q] = make_float2(0.f, 0.f);
    }
    const float bias = (p.delta_bias && row_ok) ? __ldg(p.delta_bias + d) : 0.f;
    const float Dv = (p.D && row_ok) ? __ldg(p.D + d) : 0.f;
    float* o_row = (float*)p.out + (size_t)b * p.out_batch_stride + (size_t)d * p.out_row_stride;
    const float* z_row = HAS_Z ? (const float*)p.z + (size_t)b * p.z_batch_stride + (size_t)d * p.z_row_stride : nullptr;
    // checkpoint record of a chunk: word ckpt_state_pos(state) + 2 * (r & 7) + (r >> 3)  (common.cuh)
    float* ck = p.ckpt ? p.ckpt + (size_t)task * (size_t)(nck - 1) * NS * TR + 128 * h + 2 * (r & 7) + (r >> 3) : nullptr;

    const int rsw = r & 7;                       // swizzle key of this lane's row
    const int tn = lane >> 1, tpart = lane & 1;  // B / C transposition: state, 4-step half
    const int tsw = tn & 7;
    const int tb_first = (rev ? 7 - 4 * tpart : 4 * tpart) * TBP + tn;   // scan step of memory column 4 * tpart, then +-1 step
    const int tb_step = rev ? -TBP : TBP;
    const int src_lo = r + (rev ? 16 : 0), src_hi = r + (rev ? 0 : 16);  // lanes holding scan steps 0-3 / 4-7 of this row
    const bool hi_half = (h != 0) != rev;                                // this lane's own columns are scan steps 4-7

    for (int w = 0; w < nwin; ++w) {
        const int s = w % NSTAGE;
        mbar_wait_a(sbase + OFF_BAR + 8u * (unsigned)s, (w / NSTAGE) & 1);
        const float* wu = sm.win[s][0] + r * WIN;
        const float* wd = sm.win[s][1] + r * WIN;
        const float* wB = sm.win[s][2] + tn * WIN;
        const float* wC = sm.win[s][3] + tn * WIN;
        const int kmax = min(CPW, nck - w * CPW);
        for (int k = 0; k < kmax; ++k) {
            const int c = w * CPW + k;                 // chunk index in scan order
            const int jw = rev ? CPW - 1 - k : k;      // its 32-byte slot inside the window rows
            const int l4 = (rev ? L - (c + 1) * TC : c * TC) + 4 * h;   // memory position of this lane's first column
            const bool ok4 = (unsigned)l4 < (unsigned)L;                // L % 4 == 0: its four columns are all in or all out
            const float4 uv = *reinterpret_cast<const float4*>(wu + (((2 * jw + h) ^ rsw) << 2));
            const float4 dv = *reinterpret_cast<const float4*>(wd + (((2 * jw + h) ^ rsw) << 2));
            {   // B / C chunk -> [scan step][state]
                const float4 bv = *reinterpret_cast<const float4*>(wB + (((2 * jw + tpart) ^ tsw) << 2));
                const float4 cv = *reinterpret_cast<const float4*>(wC + (((2 * jw + tpart) ^ tsw) << 2));
                float* tB = sm.tb[c & 1][0] + tb_first;
                float* tC = sm.tb[c & 1][1] + tb_first;
                tB[0] = bv.x; tB[tb_step] = bv.y; tB[2 * tb_step] = bv.z; tB[3 * tb_step] = bv.w;
                tC[0] = cv.x; tC[tb_step] = cv.y; tC[2 * tb_step] = cv.z; tC[3 * tb_step] = cv.w;
            }
            // ---- once per element: delta' = softplus(delta + bias), q = delta' * u (identity step outside the sequence) ----
            float ou[4] = {uv.x, uv.y, uv.z, uv.w};
            float odl[4] = {dv.x, dv.y, dv.z, dv.w}, oq[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = odl[e] + bias;
                if (SOFTPLUS) a = softplus_sigmoid(a).sp;
                odl[e] = ok4 ? a : 0.f;
                oq[e] = odl[e] * ou[e];
            }
            if (rev) {   // own registers in scan order
                float t;
                t = odl[0]; odl[0] = odl[3]; odl[3] = t; t = odl[1]; odl[1] = odl[2]; odl[2] = t;
                t = oq[0]; oq[0] = oq[3]; oq[3] = t; t = oq[1]; oq[1] = oq[2]; oq[2] = t;
            }
            float dl[TC], q[TC];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dl[e] = __shfl_sync(0xffffffffu, odl[e], src_lo);
                dl[4 + e] = __shfl_sync(0xffffffffu, odl[e], src_hi);
                q[e] = __shfl_sync(0xffffffffu, oq[e], src_lo);
                q[4 + e] = __shfl_sync(0xffffffffu, oq[e], src_hi);
            }
            __syncwarp();   // the transposed B / C tiles are complete
            if (ck != nullptr && c > 0) {
                float* cc = ck + (size_t)(c - 1) * NS * TR;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {   // state 8 h + 2 qd (+ 1); the 128 h term is in ck
                    __stcs(cc + ckpt_state_pos(2 * qd), x[qd].x);
                    __stcs(cc + ckpt_state_pos(2 * qd + 1), x[qd].y);
                }
            }
            // ---- recurrence: 8 steps x 4 state pairs ----
            const float* tB = sm.tb[c & 1][0] + 8 * h;
            const float* tC = sm.tb[c & 1][1] + 8 * h;
            float yp[TC];
#pragma unroll
            for (int st = 0; st < TC; ++st) {
                const float2 dd = make_float2(dl[st], dl[st]), qq = make_float2(q[st], q[st]);
#ifdef B200_DBG_NO_BC
                const float4 b0 = make_float4(dl[st], q[st], dl[st], q[st]), b1 = b0, c0 = b0, c1 = b0;
#else
                const float4 b0 = *reinterpret_cast<const float4*>(tB + st * TBP), b1 = *reinterpret_cast<const float4*>(tB + st * TBP + 4);
                const float4 c0 = *reinterpret_cast<const float4*>(tC + st * TBP), c1 = *reinterpret_cast<const float4*>(tC + st * TBP + 4);
#endif
                const float2 B2[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
                const float2 C2[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
                float2 y2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
#ifdef B200_DBG_NO_EX2
                    const float2 a = __fmul2_rn(dd, A2[qd]);
#else
                    const float2 a = ex2_2(__fmul2_rn(dd, A2[qd]));
#endif
                    x[qd] = __ffma2_rn(a, x[qd], __fmul2_rn(qq, B2[qd]));
                    y2 = __ffma2_rn(C2[qd], x[qd], y2);
                }
                yp[st] = y2.x + y2.y;
            }
            // ---- y: add the two state halves, each lane keeps its own four columns ----
            float keep[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float mine = hi_half ? yp[4 + e] : yp[e], theirs = hi_half ? yp[e] : yp[4 + e];
                keep[e] = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
            }
            if (rev) {
                float t;
                t = keep[0]; keep[0] = keep[3]; keep[3] = t; t = keep[1]; keep[1] = keep[2]; keep[2] = t;
            }
            float4 o = make_float4(fmaf(Dv, ou[0], keep[0]), fmaf(Dv, ou[1], keep[1]), fmaf(Dv, ou[2], keep[2]), fmaf(Dv, ou[3], keep[3]));
            if (ok4 && row_ok) {
                if (HAS_Z) {
                    const float4 zz = __ldcs(reinterpret_cast<const float4*>(z_row + l4));
                    o.x *= zz.x * sigmoidf_(zz.x); o.y *= zz.y * sigmoidf_(zz.y);
                    o.z *= zz.z * sigmoidf_(zz.z); o.w *= zz.w * sigmoidf_(zz.w);
                }
                __stcs(reinterpret_cast<float4*>(o_row + l4), o);
            }
        }
        __syncwarp();   // every lane is done with this window's tiles
        if (w + NSTAGE < nwin && elect_one()) {
            fence_proxy_async();
            issue(w + NSTAGE);
        }
    }
    if (p.last_state != nullptr && row_ok) {
        float* ls = p.last_state + ((size_t)b * p.dim + d) * N + 8 * h;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            if (8 * h + 2 * qd < N) ls[2 * qd] = x[qd].x;
            if (8 * h + 2 * qd + 1 < N) ls[2 * qd + 1] = x[qd].y;
        }
    }
}


// ----------------------------------------------------------------------------------------------------------------
// backward (replaces selective_scan_bwd_kernel.cuh:75-489 for the TMA-describable layouts)
//
//   * lane = (state quad sq = lane >> 3, row pair i = lane & 7) as in sscan.cu: rows (i, i + 8) ride in the two halves of
//     packed f32x2 registers, the lane owns 4 of the 16 states.  What changed is everything around the arithmetic:
//   * inputs (u, delta, dout: 16 rows x 16 steps; B, C: 16 states x 16 steps) arrive by TMA into a 3-stage ring and the
//     OUTPUTS LEAVE FROM THE SAME TILES: du overwrites u, ddelta overwrites delta, dB / dC overwrite B / C element by element
//     (each element is consumed before its gradient exists), then one TMA store (du, ddelta) or TMA reduce-add (dB, dC: the
//     fp32 atomics of selective_scan_bwd_kernel.cuh:447-461 become cp.reduce.async.bulk) per tile and window.  No global
//     load / store / atomic instruction and no address arithmetic is left in the chunk loop except the checkpoint read.
//   * the channel reduction of dB / dC (sum over the 16 rows = 8 lanes x 2 halves) is a reduce-scatter whose first two rounds
//     need no selects: lane i keeps its four states in the slot order  n = 4 sq + (slot ^ pm(i)),  pm = (bit 2, bit 1) of i,
//     so "keep slots 0-1, send slots 2-3" (partner i ^ 4) and "keep slot 0, send slot 1" (partner i ^ 2) are the same
//     registers in every lane; only the last round (partner i ^ 1, over steps) selects.  56 SHFL + 56 FADD + 16 SEL per chunk
//     instead of 56 + 56 + 112, and the slots are visited in the order 3, 1, 2, 0 so only 16 partial sums are ever parked.
//   * reversed groups read B / C from a reversed copy of the chunk (2 LDS.128 + 2 STS.128 per lane); everything else about
//     direction is a shared-memory address.
// ----------------------------------------------------------------------------------------------------------------
constexpr int WB = 16;             // steps per TMA window (64-byte tile rows, SWIZZLE_64B)
constexpr int CPWB = WB / TC;      // chunks per window
constexpr int NSTB = 2;            // ring stages: a stage is refilled once its window's stores have left it
constexpr int EXS = 24;            // exchange tiles: words per scan step (16 used; 24 keeps the STS.64 conflict free)
constexpr int RDS = 144;           // state-reduction tile: words per state quad
constexpr float kLn2 = 0.6931471805599453f;
#ifndef B200_BWD_REGS
#define B200_BWD_REGS 168   // 3 warps per SM sub-partition (16 K registers each): 12 per SM, one wave at the stage-0 shape
#endif
#ifndef B200_BWD_PRIV
#define B200_BWD_PRIV 3     // A/B over the ten scan calls of a MedMamba-T step (tools/dbg_variants.sh, two runs each): 2 -> 5.37 ms, 3 -> 5.29 ms
#endif
constexpr int PRIV_SLOTS = B200_BWD_PRIV;   // 1: dA in shared memory; 2: dA and the adjoint carry; 3: and A log2(e) (fewest spills)

struct BwdMaps {
    CUtensorMap u, delta, dout, B, C;   // loads
    CUtensorMap du, ddelta, dB, dC;     // stores (du, ddelta) / reduce-adds (dB, dC) from the same shared-memory tiles
};

struct __align__(1024) BwdSmem {
    float win[NSTB][5][TR * WB];        // [stage][u -> du, delta -> sigmoid -> ddelta, dout, B -> dB, C -> dC]
    union {
        struct {
            float rvs[2][NS * TC];      // time-reversed copy of the B / C chunk (reversed groups only)
            float exq[TC * EXS];        // delta' * u, [scan step][row pair][2]
        } a;
        float rd[4 * RDS];              // reduction over the state quads (after the state loop)
    } un;
    float exd[TC * EXS];                // delta'
    float exg[TC * EXS];                // dout
    float2 priv[PRIV_SLOTS][4][32];     // per-lane accumulators that do not fit in registers: running dA (and the adjoint carry)
    float ck[2][NS * TR];               // [chunk parity] state entering the chunk: the 1 KB record as stored (ckpt_state_pos, common.cuh)
    uint64_t bar[NSTB];
    uint64_t ckbar[2];
};

__device__ __forceinline__ int rd_pos(int quad, int cc, int i) { return quad * RDS + ((cc * 16 + 2 * i) ^ (((cc >> 1) & 1) << 4)); }
// Sum 8 per-lane row-pair values over the 4 state quads: the lane ends with the totals of scan steps s0, s1.
__device__ __forceinline__ void reduce_states(float* tile, const float2 (&v)[TC], int sq, int i, int s0, int s1, float2& r0, float2& r1) {
#pragma unroll
    for (int s = 0; s < TC; ++s) *reinterpret_cast<float2*>(tile + rd_pos(sq, s, i)) = v[s];
    __syncwarp();
    r0 = *reinterpret_cast<const float2*>(tile + rd_pos(0, s0, i));
    r1 = *reinterpret_cast<const float2*>(tile + rd_pos(0, s1, i));
#pragma unroll
    for (int qd = 1; qd < 4; ++qd) {
        r0 = __fadd2_rn(r0, *reinterpret_cast<const float2*>(tile + rd_pos(qd, s0, i)));
        r1 = __fadd2_rn(r1, *reinterpret_cast<const float2*>(tile + rd_pos(qd, s1, i)));
    }
}

struct BwdChunkCtx {
    float2 dl2[TC], q2[TC], go2[TC];   // delta', delta' * u, dout of the lane's row pair, by scan step
    float2 s1[TC], s2[TC];             // sums over the lane's states: g * B  and  A log2e * (a g x_prev)
};

// One state of the lane: recompute its 8 steps from the checkpoint, run the adjoint recurrence, leave the row-pair sums of
// the dB / dC contributions in vB / vC.
// MODE says what happens to the row-pair sums vB / vC of every step (the select-free rounds of the channel reduction):
//   0: park = partner(i ^ 4)'s sums        1: park = partner(i ^ 2)'s (sums + park)
//   2: park += partner(i ^ 4)'s sums       3: park += sums
#ifdef B200_DBG_NO_SHFL
#define BWD_SHFL_XOR(v, m) (v)
#else
#define BWD_SHFL_XOR(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#endif
template <int MODE>
__device__ __forceinline__ void bwd_state(BwdChunkCtx& cx, const float* bB, const float* bC, const int off, const float2 A2j, const float2 xm1,
                                          float2& hj, float2& dAj, float (&park)[2 * TC]) {
    // B / C of this state in scan order: 4 steps at word `off` of the tile (or reversed copy) bB / bC, the other 4 at off ^ 4
    float Bv[TC], Cv[TC];
    {
        const float4 lo = *reinterpret_cast<const float4*>(bB + off);
        const float4 hi = *reinterpret_cast<const float4*>(bB + (off ^ 4));
        Bv[0] = lo.x; Bv[1] = lo.y; Bv[2] = lo.z; Bv[3] = lo.w; Bv[4] = hi.x; Bv[5] = hi.y; Bv[6] = hi.z; Bv[7] = hi.w;
    }
    float2 a[TC], x[TC];
    {
        float2 xs = xm1;
#pragma unroll
        for (int s = 0; s < TC; ++s) {
#ifdef B200_DBG_NO_EX2
            a[s] = __fmul2_rn(cx.dl2[s], A2j);
#else
            a[s] = ex2_2(__fmul2_rn(cx.dl2[s], A2j));
#endif
            xs = __ffma2_rn(a[s], xs, __fmul2_rn(cx.q2[s], make_float2(Bv[s], Bv[s])));
            x[s] = xs;
        }
    }
    {
        const float4 lo = *reinterpret_cast<const float4*>(bC + off);
        const float4 hi = *reinterpret_cast<const float4*>(bC + (off ^ 4));
        Cv[0] = lo.x; Cv[1] = lo.y; Cv[2] = lo.z; Cv[3] = lo.w; Cv[4] = hi.x; Cv[5] = hi.y; Cv[6] = hi.z; Cv[7] = hi.w;
    }
    float2 gn = hj, dAp = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = TC - 1; s >= 0; --s) {
        const float2 xprev = s > 0 ? x[s - 1] : xm1;
        const float2 g = __ffma2_rn(cx.go2[s], make_float2(Cv[s], Cv[s]), gn);
        if (MODE == 0) {
            park[TC + s] = BWD_SHFL_XOR(fmaf(cx.go2[s].y, x[s].y, cx.go2[s].x * x[s].x), 4);   // summed over the row pair
            park[s] = BWD_SHFL_XOR(fmaf(g.y, cx.q2[s].y, g.x * cx.q2[s].x), 4);
        } else if (MODE == 1) {
            park[TC + s] = BWD_SHFL_XOR(fmaf(cx.go2[s].y, x[s].y, fmaf(cx.go2[s].x, x[s].x, park[TC + s])), 2);
            park[s] = BWD_SHFL_XOR(fmaf(g.y, cx.q2[s].y, fmaf(g.x, cx.q2[s].x, park[s])), 2);
        } else if (MODE == 2) {
            park[TC + s] += BWD_SHFL_XOR(fmaf(cx.go2[s].y, x[s].y, cx.go2[s].x * x[s].x), 4);
            park[s] += BWD_SHFL_XOR(fmaf(g.y, cx.q2[s].y, g.x * cx.q2[s].x), 4);
        } else {
            park[TC + s] = fmaf(cx.go2[s].y, x[s].y, fmaf(cx.go2[s].x, x[s].x, park[TC + s]));
            park[s] = fmaf(g.y, cx.q2[s].y, fmaf(g.x, cx.q2[s].x, park[s]));
        }
        cx.s1[s] = __ffma2_rn(g, make_float2(Bv[s], Bv[s]), cx.s1[s]);
        gn = __fmul2_rn(a[s], g);
        const float2 wv = __fmul2_rn(gn, xprev);
        cx.s2[s] = __ffma2_rn(A2j, wv, cx.s2[s]);                  // A log2(e): rescaled by ln 2 in the epilogue
        dAp = __ffma2_rn(wv, cx.dl2[s], dAp);
    }
    hj = gn;
    dAj = __fadd2_rn(dAj, dAp);
}

template <int MODE, int J>
__device__ __forceinline__ void bwd_slot(BwdChunkCtx& cx, const float* pB, const float* pC, const int off, const float2 A2j, const float2 xm1,
                                         float2& hj, float2& dAj, float2 (*priv)[4][32], int lane, float (&park)[2 * TC]) {
    if (PRIV_SLOTS == 0) {
        bwd_state<MODE>(cx, pB, pC, off, A2j, xm1, hj, dAj, park);
    } else if (PRIV_SLOTS == 1) {
        float2 dA = make_float2(0.f, 0.f);
        bwd_state<MODE>(cx, pB, pC, off, A2j, xm1, hj, dA, park);
        priv[0][J][lane] = __fadd2_rn(priv[0][J][lane], dA);
    } else {
        float2 dA = make_float2(0.f, 0.f), h = priv[1][J][lane];
        bwd_state<MODE>(cx, pB, pC, off, PRIV_SLOTS >= 3 ? priv[2][J][lane] : A2j, xm1, h, dA, park);
        priv[1][J][lane] = h;
        priv[0][J][lane] = __fadd2_rn(priv[0][J][lane], dA);
    }
}

template <bool SOFTPLUS>
__global__ void __maxnreg__(B200_BWD_REGS)
sscan_bwd2_kernel(const __grid_constant__ b200_sscan_bwd_params q, const __grid_constant__ BwdMaps tm, const unsigned tx_bytes,
                  const int zero_fill) {
    __shared__ BwdSmem sm;
    // shared-memory addresses of everything the bulk-copy engine touches, from ONE generic -> shared conversion
    const unsigned sbase = smem_addr(&sm);
    constexpr unsigned OFF_WIN = (unsigned)offsetof(BwdSmem, win), OFF_CK = (unsigned)offsetof(BwdSmem, ck);
    constexpr unsigned OFF_BAR = (unsigned)offsetof(BwdSmem, bar), OFF_CKBAR = (unsigned)offsetof(BwdSmem, ckbar);
    constexpr unsigned STAGE_BYTES = 5u * TR * WB * 4u, TILE_BYTES = (unsigned)TR * WB * 4u;
    const b200_sscan_fwd_params& p = q.f;
    const int lane = threadIdx.x;
    const int sq = lane >> 3, i = lane & 7;
    const int pm = (((i >> 2) & 1) << 1) | ((i >> 1) & 1);   // slot j of this lane holds state 4 sq + (j ^ pm)
    const int L = p.seqlen, N = p.dstate;
    const int rpg = p.dim / p.n_groups;
    const int tiles = (rpg + TR - 1) / TR;
    const int task = blockIdx.x;
    const int rt = task % tiles, bg = task / tiles;
    const int g = bg % p.n_groups, b = bg / p.n_groups;
    const int r0 = rt * TR;
    const bool okA = r0 + i < rpg, okB = r0 + i + 8 < rpg;
    const int chA = g * rpg + r0 + i, chB = chA + 8;
    const bool rev = (p.rev_mask >> g) & 1u;
    const int nwin = (L + WB - 1) / WB, nck = (L + TC - 1) / TC;

    if (zero_fill) {
        float4* w4 = reinterpret_cast<float4*>(&sm.win[0][0][0]);
        for (int k = lane; k < NSTB * 5 * TR * WB / 4; k += 32) w4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async();
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTB; ++s) mbar_init_a(sbase + OFF_BAR + 8u * s, 1);
        mbar_init_a(sbase + OFF_CKBAR, 1);
        mbar_init_a(sbase + OFF_CKBAR + 8u, 1);
        mbar_init_fence();
    }
    __syncwarp();
    const int gu = g / p.u_group_div, gd = g / (int)q.dout_group_div;
    auto win_lo = [&](int v) {   // memory position of the first column of the v-th window VISITED (scan windows last -> first)
        const int wv = nwin - 1 - v;
        return rev ? L - (wv + 1) * WB : wv * WB;
    };
    auto issue = [&](int v) {   // one thread (elect_one): the five tiles of the v-th window visited
        const unsigned s = (unsigned)v % NSTB;
        const int l_lo = win_lo(v);
        const unsigned bar = sbase + OFF_BAR + 8u * s, w = sbase + OFF_WIN + s * STAGE_BYTES;
        mbar_expect_tx_a(bar, tx_bytes);
        tma_load_4d_a(w, &tm.u, l_lo, r0, gu, b, bar);
        tma_load_4d_a(w + TILE_BYTES, &tm.delta, l_lo, r0, g, b, bar);
        tma_load_4d_a(w + 2 * TILE_BYTES, &tm.dout, l_lo, r0, gd, b, bar);
        tma_load_4d_a(w + 3 * TILE_BYTES, &tm.B, l_lo, 0, g, b, bar);
        tma_load_4d_a(w + 4 * TILE_BYTES, &tm.C, l_lo, 0, g, b, bar);
    };
    if (elect_one()) {
        issue(0);
        if (nwin > 1) issue(1);
    }

    // ---- per-lane constants and accumulators ----
    float2 A2[4], hh[4], dAacc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = 4 * sq + (j ^ pm);
        A2[j].x = (okA && n < N) ? __ldg(p.A + (size_t)chA * N + n) * kLog2e : 0.f;
        A2[j].y = (okB && n < N) ? __ldg(p.A + (size_t)chB * N + n) * kLog2e : 0.f;
        hh[j] = make_float2(0.f, 0.f);
        dAacc[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < PRIV_SLOTS; ++k) sm.priv[k][j][lane] = k == 2 ? A2[j] : make_float2(0.f, 0.f);   // private to this lane: no synchronisation
    }
    const float biasA = (p.delta_bias && okA) ? __ldg(p.delta_bias + chA) : 0.f;
    const float biasB = (p.delta_bias && okB) ? __ldg(p.delta_bias + chB) : 0.f;
    const float DA = (p.D && okA) ? __ldg(p.D + chA) : 0.f;
    const float DB = (p.D && okB) ? __ldg(p.D + chB) : 0.f;
    float dD_A = 0.f, dD_B = 0.f, dbias_A = 0.f, dbias_B = 0.f;
    const float* ck = p.ckpt + (size_t)task * (size_t)(nck - 1) * NS * TR;
    // checkpoint of the state entering chunk c (c >= 1): the 1 KB record -> tile ck[c & 1], ONE bulk copy (the record layout is
    // already the conflict-free shared-memory layout)
    auto issue_ck = [&](int c) {   // one thread (elect_one)
        const unsigned bar = sbase + OFF_CKBAR + 8u * (c & 1);
        mbar_expect_tx_a(bar, NS * TR * 4);
        bulk_load_1d_a(sbase + OFF_CK + (unsigned)(c & 1) * (NS * TR * 4), ck + (size_t)(c - 1) * NS * TR, NS * TR * 4, bar);
    };
    if (nck > 1 && elect_one()) issue_ck(nck - 1);
    int ck_uses = 0;   // completed waits on the checkpoint barriers (chunks are visited in descending order: parity alternates)

    const int c0 = 2 * sq;                                           // this lane's two memory columns of every chunk
    const int st0 = rev ? TC - 1 - c0 : c0, st1 = rev ? TC - 2 - c0 : c0 + 1;   // their scan steps
    const int rkey = (i >> 1) & 3;                                   // SWIZZLE_64B key of rows i and i + 8
    const int tn = lane >> 1, tpart = lane & 1;                      // reversed B / C copy: state, 4-step half
    const int b0 = i & 1;

    int refill = -1;   // window whose loads wait for the stage's stores to have left (issued one chunk into the next window)
    for (int v = 0; v < nwin; ++v) {
        const int wv = nwin - 1 - v;
        const int s = v % NSTB;
        mbar_wait_a(sbase + OFF_BAR + 8u * (unsigned)s, (v / NSTB) & 1);
        float* wu = sm.win[s][0];
        float* wd = sm.win[s][1];
        const float* wg = sm.win[s][2];
        float* wB = sm.win[s][3];
        float* wC = sm.win[s][4];
        const int kmax = min(CPWB, nck - wv * CPWB);
        for (int k = kmax - 1; k >= 0; --k) {
            const int c = wv * CPWB + k;                 // chunk index in scan order
            const int jw = rev ? CPWB - 1 - k : k;       // its 32-byte slot inside the window rows
            const int la = (rev ? L - (c + 1) * TC : c * TC) + c0;
            const bool v01 = (unsigned)la < (unsigned)L; // L % 4 == 0 and c0 even: both columns in or both out
            // checkpoint of the state entering this chunk: prefetched by a bulk copy while the previous chunk was computed
            if (c > 1 && elect_one()) issue_ck(c - 1);   // tile (c - 1) & 1 was last read two chunks ago
            if (c > 0) mbar_wait_a(sbase + OFF_CKBAR + 8u * (c & 1), (ck_uses >> 1) & 1);
            ck_uses += c > 0;
            const float* ckt = &sm.ck[c & 1][sq * 4 * TR + 2 * i];
            auto load_ck = [&](int j) {
                return c > 0 ? *reinterpret_cast<const float2*>(ckt + (((j ^ pm) ^ (sq & 1)) << 4)) : make_float2(0.f, 0.f);
            };
            const int own = (((2 * jw + (sq >> 1)) ^ rkey) << 2) + 2 * (sq & 1);   // word offset of (c0, c0 + 1) inside a tile row
            const int offA = i * WB + own, offB = (i + 8) * WB + own;
            {   // ---- prologue (once per element): delta', its sigmoid, delta' * u ----
                const float2 ua = *reinterpret_cast<const float2*>(wu + offA), ub = *reinterpret_cast<const float2*>(wu + offB);
                const float2 da = *reinterpret_cast<const float2*>(wd + offA), db = *reinterpret_cast<const float2*>(wd + offB);
                const float2 ga = *reinterpret_cast<const float2*>(wg + offA), gb = *reinterpret_cast<const float2*>(wg + offB);
                float dlA[2] = {da.x, da.y}, dlB[2] = {db.x, db.y}, sgA[2], sgB[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float a = dlA[e] + biasA, bb = dlB[e] + biasB, sa = 1.f, sb = 1.f;
                    if (SOFTPLUS) {
                        const SoftplusSig ra = softplus_sigmoid(a), rb = softplus_sigmoid(bb);
                        a = ra.sp; sa = ra.sig; bb = rb.sp; sb = rb.sig;
                    }
                    dlA[e] = v01 ? a : 0.f; sgA[e] = v01 ? sa : 0.f;   // identity steps outside the sequence
                    dlB[e] = v01 ? bb : 0.f; sgB[e] = v01 ? sb : 0.f;
                }
                *reinterpret_cast<float2*>(&sm.exd[st0 * EXS + 2 * i]) = make_float2(dlA[0], dlB[0]);
                *reinterpret_cast<float2*>(&sm.exd[st1 * EXS + 2 * i]) = make_float2(dlA[1], dlB[1]);
                *reinterpret_cast<float2*>(&sm.un.a.exq[st0 * EXS + 2 * i]) = make_float2(dlA[0] * ua.x, dlB[0] * ub.x);
                *reinterpret_cast<float2*>(&sm.un.a.exq[st1 * EXS + 2 * i]) = make_float2(dlA[1] * ua.y, dlB[1] * ub.y);
                *reinterpret_cast<float2*>(&sm.exg[st0 * EXS + 2 * i]) = make_float2(ga.x, gb.x);
                *reinterpret_cast<float2*>(&sm.exg[st1 * EXS + 2 * i]) = make_float2(ga.y, gb.y);
                // park sigmoid(delta + bias) where delta was: only this lane looks there again
                *reinterpret_cast<float2*>(wd + offA) = make_float2(sgA[0], sgA[1]);
                *reinterpret_cast<float2*>(wd + offB) = make_float2(sgB[0], sgB[1]);
            }
            if (rev) {   // B / C of this chunk in scan order
                const int src = tn * WB + (((2 * jw + tpart) ^ ((tn >> 1) & 3)) << 2);
                const float4 bv = *reinterpret_cast<const float4*>(wB + src), cv = *reinterpret_cast<const float4*>(wC + src);
                *reinterpret_cast<float4*>(&sm.un.a.rvs[0][tn * TC + 4 * (1 - tpart)]) = make_float4(bv.w, bv.z, bv.y, bv.x);
                *reinterpret_cast<float4*>(&sm.un.a.rvs[1][tn * TC + 4 * (1 - tpart)]) = make_float4(cv.w, cv.z, cv.y, cv.x);
            }
            __syncwarp();
            BwdChunkCtx cx;
#pragma unroll
            for (int cc = 0; cc < TC; ++cc) {
                cx.dl2[cc] = *reinterpret_cast<const float2*>(&sm.exd[cc * EXS + 2 * i]);
                cx.q2[cc] = *reinterpret_cast<const float2*>(&sm.un.a.exq[cc * EXS + 2 * i]);
                cx.go2[cc] = *reinterpret_cast<const float2*>(&sm.exg[cc * EXS + 2 * i]);
                cx.s1[cc] = make_float2(0.f, 0.f);
                cx.s2[cc] = make_float2(0.f, 0.f);
            }
            // B / C rows of slot j in scan order: 4 steps at word offset bc_off(j) of the window tile (or of the reversed copy),
            // the other 4 at bc_off(j) ^ 4
            const float* bB = rev ? sm.un.a.rvs[0] : wB;
            const float* bC = rev ? sm.un.a.rvs[1] : wC;
            auto bc_off = [&](int j) {
                const int n = 4 * sq + (j ^ pm);
                return rev ? n * TC : n * WB + (((2 * jw) ^ ((n >> 1) & 3)) << 2);
            };
            float park[2 * TC];
            // slot 3: its sums go to the partner i ^ 4, whose slot 3 is this lane's slot 1
            bwd_slot<0, 3>(cx, bB, bC, bc_off(3), A2[3], load_ck(3), hh[3], dAacc[3], sm.priv, lane, park);
            // slot 1 (+ the partner's slot 3), then to the partner i ^ 2, whose slot 1 is this lane's slot 0
            bwd_slot<1, 1>(cx, bB, bC, bc_off(1), A2[1], load_ck(1), hh[1], dAacc[1], sm.priv, lane, park);
            // slot 2: to the partner i ^ 4, whose slot 2 is this lane's slot 0
            bwd_slot<2, 2>(cx, bB, bC, bc_off(2), A2[2], load_ck(2), hh[2], dAacc[2], sm.priv, lane, park);
            // slot 0: totals over lanes {i, i^2, i^4, i^6}; the last round, over steps, with the partner i ^ 1
            bwd_slot<3, 0>(cx, bB, bC, bc_off(0), A2[0], load_ck(0), hh[0], dAacc[0], sm.priv, lane, park);
            float kB[4], kC[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                kB[e] = (b0 ? park[4 + e] : park[e]) + __shfl_xor_sync(0xffffffffu, b0 ? park[e] : park[4 + e], 1);
                kC[e] = (b0 ? park[TC + 4 + e] : park[TC + e]) + __shfl_xor_sync(0xffffffffu, b0 ? park[TC + e] : park[TC + 4 + e], 1);
            }
            __syncwarp();   // every lane has read this chunk's B / C, exq and the reversed copies
            {   // dB / dC of state 4 sq + pm (slot 0), scan steps 4 b0 .. 4 b0 + 3, written where B / C of those steps were
                const int n = 4 * sq + pm;
                const int half = rev ? 1 - b0 : b0;
                const int dst = n * WB + (((2 * jw + half) ^ ((n >> 1) & 3)) << 2);
                const float4 vb = rev ? make_float4(kB[3], kB[2], kB[1], kB[0]) : make_float4(kB[0], kB[1], kB[2], kB[3]);
                const float4 vc = rev ? make_float4(kC[3], kC[2], kC[1], kC[0]) : make_float4(kC[0], kC[1], kC[2], kC[3]);
                *reinterpret_cast<float4*>(wB + dst) = vb;
                *reinterpret_cast<float4*>(wC + dst) = vc;
            }
            // ---- per-element gradients: sums over the 16 states, then this lane's 2 rows x 2 columns ----
            float2 t1a, t1b, t2a, t2b;
            reduce_states(sm.un.rd, cx.s1, sq, i, st0, st1, t1a, t1b);
            __syncwarp();
            reduce_states(sm.un.rd, cx.s2, sq, i, st0, st1, t2a, t2b);
            {
                const float2 ua = *reinterpret_cast<const float2*>(wu + offA), ub = *reinterpret_cast<const float2*>(wu + offB);
                const float2 sa = *reinterpret_cast<const float2*>(wd + offA), sb = *reinterpret_cast<const float2*>(wd + offB);
                const float2 dl0 = *reinterpret_cast<const float2*>(&sm.exd[st0 * EXS + 2 * i]);   // (row A, row B) of column c0
                const float2 dl1 = *reinterpret_cast<const float2*>(&sm.exd[st1 * EXS + 2 * i]);
                const float2 go0 = *reinterpret_cast<const float2*>(&sm.exg[st0 * EXS + 2 * i]);
                const float2 go1 = *reinterpret_cast<const float2*>(&sm.exg[st1 * EXS + 2 * i]);
                const float duA0 = fmaf(dl0.x, t1a.x, DA * go0.x), duA1 = fmaf(dl1.x, t1b.x, DA * go1.x);
                const float duB0 = fmaf(dl0.y, t1a.y, DB * go0.y), duB1 = fmaf(dl1.y, t1b.y, DB * go1.y);
                // chain rule through softplus: sg = sigmoid(delta + bias) (1 without softplus, 0 outside the sequence)
                const float ddA0 = fmaf(ua.x, t1a.x, t2a.x * kLn2) * sa.x, ddA1 = fmaf(ua.y, t1b.x, t2b.x * kLn2) * sa.y;
                const float ddB0 = fmaf(ub.x, t1a.y, t2a.y * kLn2) * sb.x, ddB1 = fmaf(ub.y, t1b.y, t2b.y * kLn2) * sb.y;
                dbias_A += ddA0 + ddA1;
                dbias_B += ddB0 + ddB1;
                dD_A = fmaf(go0.x, ua.x, fmaf(go1.x, ua.y, dD_A));
                dD_B = fmaf(go0.y, ub.x, fmaf(go1.y, ub.y, dD_B));
                *reinterpret_cast<float2*>(wu + offA) = make_float2(duA0, duA1);
                *reinterpret_cast<float2*>(wu + offB) = make_float2(duB0, duB1);
                *reinterpret_cast<float2*>(wd + offA) = make_float2(ddA0, ddA1);
                *reinterpret_cast<float2*>(wd + offB) = make_float2(ddB0, ddB1);
            }
            __syncwarp();   // the union (rd) and the exchange tiles are free for the next chunk
            if (refill >= 0) {   // the previous window's stores were issued a whole chunk ago: they have read their tiles by now
                if (elect_one()) {
                    tma_wait_read<0>();
                    issue(refill);
                }
                refill = -1;
            }
        }
        // ---- the window's outputs leave from its own tiles ----
        const int l_lo = win_lo(v);
        if (l_lo < 0) {
            // TMA stores / reductions fault on negative coordinates (loads zero-fill them): the one window of a reversed group
            // that hangs over position 0 (L % 16 != 0) is written element by element -- once per task.
            __syncwarp();
            const int nrows = min(TR, rpg - r0);
#pragma unroll 1
            for (int e = lane; e < TR * WB; e += 32) {
                const int row = e >> 4, col = e & 15, l = l_lo + col;
                const int off = row * WB + (((col >> 2) ^ ((row >> 1) & 3)) << 2) + (col & 3);
                if (l >= 0 && l < L) {
                    if (row < nrows) {
                        ((float*)q.du)[(size_t)b * q.du_batch_stride + (size_t)(g * rpg + r0 + row) * q.du_row_stride + l] = wu[off];
                        ((float*)q.ddelta)[(size_t)b * q.ddelta_batch_stride + (size_t)g * q.ddelta_group_stride +
                                           (size_t)(r0 + row) * q.ddelta_row_stride + l] = wd[off];
                    }
                    if (row < N) {
                        atomicAdd(q.dB + (size_t)b * q.dB_batch_stride + (size_t)g * q.dB_group_stride + (size_t)row * q.dB_state_stride + l, wB[off]);
                        atomicAdd(q.dC + (size_t)b * q.dC_batch_stride + (size_t)g * q.dC_group_stride + (size_t)row * q.dC_state_stride + l, wC[off]);
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
        } else {
            fence_proxy_async();
            __syncwarp();
            if (elect_one()) {
                const unsigned w = sbase + OFF_WIN + (unsigned)s * STAGE_BYTES;
                tma_store_4d_a(&tm.du, w, l_lo, r0, g, b);
                tma_store_4d_a(&tm.ddelta, w + TILE_BYTES, l_lo, r0, g, b);
                tma_reduce_add_4d_a(&tm.dB, w + 3 * TILE_BYTES, l_lo, 0, g, b);
                tma_reduce_add_4d_a(&tm.dC, w + 4 * TILE_BYTES, l_lo, 0, g, b);
                tma_commit_group();
            }
        }
        if (v + NSTB < nwin) refill = v + NSTB;   // this stage is refilled once its stores have read it
    }
    if (elect_one()) tma_wait_read<0>();   // elect.sync is deterministic: the thread that committed every store group

#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = 4 * sq + (j ^ pm);
        if (n < N) {
            const float2 dAv = PRIV_SLOTS >= 1 ? sm.priv[0][j][lane] : dAacc[j];
            if (okA) atomicAdd(q.dA + (size_t)chA * N + n, dAv.x);
            if (okB) atomicAdd(q.dA + (size_t)chB * N + n, dAv.y);
        }
    }
#pragma unroll
    for (int m = 8; m <= 16; m <<= 1) {
        dD_A += __shfl_xor_sync(0xffffffffu, dD_A, m);
        dD_B += __shfl_xor_sync(0xffffffffu, dD_B, m);
        dbias_A += __shfl_xor_sync(0xffffffffu, dbias_A, m);
        dbias_B += __shfl_xor_sync(0xffffffffu, dbias_B, m);
    }
    if (sq == 0) {
        if (okA) {
            if (q.ddelta_bias) atomicAdd(q.ddelta_bias + chA, dbias_A);
            if (q.dD) atomicAdd(q.dD + chA, dD_A);
        }
        if (okB) {
            if (q.ddelta_bias) atomicAdd(q.ddelta_bias + chB, dbias_B);
            if (q.dD) atomicAdd(q.dD + chB, dD_B);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int64_t fix_stride(int64_t s, int n) { return n == 1 && (s <= 0 || (s & 3)) ? 4 : s; }   // a size-1 dimension's stride is free

static bool make_fwd_maps(FwdMaps* tm, const b200_sscan_fwd_params* p, unsigned* tx_bytes, int* zero_fill) {
    const int rpg = p->dim / p->n_groups;
    const int L = p->seqlen, N = p->dstate, G = p->n_groups, Bt = p->batch;
    if (p->u_group_div < 1 || G % p->u_group_div) return false;
    const int Gu = G / p->u_group_div;
    const int box_rows = rpg < TR ? rpg : TR, box_n = N < NS ? N : NS;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    if (!tma_make_map_f32(&tm->u, p->u, L, rpg, Gu, Bt, fix_stride(p->u_row_stride, rpg), fix_stride(p->u_group_stride, Gu),
                          fix_stride(p->u_batch_stride, Bt), WIN, box_rows, sw))
        return false;
    if (!tma_make_map_f32(&tm->delta, p->delta, L, rpg, G, Bt, fix_stride(p->delta_row_stride, rpg), fix_stride(p->delta_group_stride, G),
                          fix_stride(p->delta_batch_stride, Bt), WIN, box_rows, sw))
        return false;
    if (!tma_make_map_f32(&tm->B, p->B, L, N, G, Bt, fix_stride(p->B_state_stride, N), fix_stride(p->B_group_stride, G),
                          fix_stride(p->B_batch_stride, Bt), WIN, box_n, sw))
        return false;
    if (!tma_make_map_f32(&tm->C, p->C, L, N, G, Bt, fix_stride(p->C_state_stride, N), fix_stride(p->C_group_stride, G),
                          fix_stride(p->C_batch_stride, Bt), WIN, box_n, sw))
        return false;
    *tx_bytes = 2u * box_rows * WIN * 4u + 2u * box_n * WIN * 4u;
    *zero_fill = (box_rows < TR || box_n < NS) ? 1 : 0;
    return true;
}

static bool env_off(const char* name) {
    const char* v = getenv(name);
    return v && v[0] == '1';
}

// 1: launched (or failed with an error set, rc in *rc); 0: not eligible, caller falls back to sscan.cu
bool try_fwd(const b200_sscan_fwd_params* p, cudaStream_t st, int* rc) {
    static const bool off = env_off("B200_SSCAN_V1");
    if (off || p->io_dtype != B200_F32 || p->dstate > NS || (p->seqlen & 3)) return false;
    if (!aligned16(p->out) || (p->out_row_stride & 3) || (p->out_batch_stride & 3)) return false;
    if (p->z && (!aligned16(p->z) || (p->z_row_stride & 3) || (p->z_batch_stride & 3))) return false;
    FwdMaps tm;
    unsigned tx = 0;
    int zf = 0;
    if (!make_fwd_maps(&tm, p, &tx, &zf)) return false;
    const int rpg = p->dim / p->n_groups;
    const long long nt = (long long)p->batch * p->n_groups * ((rpg + TR - 1) / TR);
    auto launch = [&](auto kernel) {
        (void)func_attr_per_device((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kernel<<<(unsigned)nt, 32, 0, st>>>(*p, tm, tx, zf);
    };
    if (p->z) {
        if (p->delta_softplus) launch(sscan_fwd2_kernel<true, true>);
        else launch(sscan_fwd2_kernel<true, false>);
    } else {
        if (p->delta_softplus) launch(sscan_fwd2_kernel<false, true>);
        else launch(sscan_fwd2_kernel<false, false>);
    }
    *rc = check_launch("sscan_fwd2_kernel");
    return true;
}


static bool make_bwd_maps(BwdMaps* tm, const b200_sscan_bwd_params* q, unsigned* tx_bytes, int* zero_fill) {
    const b200_sscan_fwd_params* p = &q->f;
    const int rpg = p->dim / p->n_groups;
    const int L = p->seqlen, N = p->dstate, G = p->n_groups, Bt = p->batch;
    if (p->u_group_div < 1 || G % p->u_group_div || q->dout_group_div < 1 || G % (int)q->dout_group_div) return false;
    const int Gu = G / p->u_group_div, Gd = G / (int)q->dout_group_div;
    const int box_rows = rpg < TR ? rpg : TR, box_n = N < NS ? N : NS;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_64B;
    auto rows = [&](CUtensorMap* m, const void* base, int groups, int64_t rs, int64_t gs, int64_t bs) {
        return tma_make_map_f32(m, base, L, rpg, groups, Bt, fix_stride(rs, rpg), fix_stride(gs, groups), fix_stride(bs, Bt), WB, box_rows, sw);
    };
    auto states = [&](CUtensorMap* m, const void* base, int64_t ss, int64_t gs, int64_t bs) {
        return tma_make_map_f32(m, base, L, N, G, Bt, fix_stride(ss, N), fix_stride(gs, G), fix_stride(bs, Bt), WB, box_n, sw);
    };
    if (!rows(&tm->u, p->u, Gu, p->u_row_stride, p->u_group_stride, p->u_batch_stride)) return false;
    if (!rows(&tm->delta, p->delta, G, p->delta_row_stride, p->delta_group_stride, p->delta_batch_stride)) return false;
    if (!rows(&tm->dout, q->dout, Gd, q->dout_row_stride, q->dout_group_stride, q->dout_batch_stride)) return false;
    if (!rows(&tm->du, q->du, G, q->du_row_stride, (int64_t)rpg * q->du_row_stride, q->du_batch_stride)) return false;
    if (!rows(&tm->ddelta, q->ddelta, G, q->ddelta_row_stride, q->ddelta_group_stride, q->ddelta_batch_stride)) return false;
    if (!states(&tm->B, p->B, p->B_state_stride, p->B_group_stride, p->B_batch_stride)) return false;
    if (!states(&tm->C, p->C, p->C_state_stride, p->C_group_stride, p->C_batch_stride)) return false;
    if (!states(&tm->dB, q->dB, q->dB_state_stride, q->dB_group_stride, q->dB_batch_stride)) return false;
    if (!states(&tm->dC, q->dC, q->dC_state_stride, q->dC_group_stride, q->dC_batch_stride)) return false;
    *tx_bytes = 3u * box_rows * WB * 4u + 2u * box_n * WB * 4u;
    *zero_fill = (box_rows < TR || box_n < NS) ? 1 : 0;
    return true;
}

bool try_bwd(const b200_sscan_bwd_params* q, cudaStream_t st, int* rc) {
    static const bool off = env_off("B200_SSCAN_V1") || env_off("B200_SSCAN_BWD_V1");
    const b200_sscan_fwd_params* p = &q->f;
    if (off || p->io_dtype != B200_F32 || p->dstate > NS || (p->seqlen & 3) || p->z != nullptr) return false;
    BwdMaps tm;
    unsigned tx = 0;
    int zf = 0;
    if (!make_bwd_maps(&tm, q, &tx, &zf)) return false;
    const int rpg = p->dim / p->n_groups;
    const long long nt = (long long)p->batch * p->n_groups * ((rpg + TR - 1) / TR);
    (void)func_attr_per_device((const void*)sscan_bwd2_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    (void)func_attr_per_device((const void*)sscan_bwd2_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (p->delta_softplus) sscan_bwd2_kernel<true><<<(unsigned)nt, 32, 0, st>>>(*q, tm, tx, zf);
    else sscan_bwd2_kernel<false><<<(unsigned)nt, 32, 0, st>>>(*q, tm, tx, zf);
    *rc = check_launch("sscan_bwd2_kernel");
    return true;
}

}  // namespace v2
}  // namespace b200
