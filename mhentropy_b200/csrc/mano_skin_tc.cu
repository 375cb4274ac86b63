// Linear blend skinning of the MANO mesh on the tensor cores (reference hand/manopth/manolayer.py:232-246:
//   th_T = th_results2 . weights^T ;  verts = (th_T . [v_posed; 1])[:3] ), then centring and metres -> millimetres (:262-273).
//
// The per-vertex transform blend  T[r][v][e] = sum_j weights[v][j] A[r][j][e]  (e < 12: the 3 x 4 transform of joint j for row r) is
// 192 FMAs per (row, vertex): on the CUDA cores that made the skinning kernel instruction-bound at 12 % of the HBM peak (round 1).
// Here it is a K = 16 tcgen05 contraction with the vertices on M, so every thread of the epilogue owns ONE vertex and finds its
// twelve blended coefficients for a row in its own TMEM lane:
//
//   D[v][(r, e)] = sum_j W[v][j] . A^T[(r, e)][j]      M = 128 vertices,  N = 16 rows x 12 = 192,  K = 16 (one MMA k-step)
//
// * A CTA keeps the skinning weights of its 128 vertices resident (split half planes, 32 KB) and walks the row tiles (persistent).
// * Four producer warps convert the joint transforms A [R][16][12] fp32 of a 16-row tile into the transposed split half planes in
//   the SW128 K-major operand layout (double-buffered); one thread issues the three split-precision products (hi*hi, hi*lo, lo*hi).
// * The accumulator is double-buffered in TMEM (2 x 192 columns): eight epilogue warps apply the transform of tile i to v_posed (read
//   from the blend GEMM's output, 384 contiguous bytes per row and warp) and store the vertices while tile i + 1 is contracted.
// The kernel moves exactly its algorithmic bytes: v_posed read + vertices written (2 x 9,336 B per row) + the transforms (768 B).
#include "mano_math.cuh"
#include "tc_gemm.cuh"

namespace mhe {
namespace skin {
using namespace tc;

constexpr int kV = MHE_MANO_VERTS, kNJ = MHE_MANO_JOINTS, kJ = 16;
constexpr int kRows = 16;                      // rows per tile
constexpr int kN = kRows * 12;                 // accumulator columns per tile
constexpr int kWPlane = 128 * 128;             // bytes: 128 vertices x 64 halves (K padded to one swizzle row; 16 used)
constexpr int kBPlane = kN * 128;              // bytes: 192 (row, e) operand rows x 64 halves
constexpr int kStage = 2 * kBPlane;            // hi + lo
constexpr int kVpRow = 128 * 3 * 4;             // bytes of one row of a v_posed / vertex tile: 128 vertices x 3 floats
constexpr int kVpStage = kRows * kVpRow + 256;  // 24,576 B of v_posed + the 16 rows' centre joints (16 B each)
constexpr int kVpStages = 3;                   // ~50 KB of v_posed in flight per SM: what ~5 TB/s needs at HBM latency
constexpr int kSmem = 2 * kWPlane + 2 * kStage + kVpStages * kVpStage + 1024;
constexpr int kThreads = 14 * 32;              // warps 0-7 epilogue, 8-11 producers, 12 MMA + TMEM, 13 v_posed loader (bulk copies)
constexpr float kMM = 1000.f;

__constant__ int c_tip_vert[5] = {745, 317, 444, 556, 673};                         // manolayer.py:250
__constant__ int c_jtr_src[2][kNJ] = {{0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20},      // manolayer.py:260
                                      {0, 16, 15, 14, 13, 17, 3, 2, 1, 18, 6, 5, 4, 19, 12, 11, 10, 20, 9, 8, 7}};     // + utils.py:15

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {   // 32 columns of this lane; the caller waits (tcgen05.wait::ld)
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the 16 joint coefficients A[r][j][e], j = 0..15, of operand row n = (rr, e) of a tile
__device__ __forceinline__ void load_unit(const float* __restrict__ A, int R, int r0, int n, float (&a)[16]) {
    const int rr = n / 12, e = n - rr * 12, r = r0 + rr;
    const float* src = A + (size_t)(r < R ? r : 0) * 192 + e;
    const float keep = r < R ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = keep * __ldg(src + j * 12);
}

struct Args {
    const uint16_t* wplanes;      // [2][896][64] half planes of the skinning weights (mhe_mano_pack_posedirs_planes)
    const float* A;               // [R][16][12] joint transforms (mano_pose_fwd_kernel)
    const float* cen;             // [R][4] centre joint
    const float* vp;              // [R][ld_vp] v_posed (template + shape blend + pose blend)
    float* verts;                 // [R][778][3] mm
    float* jtr;                   // [R][21][3] mm: the five tip joints are written here (may be NULL)
    int R, ld_vp, order, tiles, slots;
    int dbg;                      // MHE_SKIN_DEBUG bits (timing experiments only): 1 no vertex stores, 2 no coefficient loads, 4 no v_posed copies, 8 no MMAs
};

__global__ void __launch_bounds__(kThreads, 1) mano_skin_tc_kernel(Args p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_acc_full[2], bar_acc_empty[2], bar_vp_full[kVpStages], bar_vp_empty[kVpStages];
    __shared__ uint32_t tmem_slot;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t wbase = smem0, bbase = smem0 + 2 * kWPlane, vbase = bbase + 2 * kStage;
    float* vs_gen = reinterpret_cast<float*>(smem_raw + (vbase - smem_u32(smem_raw)));      // generic pointer to the v_posed ring
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vt = blockIdx.x % 7, slot = blockIdx.x / 7;      // 7 vertex tiles of 128 cover the 778 vertices
    const int v0 = vt * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 4); mbar_init(smem_u32(&bar_empty[s]), 1);          // arrivals: one elected lane per warp
            mbar_init(smem_u32(&bar_acc_full[s]), 1); mbar_init(smem_u32(&bar_acc_empty[s]), 8);
        }
        for (int s = 0; s < kVpStages; ++s) { mbar_init(smem_u32(&bar_vp_full[s]), 1); mbar_init(smem_u32(&bar_vp_empty[s]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the CTA's weight planes, resident for the whole kernel: [plane][128 vertices][8 chunks of 16 B], chunks swizzled (SW128 K-major)
    for (int i = threadIdx.x; i < 2 * 128 * 8; i += kThreads) {
        const int pl = i >> 10, row = (i >> 3) & 127, ch = i & 7;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.wplanes + ((size_t)pl * 896 + v0 + row) * 64) + ch);
        st_shared_v4(wbase + pl * kWPlane + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4), v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp >= 8 && warp < 12) {
        // ===================== producers: A^T of the tile as split half planes, operand row n = (row, e), k = joint
        // Each thread owns operand rows n = t and (t < 64) t + 128 of every tile.  The coefficients of the NEXT tile are loaded into
        // registers before the ring slot of the current one is waited for: with two slots the chain load -> convert -> MMA -> free would
        // otherwise expose one HBM round trip per tile (measured: the epilogue warps waited for accumulators 18 % of the time).
        const int t = threadIdx.x - 256;
        const bool two = t < kN - 128;
        float an0[16], an1[16];
        if (slot < p.tiles) {
            load_unit(p.A, p.R, slot * kRows, t, an0);
            if (two) load_unit(p.A, p.R, slot * kRows, t + 128, an1);
        }
        uint32_t it = 0;
        for (int tile = slot; tile < p.tiles; tile += p.slots, ++it) {
            const uint32_t s = it & 1;
            mbar_wait(smem_u32(&bar_empty[s]), ((it >> 1) & 1) ^ 1);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                const int n = t + u * 128;
                uint4 h0, l0, h1, l1;
                split8<true>(u == 0 ? an0 : an1, h0, l0);
                split8<true>((u == 0 ? an0 : an1) + 8, h1, l1);
                const uint32_t row = bbase + s * kStage + (n >> 3) * 1024 + (n & 7) * 128;
                st_shared_v4(row + ((0 ^ (n & 7)) << 4), h0);
                st_shared_v4(row + ((1 ^ (n & 7)) << 4), h1);
                st_shared_v4(row + kBPlane + ((0 ^ (n & 7)) << 4), l0);
                st_shared_v4(row + kBPlane + ((1 ^ (n & 7)) << 4), l1);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(smem_u32(&bar_full[s]));      // (hundreds of arrivals on one barrier per tile serialise)
            if (tile + p.slots < p.tiles && !(p.dbg & 2)) {      // the next tile's coefficients: in flight while the ring slot is waited for
                load_unit(p.A, p.R, (tile + p.slots) * kRows, t, an0);
                if (two) load_unit(p.A, p.R, (tile + p.slots) * kRows, t + 128, an1);
            }
        }
    } else if (warp == 13) {
        // ===================== v_posed loader: 16 bulk copies per tile (one row of the CTA's 128 vertices each: 1536 contiguous bytes)
        // into a ring, kVpStages tiles ahead of the epilogue - the kernel is a stream of independent 24 KB tiles and only lives at HBM
        // speed with tens of KB in flight per SM (the register-prefetch version ran at 1.3 TB/s)
        {   // (the 17 copies of a tile are issued by 17 lanes at once: one thread issuing them back to back costs ~100 ns per copy)
            const uint32_t rowbytes = vt < 6 ? (uint32_t)kVpRow : (uint32_t)((p.ld_vp - 6 * 128 * 3) * 4);   // last tile: what is left of the row
            uint32_t it = 0;
            for (int tile = slot; tile < p.tiles; tile += p.slots, ++it) {
                const uint32_t s = it % kVpStages, use = it / kVpStages;
                const int r0 = tile * kRows, nr = (p.dbg & 4) ? 0 : min(kRows, p.R - r0);
                if (lane == 0) {
                    mbar_wait(smem_u32(&bar_vp_empty[s]), (use & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&bar_vp_full[s]), (rowbytes + 16) * nr);      // (one arrival + the bytes of the rows that exist)
                }
                __syncwarp();
                if (lane < nr)
                    bulk_load(vbase + s * kVpStage + lane * kVpRow, p.vp + (size_t)(r0 + lane) * p.ld_vp + vt * 128 * 3, rowbytes, smem_u32(&bar_vp_full[s]));
                else if (lane == 16 && nr > 0)
                    bulk_load(vbase + s * kVpStage + kRows * kVpRow, p.cen + (size_t)r0 * 4, 16 * nr, smem_u32(&bar_vp_full[s]));
            }
        }
    } else if (warp == 12) {
        // ===================== MMA issuer: D = W_hi.B_hi + W_hi.B_lo + W_lo.B_hi, one k-step (K = 16 of the 64-wide swizzle row)
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(kN, false, false, true, true);
            constexpr uint32_t kHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));      // SBO = 1024 | version | SWIZZLE_128B
            constexpr uint32_t kLo = (16u >> 4) << 16;
            uint32_t it = 0;
            for (int tile = slot; tile < p.tiles; tile += p.slots, ++it) {
                const uint32_t s = it & 1, ph = (it >> 1) & 1;
                mbar_wait(smem_u32(&bar_acc_empty[s]), ph ^ 1);
                mbar_wait(smem_u32(&bar_full[s]), ph);
                tcgen05_fence_after();
                const uint32_t acc = tmem + s * 256;
                const uint32_t a_hi = kLo | ((wbase >> 4) & 0x3FFFu), a_lo = kLo | (((wbase + kWPlane) >> 4) & 0x3FFFu);
                const uint32_t b_hi = kLo | (((bbase + s * kStage) >> 4) & 0x3FFFu), b_lo = kLo | (((bbase + s * kStage + kBPlane) >> 4) & 0x3FFFu);
                if (!(p.dbg & 8)) {
                    umma_bf16_lohi(acc, a_hi, b_hi, kHi, idesc, 0u);
                    umma_bf16_lohi(acc, a_hi, b_lo, kHi, idesc, 1u);
                    umma_bf16_lohi(acc, a_lo, b_hi, kHi, idesc, 1u);
                }
                tcgen05_commit(smem_u32(&bar_empty[s]));
                tcgen05_commit(smem_u32(&bar_acc_full[s]));
            }
        }
    } else {
        // ===================== epilogue: thread = vertex (TMEM lane), warp group h handles rows 8 h .. 8 h + 7 of the tile
        const int q = warp & 3, h = warp >> 2;
        const int v = v0 + q * 32 + lane;
        const bool vok = v < kV;
        int tip_slot = -1;                                     // output joint this vertex is (the five finger tips), if any
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (c_tip_vert[k] == v)
                for (int i = 0; i < kNJ; ++i) if (c_jtr_src[p.order][i] == kJ + k) tip_slot = i;
        uint32_t it = 0;
        for (int tile = slot; tile < p.tiles; tile += p.slots, ++it) {
            const uint32_t s = it & 1, vs_i = it % kVpStages;
            const int r0 = tile * kRows + h * 8;
            float* vs = vs_gen + vs_i * (kVpStage / 4) + h * 8 * (kVpRow / 4);      // this warp group's 8 rows of the tile [row][384 floats]
            const float* cs = vs_gen + vs_i * (kVpStage / 4) + kRows * (kVpRow / 4) + h * 8 * 4;   // ... and their centre joints [row][4]
            mbar_wait(smem_u32(&bar_vp_full[vs_i]), (it / kVpStages) & 1);
            mbar_wait(smem_u32(&bar_acc_full[s]), (it >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t acc = tmem + s * 256 + ((uint32_t)(q * 32) << 16) + h * 96;
            // the 8 rows' 96 blended coefficients of this vertex: three TMEM loads in flight, one wait
            float t[96];
            tmem_ld32_nowait(acc, t);
            tmem_ld32_nowait(acc + 32, t + 32);
            tmem_ld32_nowait(acc + 64, t + 64);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(smem_u32(&bar_acc_empty[s]));      // everything read: the accumulator goes back before the stores
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int r = r0 + rr;
                if (!vok || r >= p.R) continue;
                const float* T = t + rr * 12;
                float* pv = vs + rr * (kVpRow / 4) + (q * 32 + lane) * 3;       // stride 3 floats across the lanes: conflict-free
                const float p0 = pv[0], p1 = pv[1], p2 = pv[2];
                float o[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    o[i] = (T[i * 3 + 0] * p0 + T[i * 3 + 1] * p1 + T[i * 3 + 2] * p2 + T[9 + i] - cs[rr * 4 + i]) * kMM;
                pv[0] = o[0]; pv[1] = o[1]; pv[2] = o[2];                        // the tile becomes the output tile in place
                if (tip_slot >= 0 && p.jtr) { float* dj = p.jtr + ((size_t)r * kNJ + tip_slot) * 3; dj[0] = o[0]; dj[1] = o[1]; dj[2] = o[2]; }
            }
            __syncwarp();
            // the warp's 96 floats of every row are contiguous in the output: three fully coalesced stores per row
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int r = r0 + rr;
                if (r >= p.R) break;
                const float* src = vs + rr * (kVpRow / 4) + q * 96;
                float* dst = p.verts + (size_t)r * (kV * 3) + v0 * 3 + q * 96;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = k * 32 + lane;
                    if (v0 * 3 + q * 96 + i < kV * 3 && !(p.dbg & 1)) dst[i] = src[i];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy accesses before the next bulk copy into the slot
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(smem_u32(&bar_vp_empty[vs_i]));          // the slot may be refilled
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int launch(const uint16_t* wplanes, const float* A, const float* cen, const float* vp, int ld_vp, float* verts, float* jtr, int R, int order,
           cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(mano_skin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) {
            set_error("mano skinning: cannot raise dynamic shared memory to %d", kSmem);
            return MHE_ERR_CUDA;
        }
        attr_set = true;
    }
    static const int dbg = [] { const char* e = getenv("MHE_SKIN_DEBUG"); return e ? atoi(e) : 0; }();
    Args a{wplanes, A, cen, vp, verts, jtr, R, ld_vp, order, cdiv(R, kRows), 0, dbg};
    a.slots = a.tiles < 21 ? a.tiles : 21;         // 7 vertex tiles x 21 row-tile walkers = 147 CTAs on 148 SMs
    ProbeScope probe("mano skinning tc", stream);
    mano_skin_tc_kernel<<<7 * a.slots, kThreads, kSmem, stream>>>(a);
    return check_launch("mano skinning tc");
}

}  // namespace skin
}  // namespace mhe
