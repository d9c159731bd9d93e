// Persistent fused decoder GEMMs + NB-mixture likelihood (forward and backward sweeps), the north-star kernels.
//
// Why persistent: the one-tile-per-CTA kernels (nb_tc.cu / nb_tc_bwd.cu) spend a TMA + MMA phase of several microseconds per
// CTA during which their 8 epilogue warps idle, only two such CTAs fit an SM's tensor memory, and 316 tiles over 148 SMs
// leave 40 % of the SM-time idle (ncu r1: sm__cycles_active.avg 47.5 k of 80.3 k elapsed cycles).  Here one CTA per SM
// walks a contiguous range of fine-grained units (128 cells x 16 genes; 1252 units at C2, 8 or 9 per CTA):
//   * the A operand ([hm | zz] of the row tile, 128 x 320 bf16 = 80 KB) is loaded ONCE per row tile and stays in shared memory;
//   * warp 0 streams the units' weight tiles (mixture weight k-blocks + folded private / shared weights, 14 KB) through a
//     2-slot ring with TMA; warp 1 issues the tcgen05 MMAs into one of two TMEM accumulator buffers (pi | lp | ls);
//   * two epilogue groups of 4 warps alternate units (group = unit parity = ring slot = accumulator buffer), so the loads and
//     MMAs of unit t + 1 overlap the likelihood math of unit t, and a thread reads its cell's 16 raw counts (32 bytes) straight
//     from the count matrix, prefetched before it waits for the accumulator.
// Row partial sums are carried in registers across the units of a row tile and written once per (CTA, row tile, group).
// Reference: nn/networks.py:314-325, module/spVIPESmodule.py:751-759, 817-824; scvi log_mixture_nb.
#include <cstdlib>
#include "nb_ptc.cuh"
#include "nb_math.cuh"
#include "decoder_common.cuh"
#include "../../include/spvipes_b200.h"

namespace ptc {

static long long* g_trace = nullptr;
extern "C" int spv_debug_trace(long long* buf) {  // diagnostic hook (tools/ptc_trace.py); not part of the documented ABI
    g_trace = buf;
    return 0;
}
extern "C" long long* spv_debug_get_trace() { return g_trace; }
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The persistent kernels are opt-in (SPV_NB_PERSISTENT=1): at the BASELINE shapes the r1 version runs at par with the
// one-tile-per-CTA kernels (profiles/r1_nb_persistent_notes.md), so those stay the default.
bool enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SPV_NB_PERSISTENT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = MAX_CTAS;
        n = v < MAX_CTAS ? v : MAX_CTAS;
    }
    return n;
}

namespace {

constexpr int EPI_GROUPS = 2, EPI_GROUP_THREADS = 128, EPI_GROUP_WARPS = 4;
constexpr int THREADS = 64 + EPI_GROUPS * EPI_GROUP_THREADS;  // 320
constexpr int A_KB_BYTES = BM * BK * 2;                       // 16384
constexpr int B_KB_BYTES = BN * BK * 2;                       // 2048
constexpr int SLOT_BYTES = (KB_MAX + 2) * B_KB_BYTES;         // mixture-weight k-blocks | folded private | folded shared
constexpr int ACC_STRIDE = 64;                                // TMEM columns per accumulator buffer: pi +0, lp +16, ls +32
constexpr int TMEM_COLS = 128;
constexpr int NGC = 7;                                        // per-gene constants per unit (forward uses 6)
constexpr int OFF_B = KB_MAX * A_KB_BYTES;
constexpr int OFF_GC = OFF_B + 2 * SLOT_BYTES;
constexpr int OFF_LUT = OFF_GC + EPI_GROUPS * 2 * NGC * BN * 4;
constexpr int LUT_N = 128;                                    // raw counts below this come from the table
constexpr int OFF_BAR = OFF_LUT + LUT_N * 8;
constexpr int SMEM_BYTES = OFF_BAR + 128;  // no alignment slack: the dynamic segment is declared 1024-byte aligned (checked)
static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs (one per group of the step) must fit one SM");

struct Bars {
    uint64_t a_full, a_free, b_full[2], b_empty[2], acc_full[2], acc_empty[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ void group_barrier(int eg) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + eg), "r"(EPI_GROUP_THREADS) : "memory");
}

// warp 0: TMA producer.  A once per row tile; one ring slot per unit.
__device__ __forceinline__ void producer(const CUtensorMap* mapA, const CUtensorMap* mapB, const CUtensorMap* mapZ, uint8_t* sm,
                                         Bars* bar, Range rg, int nG, int num_kb, int kb_z, int Gp) {
    int seg = 0, prev_r = -1;
    for (int t = 0; t < rg.u1 - rg.u0; ++t) {
        const int u = rg.u0 + t, r = u / nG, j = u - r * nG;
        if (r != prev_r) {
            if (prev_r >= 0) tc::mbar_wait(&bar->a_free, (seg - 1) & 1);  // every MMA reading the previous row tile is done
            tc::mbar_expect_tx(&bar->a_full, num_kb * A_KB_BYTES);
            for (int kb = 0; kb < num_kb; ++kb) tc::tma_load_2d(mapA, &bar->a_full, sm + kb * A_KB_BYTES, kb * BK, r * BM);
            prev_r = r;
            ++seg;
        }
        const int s = t & 1, n = t >> 1;
        tc::mbar_wait(&bar->b_empty[s], (n & 1) ^ 1);
        uint8_t* slot = sm + OFF_B + s * SLOT_BYTES;
        tc::mbar_expect_tx(&bar->b_full[s], (num_kb + 2) * B_KB_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tc::tma_load_2d(mapB, &bar->b_full[s], slot + kb * B_KB_BYTES, kb * BK, j * BN);
        tc::tma_load_2d(mapZ, &bar->b_full[s], slot + KB_MAX * B_KB_BYTES, kb_z * BK, j * BN);             // folded private
        tc::tma_load_2d(mapZ, &bar->b_full[s], slot + (KB_MAX + 1) * B_KB_BYTES, kb_z * BK, Gp + j * BN);  // folded shared
    }
}

// warp 1: MMA issuer.  pi over all k-blocks, lp / ls on the latent k-block against the folded weights.
__device__ __forceinline__ void mma_issuer(uint8_t* sm, Bars* bar, uint32_t tmem_base, Range rg, int nG, int num_kb, int kb_z) {
    constexpr uint32_t idesc = tc::idesc_bf16(BM, BN, false, false);
    const uint32_t a_base = tc::smem_u32(sm);
    int seg = 0, prev_r = -1;
    const int nu = rg.u1 - rg.u0;
    for (int t = 0; t < nu; ++t) {
        const int u = rg.u0 + t, r = u / nG;
        if (r != prev_r) {
            tc::mbar_wait(&bar->a_full, seg & 1);
            prev_r = r;
            ++seg;
        }
        const int s = t & 1, n = t >> 1;
        tc::mbar_wait(&bar->b_full[s], n & 1);
        tc::mbar_wait(&bar->acc_empty[s], (n & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t acc = tmem_base + s * ACC_STRIDE;
        const uint32_t slot = tc::smem_u32(sm + OFF_B + s * SLOT_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
                tc::umma_bf16(acc, tc::smem_desc(a_base + kb * A_KB_BYTES + kk * 32, 16, 1024),
                              tc::smem_desc(slot + kb * B_KB_BYTES + kk * 32, 16, 1024), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = tc::smem_desc(a_base + kb_z * A_KB_BYTES + kk * 32, 16, 1024);
            tc::umma_bf16(acc + BN, da, tc::smem_desc(slot + KB_MAX * B_KB_BYTES + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
            tc::umma_bf16(acc + 2 * BN, da, tc::smem_desc(slot + (KB_MAX + 1) * B_KB_BYTES + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
        }
        tc::umma_commit(&bar->b_empty[s]);
        tc::umma_commit(&bar->acc_full[s]);
        if (t + 1 < nu && (u + 1) / nG != r) tc::umma_commit(&bar->a_free);
    }
}

// a cell's 16 raw counts of the unit (uint16 source): two 16-byte loads when the segment is aligned and inside the matrix
__device__ __forceinline__ void load_counts16(uint32_t (&cw)[8], const unsigned short* X16, long xrow, int n0, int G, bool mok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) cw[i] = 0u;
    if (!mok) return;
    const unsigned short* src = X16 + xrow + n0;
    if (n0 + BN <= G && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src)), b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
        cw[0] = a.x; cw[1] = a.y; cw[2] = a.z; cw[3] = a.w; cw[4] = b.x; cw[5] = b.y; cw[6] = b.z; cw[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = n0 + 2 * i;
            const uint32_t c0 = g < G ? __ldg(src + 2 * i) : 0u, c1 = g + 1 < G ? __ldg(src + 2 * i + 1) : 0u;
            cw[i] = c0 | (c1 << 16);
        }
    }
}

// slow form of one element (count beyond the table or rho below ~1e-6): one out-of-line copy, called from the epilogue
__device__ __noinline__ NbOut fwd_slow(bool is_count, uint32_t c, float tv, float accp, float accs, float accpi, NbGene ge, float Rpl,
                                       float Rsl) {
    float2 tl;
    if (is_count) tl = nb_count_terms_exact(c);
    else { tl.x = tv; tl.y = lgamma_pos_fast(tv + 1.0f); }
    bool dummy = false;
    return nb_forward_v3<true>(tl.x, tl.y, accp, accs, accpi, ge, Rpl, Rsl, dummy);
}

template <int SRC>
__global__ void __launch_bounds__(384, 2) nb_ptc_fwd_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapB,
                                                                const __grid_constant__ CUtensorMap mapZ, FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sm = smem_raw;
    if (tc::smem_u32(sm) & 1023u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    float* s_gc = reinterpret_cast<float*>(sm + OFF_GC);     // [group][2][NGC][BN]
    float2* s_lut = reinterpret_cast<float2*>(sm + OFF_LUT);  // [LUT_N]: (log1p(c), lgamma(log1p(c) + 1)) per raw count
    Bars* bar = reinterpret_cast<Bars*>(sm + OFF_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Range rg = cta_range(blockIdx.x, gridDim.x, p.units);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&mapA);
        tc::tma_prefetch_desc(&mapB);
        tc::tma_prefetch_desc(&mapZ);
        tc::mbar_init(&bar->a_full, 1);
        tc::mbar_init(&bar->a_free, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bar->b_full[s], 1);
            tc::mbar_init(&bar->b_empty[s], 1);
            tc::mbar_init(&bar->acc_full[s], 1);
            tc::mbar_init(&bar->acc_empty[s], EPI_GROUP_WARPS);
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(&bar->tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = bar->tmem_slot;

    if (warp == 0) {
        if (tc::elect_one()) producer(&mapA, &mapB, &mapZ, sm, bar, rg, p.nG, p.num_kb, p.kb_z, p.Gp);
    } else if (warp == 1) {
        if (tc::elect_one()) mma_issuer(sm, bar, tmem_base, rg, p.nG, p.num_kb, p.kb_z);
    } else {
        const int eg = (warp - 2) >> 2;                              // epilogue group = unit parity
        const int et = threadIdx.x - 64 - eg * EPI_GROUP_THREADS;    // 0..127 within the group
        const int q = warp & 3;                                      // TMEM lane quarter this warp may access
        const int rloc = q * 32 + lane;
        const long G = p.G;
        if (SRC == SPV_SRC_U16_LOG1P) {
            if (threadIdx.x - 64 < LUT_N) nb_fill_count_lut(s_lut, threadIdx.x - 64);
            asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(eg * ACC_STRIDE);
        const bool vec_pi = p.pi && ((G & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.pi) & 15) == 0);
        // This group's units t = eg, eg + 2, ... of the CTA's range, software pipelined: the per-gene constants (one value per
        // loader thread) and the thread's 16 raw counts of unit t + 2 are fetched into registers while unit t is computed, so
        // no global-memory latency sits between two units.  Row partial sums: one set per row tile of the range (at most two).
        const int nu = rg.u1 - rg.u0;
        const int r_first = rg.u0 / p.nG;
        float acc_seg[2][3] = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
        auto row_setup = [&](int r, bool& mok, float& Rpl, float& Rsl, long& xrow, int& m) {
            m = r * BM + rloc;
            mok = m < p.B;
            const int mm = mok ? m : 0;
            Rpl = NB_LOG2E * __ldg(p.rowc + (long)mm * 4 + 0);
            Rsl = NB_LOG2E * __ldg(p.rowc + (long)mm * 4 + 1);
            xrow = (p.rows ? (long)__ldg(p.rows + mm) : (long)mm) * p.ldx;
        };
        // per-gene constants of nb_forward_v3 (cpl, csl, bm, th, thE, K0; ready-made rows of genec): asynchronous 4-byte
        // copies straight into the group's constant buffer, no register (a register prefetch spills at this budget and the
        // spill store would wait for the load)
        auto fetch_gc = [&](int n0, float* dstbuf) {
            if (et < 6 * BN) {
                const int k = et / BN, c = et - k * BN, g = n0 + c;
                if (g < p.G) {
                    const float* src = k == 0 ? p.genec + GC_CPL * G + g : k == 1 ? p.genec + GC_CSL * G + g : k == 2 ? p.bm + g
                                     : k == 3 ? p.genec + GC_THETA * G + g : k == 4 ? p.genec + GC_THE * G + g : p.genec + GC_K0 * G + g;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dstbuf + et)), "l"(src) : "memory");
                } else {
                    dstbuf[et] = (k == 3 || k == 4) ? 1.0f : 0.0f;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int r_cur = -1, m = 0;
        bool mok = false;
        float Rpl = 0.0f, Rsl = 0.0f;
        long xrow = 0;
        if (eg < nu) {  // prologue: first unit of the group
            const int u = rg.u0 + eg, r = u / p.nG;
            row_setup(r, mok, Rpl, Rsl, xrow, m);
            r_cur = r;
            fetch_gc((u - r * p.nG) * BN, s_gc + (eg * 2) * NGC * BN);
        }
        for (int t = eg; t < nu; t += 2) {
            const int u = rg.u0 + t, r = u / p.nG, n = t >> 1;
            const int n0 = (u - r * p.nG) * BN;
            if (r != r_cur) {  // new row tile
                row_setup(r, mok, Rpl, Rsl, xrow, m);
                r_cur = r;
            }
#ifdef PTC_TRACE  // diagnostic build (tools/ptc_trace.py): per-unit time stamps
            long long* tr = (p.trace && et == 0) ? p.trace + (((long)blockIdx.x * 2 + eg) * 16 + (n & 15)) * 4 : nullptr;
            if (tr) tr[0] = gtime();
#endif
            float* gc = s_gc + (eg * 2 + (n & 1)) * NGC * BN;  // double buffered: one group barrier per unit
            asm volatile("cp.async.wait_group 0;" ::: "memory");  // this unit's constants have landed (issued one unit ago)
            uint32_t cw[8];  // this cell's 16 raw counts: the line was prefetched into L1 / L2 two units ago
            if (SRC == SPV_SRC_U16_LOG1P) load_counts16(cw, reinterpret_cast<const unsigned short*>(p.X), xrow, n0, p.G, mok);
            group_barrier(eg);
            // prefetch for unit t + 2 (skipped for the counts when it starts a new row tile: once per range at most)
            if (t + 2 < nu) {
                const int u2 = u + 2, r2 = u2 / p.nG;
                fetch_gc((u2 - r2 * p.nG) * BN, s_gc + (eg * 2 + ((n + 1) & 1)) * NGC * BN);
                if (SRC == SPV_SRC_U16_LOG1P && mok && r2 == r) {
                    const unsigned short* nx = reinterpret_cast<const unsigned short*>(p.X) + xrow + (n0 + 2 * BN);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
                }
            }
#ifdef PTC_TRACE
            if (tr) tr[1] = gtime();
#endif
            tc::mbar_wait(&bar->acc_full[eg], n & 1);
            tc::fence_after_sync();
#ifdef PTC_TRACE
            if (tr) tr[2] = gtime();
#endif
            float sll = 0.0f, sep = 0.0f, ses = 0.0f;
#pragma unroll 1
            for (int j4 = 0; j4 < BN; j4 += 4) {  // not unrolled (the 16-column body overflows the instruction cache); the
                                                  // count registers are rotated so that this round's four counts are cw[0..1]
                uint32_t rpi[4], rlp[4], rls[4];
                tc::tmem_ld4(lane_addr + (uint32_t)j4, rpi);
                tc::tmem_ld4(lane_addr + (uint32_t)(BN + j4), rlp);
                tc::tmem_ld4(lane_addr + (uint32_t)(2 * BN + j4), rls);
                tc::tmem_ld_wait();
                if (j4 == BN - 4) {  // accumulator buffer fully read: hand it back to the MMA warp
                    tc::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar->acc_empty[eg]);
                }
                // branch-free over the four columns, so that their dependency chains interleave; elements that need the
                // slow forms (count beyond the table, rho below ~1e-6) are flagged and redone afterwards (rare)
                float pv[4];
                bool redo[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int gl = j4 + jj;
                    const bool ok = mok && n0 + gl < p.G;
                    NbGene ge;
                    ge.cpl = gc[0 * BN + gl]; ge.csl = gc[1 * BN + gl]; ge.bm = gc[2 * BN + gl];
                    ge.th = gc[3 * BN + gl]; ge.thE = gc[4 * BN + gl]; ge.K = gc[5 * BN + gl];
                    pv[jj] = __uint_as_float(rpi[jj]) + ge.bm;
                    float2 tl;
                    bool rare = false;
                    if (SRC == SPV_SRC_U16_LOG1P) {
                        const uint32_t w = cw[jj >> 1];
                        const uint32_t c = (jj & 1) ? (w >> 16) : (w & 0xffffu);
                        tl = s_lut[c < (uint32_t)LUT_N ? c : 0u];
                        rare = c >= (uint32_t)LUT_N;
                    } else {
                        tl.x = ok ? load_src<SRC>(p.X, xrow + n0 + gl) : 0.0f;
                        tl.y = lgamma_pos_fast(tl.x + 1.0f);
                    }
                    const NbOut o = nb_forward_v3<false>(tl.x, tl.y, __uint_as_float(rlp[jj]), __uint_as_float(rls[jj]),
                                                         __uint_as_float(rpi[jj]), ge, Rpl, Rsl, rare);
                    const bool use = ok && !rare;
                    sll += use ? o.ll : 0.0f;
                    sep += use ? o.ep : 0.0f;
                    ses += use ? o.es : 0.0f;
                    redo[jj] = ok && rare;
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if (redo[jj]) {  // rare: out-of-line slow form
                        const int gl = j4 + jj;
                        float tv = 0.0f;
                        uint32_t c = 0u;
                        if (SRC == SPV_SRC_U16_LOG1P) {
                            const uint32_t w = cw[jj >> 1];
                            c = (jj & 1) ? (w >> 16) : (w & 0xffffu);
                        } else {
                            tv = load_src<SRC>(p.X, xrow + n0 + gl);
                        }
                        NbGene ge;
                        ge.cpl = gc[0 * BN + gl]; ge.csl = gc[1 * BN + gl]; ge.bm = gc[2 * BN + gl];
                        ge.th = gc[3 * BN + gl]; ge.thE = gc[4 * BN + gl]; ge.K = gc[5 * BN + gl];
                        const NbOut o = fwd_slow(SRC == SPV_SRC_U16_LOG1P, c, tv, __uint_as_float(rlp[jj]), __uint_as_float(rls[jj]),
                                                 __uint_as_float(rpi[jj]), ge, Rpl, Rsl);
                        sll += o.ll; sep += o.ep; ses += o.es;
                    }
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) cw[i] = cw[i + 2];
                if (!mok) continue;
                if (p.pi) {
                    const int g = n0 + j4;
                    float* dst = p.pi + (long)m * G + g;
                    if (vec_pi && g + 3 < p.G) *reinterpret_cast<float4*>(dst) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    else
                        for (int jj = 0; jj < 4; ++jj)
                            if (g + jj < p.G) dst[jj] = pv[jj];
                }
            }
            if (r == r_first) { acc_seg[0][0] += sll; acc_seg[0][1] += sep; acc_seg[0][2] += ses; }
            else { acc_seg[1][0] += sll; acc_seg[1][1] += sep; acc_seg[1][2] += ses; }
        }
#ifdef PTC_TRACE
        if (p.trace && et == 0) p.trace[(((long)blockIdx.x * 2 + eg) * 16 + 15) * 4 + 3] = gtime();
#endif
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            float* o = p.part + ((((long)blockIdx.x * 2 + sg) * 2 + eg) * BM + rloc) * 3;
            o[0] = acc_seg[sg][0]; o[1] = acc_seg[sg][1]; o[2] = acc_seg[sg][2];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// rec[b] = - sum of the row partials of every (CTA, group) whose range touches b's row tile; rowc[b, 2:4] = the two
// softmax-backward row sums.  One warp per row, lanes over the CTAs (fixed order: deterministic).
__global__ void rownb_ptc_kernel(const float* __restrict__ part, int n_ctas, int units, int nG, int B, float* __restrict__ rowc,
                                 float* __restrict__ rec) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const int r = b / BM, rloc = b - r * BM;
    float ll = 0.0f, dp = 0.0f, ds = 0.0f;
    for (int c = lane; c < n_ctas; c += 32) {
        const Range rg = cta_range(c, n_ctas, units);
        if (rg.u1 <= rg.u0) continue;
        const int r0 = rg.u0 / nG, r1 = (rg.u1 - 1) / nG;
        if (r < r0 || r > r1) continue;
        const float* o = part + ((((long)c * 2 + (r - r0)) * 2) * BM + rloc) * 3;
        ll += o[0] + o[BM * 3 + 0];
        dp += o[1] + o[BM * 3 + 1];
        ds += o[2] + o[BM * 3 + 2];
    }
    ll = warp_sum(ll); dp = warp_sum(dp); ds = warp_sum(ds);
    if (lane == 0) {
        rec[b] = -ll;  // reference :823-824
        rowc[(long)b * 4 + 2] = dp;
        rowc[(long)b * 4 + 3] = ds;
    }
}

}  // namespace

extern "C" int spv_debug_ptc_occupancy() {  // diagnostic: resident CTAs per SM of the persistent forward kernel
    cudaFuncSetAttribute(nb_ptc_fwd_kernel<SPV_SRC_U16_LOG1P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    int n = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, nb_ptc_fwd_kernel<SPV_SRC_U16_LOG1P>, THREADS, SMEM_BYTES);
    return n * 1000000 + SMEM_BYTES;
}

int fwd_launch(int src, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapZ, const FwdParams& p,
               cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(nb_ptc_fwd_kernel<SPV_SRC_U16_LOG1P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(nb_ptc_fwd_kernel<SPV_SRC_F32_LOG1P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
            return SPV_ERR_LAUNCH;
        configured = true;
    }
    const int n_ctas = p.units < sm_count() ? p.units : sm_count();
    FwdParams q = p;
    q.trace = g_trace;
    if (src == SPV_SRC_U16_LOG1P) nb_ptc_fwd_kernel<SPV_SRC_U16_LOG1P><<<n_ctas, THREADS, SMEM_BYTES, st>>>(mapA, mapB, mapZ, q);
    else if (src == SPV_SRC_F32_LOG1P) nb_ptc_fwd_kernel<SPV_SRC_F32_LOG1P><<<n_ctas, THREADS, SMEM_BYTES, st>>>(mapA, mapB, mapZ, q);
    else return SPV_ERR_ARG;
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

int fwd_rowreduce(const float* part, int G, int B, float* rowc, float* rec, cudaStream_t st) {
    const int nG = (G + BN - 1) / BN, units = ((B + BM - 1) / BM) * nG;
    const int n_ctas = units < sm_count() ? units : sm_count();
    rownb_ptc_kernel<<<(B + 7) / 8, 256, 0, st>>>(part, n_ctas, units, nG, B, rowc, rec);
    SPV_CHECK_LAUNCH();
    return SPV_OK;
}

}  // namespace ptc
