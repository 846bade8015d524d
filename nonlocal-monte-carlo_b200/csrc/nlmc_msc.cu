// nlmc_msc.cu -- K2 production path: multi-spin-coded heat-bath sweeps, bit-sliced energies (K4') and
// replica-exchange swaps (K6) for +-J instances with h = 0 and degrees <= 6 (periodic and open 2D/3D lattices,
// Chimera-like graphs; the EA configs C2/C4/C5 of BASELINE.json).
//
// Layout.  A "ladder" is one NPT run (one replica per inverse temperature).  32 independent ladders
// share a 32-bit word, one bit each (bit = 1 <=> spin +1); word w = b*G + g (slot b, ladder group g,
// G = n_ladders/32, n_ladders a multiple of 128 so that four consecutive words -- a quad -- share one slot).
// The state is stored quad-major and in colour order, uint4 S[W/4][n] indexed by (quad, position of the site in
// the colour-sorted site list): a warp of the sweep kernel owns 32 consecutive positions of one quad (see MscDev).
// For C5 (L=64, 32 betas x 128 ladders = 4096 replicas) the whole state is 134 MB.  The C ABI exchanges the
// packed state site-major, packed[site][W]; the transposition runs on the device.
//
// Update rule (same distribution as the reference's sign(tanh(beta*x) - 2u + 1), NMC/nmc.py:87):
// with c = number of neighbours with J_ij*s_j = +1 and field f = 2c - 6 (even degree) or 2c - 5 (odd degree),
//     s_i <- [f > 0] XOR g,   g ~ Bernoulli(q(|f|)),  q(0) = 1/2,  q(a) = 1/(1 + exp(2*beta*a)).
// c is formed for 32 ladders at once with bit-sliced full adders (LOP3), and g is drawn for 32
// ladders at once by a bit-serial comparison of a uniform with the 32-bit threshold of each ladder's
// |f| level, most significant bit first (msc_sweep_kernel).  Random words are Philox4x32-10 keyed by
// (seed; site, (slot, ladder quad), sweep, bit step): results do not depend on the launch geometry or on how
// ladders are sharded over GPUs.  Sites are updated colour by colour (checkerboard on bipartite
// lattices, greedy colouring otherwise), all sites of a colour in parallel.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include <cooperative_groups.h>

#include "nlmc_common.cuh"
#include "nlmc_hostpar.h"

// comparison steps of the merged round (4 or 8: one or two Philox calls), see msc_sweep_kernel.  8 measured slower: C5 sweep
// 0.2932 vs 0.2859 ms, label form 0.3399 vs 0.3073, 4-slot block 0.0482 vs 0.0434 (the second Philox call and its four steps
// cost more than the shorter straggler loop saves)
#ifndef NLMC_MERGED_STEPS
#define NLMC_MERGED_STEPS 4
#endif

struct nlmc_msc {
    nlmc_instance *inst = nullptr;
    int n = 0, W = 0, n_beta = 0, n_ladders = 0, G = 0, n_colours = 0;
    int ladder_offset = 0;        // global index of this handle's first ladder (multiple of 128)
    // beta-label exchange (north_star 4): the handle owns the slots [slot_begin, slot_begin + n_beta) of a ladder of
    // n_beta_total temperatures; configurations never move, every (slot, ladder) carries the index of the beta it is
    // currently simulated at.  label_mode = 0: the classic layout (slot b is at betas[b], exchanges move bits).
    int label_mode = 0, n_beta_total = 0, slot_begin = 0;
    uint8_t *labels = nullptr;    // [n_beta_total][n_ladders] beta index held by (slot, ladder) -- replicated on every rank
    uint8_t *slot_of = nullptr;   // [n_beta_total][n_ladders] inverse: slot holding beta index i of the ladder
    uint32_t *thr_total = nullptr;  // [n_beta_total][4] thresholds of every temperature
    double *betas_total = nullptr;  // [n_beta_total]
    uint32_t *thrbits = nullptr;  // [k_steps][3][W] bit planes: bit l of word w = bit (31-p) of the level-L threshold of lane l
    uint32_t *thr_lane = nullptr; // [W][32][4] full thresholds of every lane (levels 1..3), for the stragglers
    int32_t *accepted_rounds = nullptr;  // [kRoundLog] accepted exchanges of the last rounds (slot = round % kRoundLog)
    bool own_stream = true;
    long long n_bonds = 0;
    uint32_t *S = nullptr;        // quad-major, colour-sorted: uint4 S[W/4][n] (see MscDev)
    int32_t *rec = nullptr;       // [n][8] record of the site at position p of site_list (sites sorted by colour):
                                  // positions of its 6 neighbours (-1 = padding), sign bits (bit d: J_{i,nbr d} < 0), site index
    int32_t *site_list = nullptr; // [n] sites sorted by colour
    int32_t *pos_of = nullptr;    // [n] position of a site in site_list
    uint32_t *thr_nz = nullptr;   // label mode: [W/4] steps at which the quad's threshold planes are not all zero
    uint32_t *rows = nullptr;     // scratch [n][W]: the site-major packed state of the C ABI (nlmc_msc_get/set_packed)
    std::vector<int> colour_ptr;  // [n_colours+1]
    // launch classes: the sites of a colour, split by the parity of their degree (first position, count, odd flag);
    // a launch of the sweep kernel covers one class, so the thresholds (|f| = 2,4,6 or 1,3,5) are uniform over it
    struct SiteClass { int first, count, odd; };
    std::vector<SiteClass> classes;
    bool has_odd = false;         // some site has odd degree: the label form keeps a second set of threshold tables
    uint32_t *thr = nullptr;      // [n_beta][4] thresholds of |f| = 0,2,4,6
    double *betas = nullptr;      // [n_beta]
    int32_t *E_acc = nullptr;     // [W*32] sum_i (unsatisfied bonds at i) per (word, lane)
    double *E = nullptr;          // [n_beta][n_ladders]
    uint32_t *swapmask = nullptr; // [n_beta-1][G]
    int32_t *accepted = nullptr;  // [1]
    int8_t *scratch_spins = nullptr;  // [n]
    unsigned long long seed = 0;
    uint32_t *d_counters = nullptr;  // [4] sweep counter, round counter, record slot: on the device so that captured graphs replay
    int8_t *recM = nullptr;          // grow-only record buffers of nlmc_msc_sweep_record
    double *recE = nullptr;
    size_t recM_cap = 0, recE_cap = 0;
    struct RecGraph { int ladder; bool has_M, has_E; int n_sweeps_T; cudaGraphExec_t exec; };
    std::vector<RecGraph> rec_graphs;  // one recorded sweep (sweep + unpack + energies + slot bump), replayed per sweep
    int k_steps = 5;  // unconditional bit steps of the Bernoulli comparison (tuning knob NLMC_MSC_STEPS)
    int k_merged = NLMC_MERGED_STEPS; // further steps on the four words of a thread merged into one (0 or NLMC_MERGED_STEPS; NLMC_MSC_MERGED)
    struct RoundGraph { int n_sweeps, pairs; bool with_energy_swap; cudaGraphExec_t exec; };
    std::vector<RoundGraph> graphs;  // whole rounds captured once per (n_sweeps, pairs) and replayed
    bool use_graphs = true;
    bool batch_default = false;  // sweeps of a batch as one cooperative launch (short rows)
    int batch_ctas = 0;          // co-resident CTAs of the batch kernel (0: not asked yet, -1: unavailable)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<double> h_betas;
    std::vector<uint32_t> h_thr, h_thr_odd;  // [n_beta][4] thresholds (even / odd degree): launch arguments of the sweep kernel
};

namespace nlmc {

constexpr uint32_t kTagSweep = 0x53574550u, kTagInit = 0x494e4954u, kTagSwap = 0x53574150u;

#ifndef NLMC_PHILOX_ROUNDS
#define NLMC_PHILOX_ROUNDS 10  // Philox4x32-10 (the Random123 / cuRAND default); 7 is the smallest Crush-resistant count
#endif
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < NLMC_PHILOX_ROUNDS; ++i) {
#ifdef NLMC_PHILOX_HILO
            const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = h1 ^ c1 ^ a;
            const uint32_t n2 = h0 ^ c3 ^ b;
            c1 = l1;
            c3 = l0;
#else
            const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
            const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b;
            c1 = (uint32_t)p1;
            c3 = (uint32_t)p0;
#endif
            c0 = n0;
            c2 = n2;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

constexpr int kMaxBeta = 128;

// Device view of a handle.  The state is stored QUAD-MAJOR and in COLOUR ORDER: uint4 S[quad][pos], quad = four
// consecutive words (128 ladders of one slot), pos = position of the site in the colour-sorted site list.  A warp of the
// sweep kernel owns 32 consecutive positions of one quad, so (a) every lane of a warp sits at the same slot -- thresholds,
// their bits and the per-step "all threshold bits are zero" test live in uniform registers --, (b) rows of any length
// (a block of a ladder sharded over GPUs has 16 words per site) fill warps completely, and (c) on a lattice the six
// neighbour loads and the store of a warp are contiguous 512-byte runs (consecutive sites of a colour have consecutive
// neighbours in the other colour).
struct MscDev {
    int n, W, G, n_beta, quad_offset;  // quad_offset = ladder_offset / 128: global index of ladder quad 0
    int slot_begin;       // global index of slot 0 (label mode; 0 otherwise): random streams are keyed by the GLOBAL slot
    int qpb;              // quads per slot: G / 4
    const uint32_t *thrbits;    // label mode: [kSteps][3][W]
    const uint4 *thr_lane;      // label mode: [W][32] thresholds {-, T1, T2, T3} of every lane
    const uint32_t *thr_nz;     // label mode: [W/4] bit p set <=> some threshold plane of the quad is non-zero at step p
    // label mode, sites of odd degree: the same three tables for the levels |f| = 1, 3, 5 (NULL when every degree is even)
    const uint32_t *thrbits_odd;
    const uint4 *thr_lane_odd;
    const uint32_t *thr_nz_odd;
    uint32_t *S;
    const int4 *rec;      // [n][2] by position: {nbr0..3}, {nbr4, nbr5, sign bits, site}; neighbours as POSITIONS (-1 = padding)
    const int32_t *site_list;   // [n] site at a position
    const int32_t *pos_of;      // [n] position of a site
    uint32_t seed_lo, seed_hi;
};

// thresholds of the handle's slots for the levels |f| = 2, 4, 6 as a kernel parameter: read through the constant bank
// with a uniform index, they (and every bit test on them) stay in uniform registers
// (t: |f| = 2, 4, 6 for sites of even degree; t_odd: |f| = 1, 3, 5 for sites of odd degree)
struct MscThr { uint32_t t[kMaxBeta * 3]; uint32_t t_odd[kMaxBeta * 3]; };

__device__ __forceinline__ size_t word_index(const MscDev &a, int pos, int w) {
    return ((size_t)(w >> 2) * a.n + pos) * 4 + (w & 3);
}

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

// bit-sliced population count of six one-bit inputs -> (c0, c1, c2)
__device__ __forceinline__ void count6(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                       uint32_t &c0, uint32_t &c1, uint32_t &c2) {
    const uint32_t s1 = a0 ^ a1 ^ a2, k1 = maj3(a0, a1, a2);
    const uint32_t s2 = a3 ^ a4 ^ a5, k2 = maj3(a3, a4, a5);
    c0 = s1 ^ s2;
    const uint32_t k3 = s1 & s2;
    c1 = k1 ^ k2 ^ k3;
    c2 = maj3(k1, k2, k3);
}

// Threshold bit of every lane's |f| level: t = (p1 ? I1 : 0) | (p2 ? I2 : 0) | (p3 ? I3 : 0) with p in {0, 1}.  The lane
// sets are disjoint, so the ORs are sums and the selects are products: three integer multiply-adds on the FMA pipe
// instead of select/logic operations on the ALU pipe, which is the busier one in this kernel.
__device__ __forceinline__ uint32_t level_select(uint32_t i1, uint32_t i2, uint32_t i3, uint32_t p1, uint32_t p2,
                                                 uint32_t p3) {
    uint32_t t;  // inline PTX so that the compiler does not turn the products back into selects
    asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %1, %4;\n\tmad.lo.u32 a, %2, %5, a;\n\tmad.lo.u32 %0, %3, %6, a;\n\t}"
        : "=r"(t) : "r"(i1), "r"(i2), "r"(i3), "r"(p1), "r"(p2), "r"(p3));
    return t;
}

// 0xffffffff for a negative argument, else 0 (opaque to the compiler, which would otherwise fold the mask into a predicate
// and spend a select plus a logic operation per word where one three-input logic operation does)
__device__ __forceinline__ uint32_t sign_mask(int x) {
    uint32_t m;
    asm("shr.s32 %0, %1, 31;" : "=r"(m) : "r"(x));
    return m;
}

__device__ __forceinline__ uint32_t comp(const uint4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// One colour of one sweep: a warp per (quad of words, 32 consecutive positions of the colour), a lane per site.
//
// The Bernoulli draw g ~ B(q(|f|)) of the 128 ladders of a thread is a bit-serial comparison u < T_level, most
// significant bit first.  kSteps steps run unconditionally (fully unrolled); after them a lane is still undecided with
// probability 2^-kSteps, and those few lanes are finished one by one against the remaining threshold bits with a fresh
// 32-bit uniform each (exactly the conditional probability).  The thresholds of a warp are uniform, so a step at which
// all three levels have a zero threshold bit -- 4.6 of the 6 steps on average over the ladder of config C5, all six
// above beta = 1.04 -- skips the level select: undecided lanes whose random bit is 1 are decided (u > T).
//
// kPerBit (beta-label exchange): the ladders of a word sit at DIFFERENT temperatures, so the threshold bit of a ladder
// comes from per-word bit planes thrbits[p][level][w] (rebuilt after every exchange) instead of three scalars; the planes
// of a warp's quad are the same for all its lanes (broadcast loads), thr_nz says at which steps they are all zero, and
// stragglers look their threshold up through the ladder's label.
//
// kMerged (0 or 4): after the unconditional steps a thread's four words are nearly empty of undecided lanes (2^-kSteps
// each), so they are OR-ed into ONE word -- a bit column goes to the first of the four words that is still undecided
// there -- and kMerged further comparison steps run on that word with one Philox call; the results are scattered back.
// A lane that lost its column to another word simply waits for the straggler loop.  Every random bit is still used by
// at most one lane, chosen by the past only, so the draw stays exact.
template <int kSteps, bool kPerBit, int kMerged, bool kOdd, bool kClamp = false>
// CTAs of 256 threads per SM.  5 (47-48 registers) measured 1.6-2.1 % faster than 4 (56-60 registers) on the quad-major layout
// in all three forms -- C5 sweep 0.2906 -> 0.2859 ms, label form 0.3137 -> 0.3072, 4-slot block 0.0441 -> 0.0434, identical
// trajectories (profiles/r2d_occupancy_ab.log); the label form and the odd-degree classes spill 12-24 bytes at 48 registers.
#ifndef NLMC_PERBIT_CTAS
#define NLMC_PERBIT_CTAS 5
#endif
#ifndef NLMC_SCALAR_CTAS
#define NLMC_SCALAR_CTAS 5
#endif
#ifndef NLMC_SWEEP_THREADS
#define NLMC_SWEEP_THREADS 256
#endif
__device__ __forceinline__ void msc_sweep_site(const MscDev &a, const MscThr &thr, int first, int n_sites, int idx, int b,
                                               int qin, const uint32_t *__restrict__ counters, uint32_t sweep_in_batch,
                                               int pdl) {
    constexpr bool odd = kOdd;  // the launch holds sites of odd degree (levels |f| = 1, 3, 5)
    // kClamp (the persistent batch kernel): a thread past the end recomputes the last site and skips the store, so the
    // warp stays converged inside the tile loop and the compiler keeps the uniform values on the uniform datapath
    const bool valid = idx < n_sites;
    if (kClamp) idx = min(idx, n_sites - 1);
    else if (!valid) return;
    const int qd = b * a.qpb + qin;
    const int pos = first + idx;
    const int4 r0 = __ldg(a.rec + (size_t)pos * 2), r1 = __ldg(a.rec + (size_t)pos * 2 + 1);
    const int nb[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
    const uint32_t meta = (uint32_t)r1.z;
    const int site = r1.w;
    uint4 *Sq = reinterpret_cast<uint4 *>(a.S) + (size_t)qd * a.n;
    // the record above is constant data; the spins and the sweep counter are written by the launch before this one
    if (pdl & 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t sweep = counters[0] + sweep_in_batch;
    // all six loads are issued back to back (a padding slot reads position 0 and is masked below)
    uint4 x[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) x[d] = Sq[max(nb[d], 0)];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        // x <- (x & keep) ^ flip: a neighbour keeps its bits and is inverted when J < 0; padding comes in (+1, -1) pairs
        // (plus one slot at -1 when the degree is odd; the record carries flip = 1 for the even padding slots).  One three-input logic operation
        // per word; the masks are formed with shifts so that they stay in registers (as selects on predicates the
        // compiler spent two instructions per word).
        const uint32_t keep = sign_mask(~nb[d]);
        const uint32_t flip = sign_mask((int)(meta << (31 - d)));
        x[d].x = (x[d].x & keep) ^ flip;
        x[d].y = (x[d].y & keep) ^ flip;
        x[d].z = (x[d].z & keep) ^ flip;
        x[d].w = (x[d].w & keep) ^ flip;
    }
    // Lane sets of the |f| levels 1..3, level 0 being the rest, and the sign plane.  Sites of even degree (padded to six
    // slots with neutral pairs): f = 2c - 6, levels |f| = 2, 4, 6, level 0 = zero field, sign = [c >= 4].  Sites of odd
    // degree (neutral pairs plus one slot at -1; a launch holds sites of one parity, `odd` is uniform): f = 2c - 5 with
    // c in 0..5, levels |f| = 1, 3, 5, no zero-field lanes, sign = [c >= 3].
    uint32_t I1[4], I2[4], I3[4], sgn[4], res[4], und[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t c0, c1, c2;
        count6(comp(x[0], k), comp(x[1], k), comp(x[2], k), comp(x[3], k), comp(x[4], k), comp(x[5], k), c0, c1, c2);
        if (!odd) {
            const uint32_t m0 = ~c0;                                 // |c - 3| = (m1 m0)
            const uint32_t m1 = (c2 & (c1 | c0)) | (~c2 & ~c1);
            I1[k] = ~m1 & m0;
            I2[k] = m1 & ~m0;
            I3[k] = m1 & m0;
            sgn[k] = c2;
        } else {
            I1[k] = ~c2 & c1;                                        // c in {2, 3}: |f| = 1
            I2[k] = ~c1 & (c2 ^ c0);                                 // c in {1, 4}: |f| = 3
            I3[k] = ~c1 & ~(c2 ^ c0);                                // c in {0, 5}: |f| = 5
            sgn[k] = c2 | (c1 & c0);
        }
    }
    uint32_t T1 = 0, T2 = 0, T3 = 0, nzmask;
    const uint32_t *thrbits = odd ? a.thrbits_odd : a.thrbits;
    const uint4 *thr_lane = odd ? a.thr_lane_odd : a.thr_lane;
    if (kPerBit) {
        nzmask = __ldg((odd ? a.thr_nz_odd : a.thr_nz) + qd);
    } else {
        const uint32_t *tt = odd ? thr.t_odd : thr.t;
        T1 = tt[b * 3]; T2 = tt[b * 3 + 1]; T3 = tt[b * 3 + 2];
        nzmask = __brev(T1 | T2 | T3);  // bit p <=> some level has threshold bit 31 - p set
    }
    const Philox rng{a.seed_lo, a.seed_hi ^ kTagSweep};
    // stream id = (GLOBAL slot index, GLOBAL ladder quad): independent of how ladders / beta blocks are sharded
    const uint32_t sid = ((uint32_t)(b + a.slot_begin) << 20) | (uint32_t)(a.quad_offset + qin);
    // Bit-serial comparison state per word: und = lanes whose uniform still equals the threshold on the prefix seen so
    // far; v = the uniform's bit at the last step a lane was undecided, i.e. for a decided lane the bit at its first
    // difference (v = 0 there <=> uniform < threshold <=> g = 1).  g = ~v & ~und is formed once after the last step.
    uint32_t v[4];
#pragma unroll
    for (int p = 0; p < kSteps; ++p) {
        const uint4 r4 = rng((uint32_t)site, sid, sweep, (uint32_t)p);
        if ((nzmask >> p) & 1u) {
            uint4 A1, A2, A3;
            uint32_t p1 = 0, p2 = 0, p3 = 0;
            if (kPerBit) {
                const uint32_t *tb = thrbits + (size_t)p * 3 * a.W + qd * 4;
                A1 = __ldg(reinterpret_cast<const uint4 *>(tb));
                A2 = __ldg(reinterpret_cast<const uint4 *>(tb + a.W));
                A3 = __ldg(reinterpret_cast<const uint4 *>(tb + 2 * a.W));
            } else {
                p1 = (T1 >> (31 - p)) & 1u; p2 = (T2 >> (31 - p)) & 1u; p3 = (T3 >> (31 - p)) & 1u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t r = comp(r4, k);
                const uint32_t t = kPerBit ? ((I1[k] & comp(A1, k)) | (I2[k] & comp(A2, k)) | (I3[k] & comp(A3, k)))
                                           : level_select(I1[k], I2[k], I3[k], p1, p2, p3);
                if (p == 0) {  // level-0 lanes (q = 1/2) are decided by the first bit alone (g = ~r, so v = r fits them too)
                    v[k] = r;
                    und[k] = ~(r ^ t) & (I1[k] | I2[k] | I3[k]);
                } else {
                    v[k] = (und[k] & r) | (~und[k] & v[k]);  // undecided lanes take this step's bit
                    und[k] &= ~(r ^ t);                      // still equal on this prefix
                }
            }
        } else {  // every threshold bit of this step is zero: an undecided lane with random bit 1 is decided (u > T, v = 1)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t r = comp(r4, k);
                if (p == 0) {
                    v[k] = r;
                    und[k] = ~r & (I1[k] | I2[k] | I3[k]);
                } else {
                    v[k] |= und[k];
                    und[k] &= ~r;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) res[k] = ~v[k] & ~und[k];
    uint32_t adv[4] = {0u, 0u, 0u, 0u};  // lanes that took part in the merged steps (kSteps + kMerged threshold bits consumed)
    if (kMerged > 0 && (und[0] | und[1] | und[2] | und[3])) {
        const uint32_t t01 = und[0] | und[1], t012 = t01 | und[2];
        adv[0] = und[0]; adv[1] = und[1] & ~und[0]; adv[2] = und[2] & ~t01; adv[3] = und[3] & ~t012;
        const uint32_t U = t012 | und[3];
        uint32_t um = U, vm = 0u;
        uint32_t J1 = 0u, J2 = 0u, J3 = 0u;  // level planes of the merged word (not needed when every merged step is a zero step)
        if (!kPerBit && ((nzmask >> kSteps) & ((1u << kMerged) - 1u))) {
            J1 = (adv[0] & I1[0]) | (adv[1] & I1[1]) | (adv[2] & I1[2]) | (adv[3] & I1[3]);
            J2 = (adv[0] & I2[0]) | (adv[1] & I2[1]) | (adv[2] & I2[2]) | (adv[3] & I2[3]);
            J3 = (adv[0] & I3[0]) | (adv[1] & I3[1]) | (adv[2] & I3[2]) | (adv[3] & I3[3]);
        }
        const uint4 r4 = rng((uint32_t)site, sid, sweep, 16u);
        uint4 r4b = r4;
        if (kMerged > 4) r4b = rng((uint32_t)site, sid, sweep, 17u);   // steps 5..8 of the merged round
#pragma unroll
        for (int q = 0; q < kMerged; ++q) {
            const int p = kSteps + q;
            const uint32_t r = comp(q < 4 ? r4 : r4b, q & 3);
            if ((nzmask >> p) & 1u) {
                uint32_t t;
                if (kPerBit) {
                    const uint32_t *tb = thrbits + (size_t)p * 3 * a.W + qd * 4;
                    const uint4 A1 = __ldg(reinterpret_cast<const uint4 *>(tb));
                    const uint4 A2 = __ldg(reinterpret_cast<const uint4 *>(tb + a.W));
                    const uint4 A3 = __ldg(reinterpret_cast<const uint4 *>(tb + 2 * a.W));
                    t = 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        t |= adv[k] & ((I1[k] & comp(A1, k)) | (I2[k] & comp(A2, k)) | (I3[k] & comp(A3, k)));
                } else {
                    t = level_select(J1, J2, J3, (T1 >> (31 - p)) & 1u, (T2 >> (31 - p)) & 1u, (T3 >> (31 - p)) & 1u);
                }
                vm = (um & r) | (~um & vm);
                um &= ~(r ^ t);
            } else {
                vm |= um;
                um &= ~r;
            }
        }
        const uint32_t gm = U & ~um & ~vm;  // decided in the merged steps with uniform < threshold
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            res[k] |= adv[k] & gm;
            und[k] &= ~adv[k] | um;
        }
    }
    // stragglers: each gets a fresh 32-bit uniform against the remaining threshold bits.  One Philox call
    // serves the lowest undecided lane of each of the four words.  (A select-only body measured 2 % slower than this
    // one, whose branches skip the words without a straggler.)
    if (und[0] | und[1] | und[2] | und[3]) {
        const uint32_t R1 = T1 << kSteps, R2 = T2 << kSteps, R3 = T3 << kSteps;
        const uint32_t Q1 = T1 << (kSteps + kMerged), Q2 = T2 << (kSteps + kMerged), Q3 = T3 << (kSteps + kMerged);
        uint32_t call = 32u;
        do {
            const uint4 r4 = rng((uint32_t)site, sid, sweep, call++);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t bit = und[k] & (0u - und[k]);  // 0 when the word has no straggler
                uint32_t rem;
                if (kPerBit) {
                    rem = 0u;
                    if (bit) {
                        const uint4 tl = __ldg(thr_lane + (size_t)(qd * 4 + k) * 32 + (__ffs((int)bit) - 1));
                        rem = (I3[k] & bit) ? tl.w : (I2[k] & bit) ? tl.z : tl.y;
                        rem <<= (kMerged > 0 && (adv[k] & bit)) ? kSteps + kMerged : kSteps;
                    }
                } else if (kMerged > 0 && (adv[k] & bit)) {
                    rem = (I3[k] & bit) ? Q3 : (I2[k] & bit) ? Q2 : Q1;
                } else {
                    rem = (I3[k] & bit) ? R3 : (I2[k] & bit) ? R2 : R1;  // level 0 never gets here
                }
                if (comp(r4, k) < rem) res[k] |= bit;
                und[k] ^= bit;
            }
        } while (und[0] | und[1] | und[2] | und[3]);
    }
    uint4 out;
    out.x = sgn[0] ^ res[0];
    out.y = sgn[1] ^ res[1];
    out.z = sgn[2] ^ res[2];
    out.w = sgn[3] ^ res[3];
    if (!kClamp || valid) Sq[pos] = out;
}

template <int kSteps, bool kPerBit, int kMerged, bool kOdd>
__global__ void __launch_bounds__(NLMC_SWEEP_THREADS, (kPerBit ? NLMC_PERBIT_CTAS : NLMC_SCALAR_CTAS) * (256 / NLMC_SWEEP_THREADS))
msc_sweep_kernel(const MscDev a, const MscThr thr, int first, int n_sites, const uint32_t *__restrict__ counters,
                 uint32_t sweep_in_batch, int pdl) {
    // programmatic dependent launch (pdl): the next colour's grid may be scheduled while this one drains
    if (pdl & 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int idx = (int)(blockIdx.x * (unsigned)NLMC_SWEEP_THREADS + threadIdx.x);  // position within the colour
    // slot and quad within the slot come from the block index: uniform over the CTA
    msc_sweep_site<kSteps, kPerBit, kMerged, kOdd>(a, thr, first, n_sites, idx, (int)blockIdx.y, (int)blockIdx.z, counters,
                                                   sweep_in_batch, pdl);
}

// A whole batch of sweeps in ONE cooperative launch (aimed at blocks of a ladder sharded over GPUs: a colour launch of a
// 4-slot block is 20 us of work plus launch gap and drain).  The grid is persistent -- as many CTAs as are co-resident --,
// a CTA takes the tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of a launch class (tile = 256 positions of one quad),
// and a grid barrier (1.5 us on a B200, tools/micro/gridsync.cu) stands where the launch boundary was.  MEASURED SLOWER
// than the launch chain (0.41 vs 0.29 ms per C5 sweep, 0.071 vs 0.044 ms on a 4-slot block): a quarter of the warp time
// waits at the barrier (static tiles, SMs of unequal speed) and the compiler leaves the uniform datapath inside the tile
// loop (+21 % instructions) -- profiles/r2d_batch_kernel_summary.md.  Off unless NLMC_MSC_BATCH=1; trajectories are
// identical to the launch chain's.
constexpr int kMaxClasses = 16;
struct MscClasses { int n; int first[kMaxClasses], count[kMaxClasses], odd[kMaxClasses]; };

#ifndef NLMC_BATCH_CTAS
#define NLMC_BATCH_CTAS 4
#endif
template <int kSteps, bool kPerBit, int kMerged>
__global__ void __launch_bounds__(NLMC_SWEEP_THREADS, NLMC_BATCH_CTAS * (256 / NLMC_SWEEP_THREADS))
msc_sweep_batch_kernel(const MscDev a, const MscThr thr, const MscClasses cl, int n_sweeps, uint32_t *counters) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int s = 0; s < n_sweeps; ++s) {
        for (int c = 0; c < cl.n; ++c) {
            const int first = cl.first[c], cnt = cl.count[c];
            const int nx = (cnt + NLMC_SWEEP_THREADS - 1) / NLMC_SWEEP_THREADS;
            const int n_tiles = nx * a.n_beta * a.qpb;
            for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
                const int xb = t % nx, rest = t / nx;
                const int b = rest % a.n_beta, qin = rest / a.n_beta;
                const int idx = xb * NLMC_SWEEP_THREADS + (int)threadIdx.x;
                if (cl.odd[c]) msc_sweep_site<kSteps, kPerBit, kMerged, true, true>(a, thr, first, cnt, idx, b, qin, counters, (uint32_t)s, 0);
                else msc_sweep_site<kSteps, kPerBit, kMerged, false, true>(a, thr, first, cnt, idx, b, qin, counters, (uint32_t)s, 0);
            }
            grid.sync();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] += (uint32_t)n_sweeps;   // after the last barrier: nobody reads it any more
}

// Uniform random initial spins (the production counterpart of sign(2*rand-1), NPT/npt.py:612).
__global__ void msc_init_kernel(MscDev a, uint32_t stream_id) {
    const size_t quad = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // (quad of words, position)
    if (quad >= (size_t)a.n * (a.W / 4)) return;
    const int qd = (int)(quad / a.n), pos = (int)(quad % a.n);
    const int b = qd / a.qpb, qin = qd % a.qpb;
    const uint32_t sid = ((uint32_t)(b + a.slot_begin) << 20) | (uint32_t)(a.quad_offset + qin);
    const Philox rng{a.seed_lo, a.seed_hi ^ kTagInit};
    reinterpret_cast<uint4 *>(a.S)[quad] = rng((uint32_t)__ldg(a.site_list + pos), sid, stream_id, 0u);
}

// K4': per (word, ladder) sum over sites of the number of unsatisfied bonds at the site, with bit-sliced
// vertical counters (10 bit planes) flushed every kEnergyChunk sites.  E = sum_i unsat_i - n_bonds.
constexpr int kEnergyChunk = 128;  // at most 128 sites per lane: 128 sites * 6 bonds < 2^10
// On a two-colourable graph every bond joins the two colour classes, so the sites of ONE class see every bond exactly
// once: the launcher then passes only that class (the first n_list positions) and the finish kernel doubles the sum.
// Work item = (quad of words, 32 * chunk consecutive positions); a warp takes one item, a lane the positions
// base + lane + 32 k (coalesced 512-byte rows of the quad-major layout).  A CTA loops over items with a grid stride and
// gathers its counts in SHARED memory (native shared atomics), so that every (word, ladder) address of E_acc receives one
// global atomic per CTA instead of one per item.
__global__ void __launch_bounds__(512) msc_energy_kernel(MscDev a, int32_t *E_acc, int chunk, int n_list, int use_smem) {
    extern __shared__ int32_t acc_s[];  // [W * 32] when use_smem
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (use_smem) {
        for (int i = tid; i < a.W * 32; i += (int)blockDim.x) acc_s[i] = 0;
        __syncthreads();
    }
    const int quads = a.W >> 2;
    const int span = 32 * chunk;
    const int site_chunks = (n_list + span - 1) / span;
    const long long n_items = (long long)site_chunks * quads;
    const int wpc = (int)blockDim.x >> 5;  // warps per CTA
    for (long long g = (long long)blockIdx.x * wpc + warp; g < n_items; g += (long long)gridDim.x * wpc) {
        const int qd = (int)(g % quads);
        const int s_begin = (int)(g / quads) * span, s_end = min(n_list, s_begin + span);
        const uint4 *Sq = reinterpret_cast<const uint4 *>(a.S) + (size_t)qd * a.n;
        uint32_t v[4][10];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int bb = 0; bb < 10; ++bb) v[k][bb] = 0u;
        for (int s = s_begin + lane; s < s_end; s += 32) {
            const int4 r0 = __ldg(a.rec + (size_t)s * 2), r1 = __ldg(a.rec + (size_t)s * 2 + 1);
            const int nb[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
            const uint32_t meta = (uint32_t)r1.z;
            const uint4 own = Sq[s];
            uint4 x[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) x[d] = Sq[max(nb[d], 0)];
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                // unsatisfied <=> J*s_i*s_j = -1 <=> s_i xor s_j xor [J<0]; padding contributes nothing
                const uint32_t keep = sign_mask(~nb[d]);
                const uint32_t neg = sign_mask((int)(meta << (31 - d)));
                x[d].x = (x[d].x ^ own.x ^ neg) & keep; x[d].y = (x[d].y ^ own.y ^ neg) & keep;
                x[d].z = (x[d].z ^ own.z ^ neg) & keep; x[d].w = (x[d].w ^ own.w ^ neg) & keep;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t c0, c1, c2;
                count6(comp(x[0], k), comp(x[1], k), comp(x[2], k), comp(x[3], k), comp(x[4], k), comp(x[5], k), c0, c1, c2);
                // ripple-add the 3-bit count into the 10-bit vertical counter
                uint32_t carry = v[k][0] & c0;
                v[k][0] ^= c0;
                uint32_t t = v[k][1] ^ c1 ^ carry;
                carry = maj3(v[k][1], c1, carry);
                v[k][1] = t;
                t = v[k][2] ^ c2 ^ carry;
                carry = maj3(v[k][2], c2, carry);
                v[k][2] = t;
#pragma unroll
                for (int bb = 3; bb < 10; ++bb) {
                    t = v[k][bb] & carry;
                    v[k][bb] ^= carry;
                    carry = t;
                }
            }
        }
        // the 32 lanes of the warp hold partial counts of the SAME 128 (word, ladder) addresses: lane l starts at ladder
        // l, so that the atomics of one instruction go to 32 different addresses
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            for (int i = 0; i < 32; ++i) {
                const int l = (i + lane) & 31;
                int val = 0;
#pragma unroll
                for (int bb = 0; bb < 10; ++bb) val |= (int)((v[k][bb] >> l) & 1u) << bb;
                if (val) {
                    if (use_smem) atomicAdd(acc_s + (qd * 4 + k) * 32 + l, val);
                    else atomicAdd(E_acc + (size_t)(qd * 4 + k) * 32 + l, val);
                }
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = tid; i < a.W * 32; i += (int)blockDim.x)
            if (acc_s[i]) atomicAdd(E_acc + i, acc_s[i]);
    }
}

__global__ void msc_energy_finish_kernel(int W, int G, int n_ladders, long long n_bonds, int factor, const int32_t *E_acc,
                                         double *E) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (word, lane)
    if (idx >= W * 32) return;
    const int w = idx >> 5, l = idx & 31;
    const int b = w / G, g = w % G;
    E[(size_t)b * n_ladders + g * 32 + l] = (double)factor * (double)E_acc[idx] - (double)n_bonds;
}

// Pair selection of the exchange (NPT/npt.py:514-533): the pairs still selectable are kept as a bit mask (bit i <=> pair
// (i, i+1)); the pick-th set bit is found with popcounts and __fns instead of a scan over a byte array in local memory.
struct PairSet {
    uint32_t m[kMaxBeta / 32];
    int n_avail;
    __device__ __forceinline__ void init(int n_pairs) {
#pragma unroll
        for (int w = 0; w < kMaxBeta / 32; ++w) {
            const int lo = w * 32;
            m[w] = n_pairs >= lo + 32 ? 0xffffffffu : (n_pairs > lo ? (1u << (n_pairs - lo)) - 1u : 0u);
        }
        n_avail = n_pairs;
    }
    // index of the pick-th (0-based) selectable pair; removes it and its two overlapping neighbours
    __device__ __forceinline__ int take(int pick) {
        int i = 0;
#pragma unroll
        for (int w = 0; w < kMaxBeta / 32; ++w) {
            const int c = __popc(m[w]);
            if (pick >= 0 && pick < c) { i = w * 32 + (int)__fns(m[w], 0u, pick + 1); pick = -1; }
            else if (pick >= 0) pick -= c;
        }
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
            const int q = i + j;
            if (q >= 0 && q < kMaxBeta) {
                const uint32_t bit = 1u << (q & 31);
#pragma unroll
                for (int w = 0; w < kMaxBeta / 32; ++w)
                    if (w == (q >> 5) && (m[w] & bit)) { m[w] &= ~bit; --n_avail; }
            }
        }
        return i;
    }
};

// K6: replica exchange, one thread per ladder.  Pair selection and acceptance follow the reference
// (NPT/npt.py:514-533,652-680): num_pairs non-overlapping adjacent pairs drawn one after the other
// uniformly from the pairs still available; accept with min(1, exp((b_next-b_sel)*(E_next-E_sel))).
// Accepted exchanges are recorded as lane masks per temperature boundary and applied to the
// configurations by msc_swap_apply_kernel (the reference swaps configurations, not labels).
constexpr int kRoundLog = 4096;  // per-round acceptance counts kept for the last kRoundLog rounds
__global__ void msc_swap_decide_kernel(int n_beta, int n_ladders, int G, int num_pairs, const double *betas, double *E,
                                       uint32_t *swapmask, int32_t *accepted, int32_t *accepted_rounds, uint32_t seed_lo,
                                       uint32_t seed_hi, const uint32_t *__restrict__ counters, int ladder_offset) {
    const uint32_t round = counters[1];
    const int ladder = blockIdx.x * blockDim.x + threadIdx.x;
    if (ladder >= n_ladders) return;
    const Philox rng{seed_lo, seed_hi ^ kTagSwap};
    PairSet avail;
    avail.init(n_beta - 1);
    int acc = 0;
    for (int k = 0; k < num_pairs && avail.n_avail > 0; ++k) {
        const uint4 r = rng((uint32_t)(ladder + ladder_offset), round, (uint32_t)k, 0u);
        const int i = avail.take((int)(((unsigned long long)r.x * (unsigned)avail.n_avail) >> 32));
        const double E_sel = E[(size_t)i * n_ladders + ladder], E_next = E[(size_t)(i + 1) * n_ladders + ladder];
        const double x = (betas[i + 1] - betas[i]) * (E_next - E_sel);
        const double u = ((double)r.y * 4294967296.0 + (double)r.z + 0.5) * (1.0 / 18446744073709551616.0);
        if (u < fmin(1.0, exp(x))) {
            ++acc;
            atomicOr(swapmask + (size_t)i * G + (ladder >> 5), 1u << (ladder & 31));
            E[(size_t)i * n_ladders + ladder] = E_next;
            E[(size_t)(i + 1) * n_ladders + ladder] = E_sel;
        }
    }
    if (acc) {
        atomicAdd(accepted, acc);
        atomicAdd(accepted_rounds + (round % kRoundLog), acc);
    }
}

__global__ void msc_swap_apply_kernel(MscDev a, const uint32_t *swapmask) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // (position, g)
    if (idx >= (size_t)a.n * a.G) return;
    const int g = (int)(idx % a.G), pos = (int)(idx / a.G);
    uint32_t A = a.S[word_index(a, pos, g)];
    bool A_dirty = false;
    for (int b = 0; b + 1 < a.n_beta; ++b) {
        uint32_t B = a.S[word_index(a, pos, (b + 1) * a.G + g)];
        bool B_dirty = false;
        const uint32_t mask = __ldg(swapmask + (size_t)b * a.G + g);
        if (mask) {
            const uint32_t t = (A ^ B) & mask;  // lanes whose two configurations differ at this site
            if (t) {
                A ^= t;
                B ^= t;
                A_dirty = B_dirty = true;
            }
        }
        if (A_dirty) a.S[word_index(a, pos, b * a.G + g)] = A;
        A = B;
        A_dirty = B_dirty;
    }
    if (A_dirty) a.S[word_index(a, pos, (a.n_beta - 1) * a.G + g)] = A;
}

// K6, beta-label form (north_star 4): configurations stay in their slots, every (slot, ladder) carries the
// index of the temperature it is simulated at.  E[slot][ladder] holds the energies of ALL slots of the ladder set
// (gathered over the ranks when the beta range is sharded); every rank runs this kernel on identical inputs and
// arrives at the identical permutation -- only 8 bytes per replica ever cross the GPUs.  Pair selection and
// acceptance as in msc_swap_decide_kernel (NPT/npt.py:514-533,652-680), pairs being adjacent TEMPERATURES.
__device__ __forceinline__ void label_swap_ladder(int ladder, uint32_t round, int n_beta, int n_ladders, int num_pairs,
                                                  const double *betas, const double *E, uint8_t *labels, uint8_t *slot_of,
                                                  int32_t *accepted, int32_t *accepted_rounds, uint32_t seed_lo,
                                                  uint32_t seed_hi, int ladder_offset) {
    const Philox rng{seed_lo, seed_hi ^ kTagSwap};
    PairSet avail;
    avail.init(n_beta - 1);
    int acc = 0;
    for (int k = 0; k < num_pairs && avail.n_avail > 0; ++k) {
        const uint4 r = rng((uint32_t)(ladder + ladder_offset), round, (uint32_t)k, 0u);
        const int i = avail.take((int)(((unsigned long long)r.x * (unsigned)avail.n_avail) >> 32));
        const int sa = slot_of[(size_t)i * n_ladders + ladder], sb = slot_of[(size_t)(i + 1) * n_ladders + ladder];
        const double E_sel = E[(size_t)sa * n_ladders + ladder], E_next = E[(size_t)sb * n_ladders + ladder];
        const double x = (betas[i + 1] - betas[i]) * (E_next - E_sel);
        const double u = ((double)r.y * 4294967296.0 + (double)r.z + 0.5) * (1.0 / 18446744073709551616.0);
        if (u < fmin(1.0, exp(x))) {
            ++acc;
            labels[(size_t)sa * n_ladders + ladder] = (uint8_t)(i + 1);
            labels[(size_t)sb * n_ladders + ladder] = (uint8_t)i;
            slot_of[(size_t)i * n_ladders + ladder] = (uint8_t)sb;
            slot_of[(size_t)(i + 1) * n_ladders + ladder] = (uint8_t)sa;
        }
    }
    if (acc) {
        atomicAdd(accepted, acc);
        atomicAdd(accepted_rounds + (round % kRoundLog), acc);
    }
}

__global__ void msc_label_swap_kernel(int n_beta, int n_ladders, int num_pairs, const double *betas, const double *E,
                                      uint8_t *labels, uint8_t *slot_of, int32_t *accepted, int32_t *accepted_rounds,
                                      uint32_t seed_lo, uint32_t seed_hi, const uint32_t *__restrict__ counters,
                                      int ladder_offset) {
    const uint32_t round = counters[1];
    const int ladder = blockIdx.x * blockDim.x + threadIdx.x;
    if (ladder >= n_ladders) return;
    label_swap_ladder(ladder, round, n_beta, n_ladders, num_pairs, betas, E, labels, slot_of, accepted, accepted_rounds,
                      seed_lo, seed_hi, ladder_offset);
}

__device__ __forceinline__ void thrbits_item(int idx, int planes, int W, int G, const uint8_t *labels_local,
                                             const uint32_t *thr_total, uint32_t *thrbits, uint4 *thr_lane) {
    if (idx < W * 32) {  // full thresholds of lane (w, l)
        const int w = idx >> 5, l = idx & 31;
        const int lab = labels_local[(size_t)(w / G) * (G * 32) + (w % G) * 32 + l];
        thr_lane[idx] = *reinterpret_cast<const uint4 *>(thr_total + lab * 4);
    }
    if (idx >= planes * 3 * W) return;
    const int w = idx % W, lev = (idx / W) % 3 + 1, p = idx / (3 * W);
    const int b = w / G, g = w % G;
    const uint8_t *lab = labels_local + (size_t)b * (G * 32) + g * 32;
    uint32_t word = 0u;
    for (int l = 0; l < 32; ++l) word |= ((thr_total[lab[l] * 4 + lev] >> (31 - p)) & 1u) << l;
    thrbits[idx] = word;
}

__device__ __forceinline__ void thrnz_item(int qd, int planes, int W, const uint32_t *thrbits, uint32_t *thr_nz) {
    uint32_t nz = 0u;
    for (int p = 0; p < planes; ++p) {
        uint32_t any = 0u;
        for (int lev = 0; lev < 3; ++lev)
            for (int k = 0; k < 4; ++k) any |= thrbits[((size_t)p * 3 + lev) * W + qd * 4 + k];
        if (any) nz |= 1u << p;
    }
    thr_nz[qd] = nz;
}

// The whole label exchange of a round in ONE launch of one CTA (ladder sets of up to 1024 ladders): clear the round's
// acceptance slot, decide and permute the labels, rebuild this handle's threshold planes, bump the round counter.  As five
// tiny launches the same work took 39 us per round, which on a 4-slot block of a ladder sharded over 8 GPUs is 5 % of the
// round.
__global__ void __launch_bounds__(1024) msc_label_round_kernel(
    int n_beta, int n_ladders, int num_pairs, const double *betas, const double *E, uint8_t *labels, uint8_t *slot_of,
    int32_t *accepted, int32_t *accepted_rounds, uint32_t seed_lo, uint32_t seed_hi, uint32_t *counters, int ladder_offset,
    int planes, int W, int G, int slot_begin, const uint32_t *thr_total, uint32_t *thrbits, uint4 *thr_lane, uint32_t *thr_nz,
    int sets) {
    const uint32_t round = counters[1];
    const int tid = threadIdx.x;
    if (tid == 0) accepted_rounds[round % kRoundLog] = 0;
    __syncthreads();
    if (n_beta >= 2 && num_pairs > 0) {
        if (tid < n_ladders)
            label_swap_ladder(tid, round, n_beta, n_ladders, num_pairs, betas, E, labels, slot_of, accepted, accepted_rounds,
                              seed_lo, seed_hi, ladder_offset);
        __syncthreads();
        const int items = max(planes * 3, 32) * W;
        const uint8_t *labels_local = labels + (size_t)slot_begin * n_ladders;
        for (int idx = tid; idx < items * sets; idx += (int)blockDim.x) {
            const int set = idx / items;   // 0: even degree, 1: odd degree (tables of set 1 follow those of set 0)
            thrbits_item(idx - set * items, planes, W, G, labels_local, thr_total + (size_t)set * n_beta * 4,
                         thrbits + (size_t)set * planes * 3 * W, thr_lane + (size_t)set * W * 32);
        }
        __syncthreads();
        for (int q = tid; q < (W / 4) * sets; q += (int)blockDim.x) {
            const int set = q / (W / 4);
            thrnz_item(q - set * (W / 4), planes, W, thrbits + (size_t)set * planes * 3 * W, thr_nz + (size_t)set * (W / 4));
        }
    }
    __syncthreads();
    if (tid == 0) counters[1] = round + 1u;
}

// Threshold bit planes of this handle's slots from the labels: one thread per (step p, level, word).
__global__ void msc_thrbits_kernel(int planes, int W, int G, const uint8_t *__restrict__ labels_local,
                                   const uint32_t *__restrict__ thr_total, uint32_t *thrbits, uint4 *thr_lane) {
    thrbits_item(blockIdx.x * blockDim.x + threadIdx.x, planes, W, G, labels_local, thr_total, thrbits, thr_lane);
}

// thr_nz[quad]: bit p set <=> at step p some level's plane of one of the quad's four words is non-zero
__global__ void msc_thrnz_kernel(int planes, int W, const uint32_t *__restrict__ thrbits, uint32_t *thr_nz) {
    const int qd = blockIdx.x * blockDim.x + threadIdx.x;
    if (qd < W / 4) thrnz_item(qd, planes, W, thrbits, thr_nz);
}

__global__ void msc_labels_identity_kernel(int n_beta, int n_ladders, uint8_t *labels, uint8_t *slot_of) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_beta * n_ladders) return;
    labels[idx] = slot_of[idx] = (uint8_t)(idx / n_ladders);
}

__global__ void msc_clear_round_slot_kernel(int32_t *accepted_rounds, const uint32_t *__restrict__ counters) {
    accepted_rounds[counters[1] % kRoundLog] = 0;
}

__global__ void msc_bump_kernel(uint32_t *counters, int which, uint32_t by = 1u) { counters[which] += by; }

// all replicas (every beta) of one ladder as int8 +-1: out[b][site]
__global__ void msc_unpack_ladder_kernel(MscDev a, int g, int lane, int8_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i < a.n) out[(size_t)b * a.n + i] = ((a.S[word_index(a, __ldg(a.pos_of + i), b * a.G + g)] >> lane) & 1u) ? 1 : -1;
}

// the same into slot counters[2] of a record buffer (captured once, replayed per recorded sweep)
// n_sweeps_T == 0: slot-major record [sweep][beta][site]; n_sweeps_T > 0: [beta][site][sweep] -- the layout of the
// reference's M ((R*N) x sweeps, NPT/npt.py:640-644), so the host only widens int8 to float64
__global__ void msc_unpack_ladder_rec_kernel(MscDev a, int g, int lane, int8_t *base, size_t stride,
                                             const uint32_t *__restrict__ counters, int n_sweeps_T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= a.n) return;
    const int8_t v = ((a.S[word_index(a, __ldg(a.pos_of + i), b * a.G + g)] >> lane) & 1u) ? 1 : -1;
    if (n_sweeps_T) base[((size_t)b * a.n + i) * n_sweeps_T + counters[2]] = v;
    else base[(size_t)counters[2] * stride + (size_t)b * a.n + i] = v;
}

__global__ void msc_record_energy_kernel(const double *__restrict__ E, double *base, size_t stride,
                                         const uint32_t *__restrict__ counters) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < stride) base[(size_t)counters[2] * stride + i] = E[i];
}

__global__ void msc_unpack_kernel(MscDev a, int w, int lane, int8_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) out[i] = ((a.S[word_index(a, __ldg(a.pos_of + i), w)] >> lane) & 1u) ? 1 : -1;
}

__global__ void msc_pack_kernel(MscDev a, int w, int lane, const int8_t *in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    uint32_t *p = a.S + word_index(a, __ldg(a.pos_of + i), w);
    const uint32_t bit = 1u << lane;
    *p = in[i] > 0 ? (*p | bit) : (*p & ~bit);  // one thread per site: no other writer of this word
}

// The packed state as the C ABI exchanges it -- rows[site][W], site-major -- from / to the quad-major device layout.
__global__ void msc_to_rows_kernel(MscDev a, uint4 *rows) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // (site, quad), quad fastest
    const int quads = a.W >> 2;
    if (idx >= (size_t)a.n * quads) return;
    const int i = (int)(idx / quads), qd = (int)(idx % quads);
    rows[idx] = reinterpret_cast<const uint4 *>(a.S)[(size_t)qd * a.n + __ldg(a.pos_of + i)];
}

__global__ void msc_from_rows_kernel(MscDev a, const uint4 *rows) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int quads = a.W >> 2;
    if (idx >= (size_t)a.n * quads) return;
    const int i = (int)(idx / quads), qd = (int)(idx % quads);
    reinterpret_cast<uint4 *>(a.S)[(size_t)qd * a.n + __ldg(a.pos_of + i)] = rows[idx];
}

// 32-bit thresholds of q(|f|) = 1/(1 + exp(2*beta*|f|)) for |f| = 0, 2, 4, 6 (sites of even degree) or, with odd = true,
// for |f| = -, 1, 3, 5 (sites of odd degree; entry 0 unused)
static std::vector<uint32_t> msc_thresholds(int n_beta, const double *betas, bool odd = false) {
    std::vector<uint32_t> thr((size_t)n_beta * 4);
    for (int b = 0; b < n_beta; ++b) {
        thr[(size_t)b * 4] = 0x80000000u;
        for (int a = 1; a < 4; ++a) {
            const double f = odd ? 2.0 * a - 1.0 : 2.0 * a;
            const double q = 1.0 / (1.0 + std::exp(2.0 * betas[b] * f));
            thr[(size_t)b * 4 + a] = (uint32_t)std::min(4294967295.0, std::floor(q * 4294967296.0));
        }
    }
    return thr;
}

static MscDev dev_view(const nlmc_msc *M) {
    MscDev d;
    d.n = M->n; d.W = M->W; d.G = M->G; d.n_beta = M->n_beta; d.quad_offset = M->ladder_offset / 128;
    d.S = M->S; d.rec = reinterpret_cast<const int4 *>(M->rec); d.site_list = M->site_list; d.pos_of = M->pos_of;
    d.seed_lo = (uint32_t)M->seed; d.seed_hi = (uint32_t)(M->seed >> 32);
    d.slot_begin = M->slot_begin;
    d.qpb = M->G / 4;
    d.thrbits = M->thrbits;
    d.thr_lane = reinterpret_cast<const uint4 *>(M->thr_lane);
    d.thr_nz = M->thr_nz;
    // the tables of the odd-degree sites follow those of the even-degree ones
    d.thrbits_odd = M->thrbits ? M->thrbits + (size_t)(M->k_steps + M->k_merged) * 3 * M->W : nullptr;
    d.thr_lane_odd = M->thr_lane ? reinterpret_cast<const uint4 *>(M->thr_lane) + (size_t)M->W * 32 : nullptr;
    d.thr_nz_odd = M->thr_nz ? M->thr_nz + M->W / 4 : nullptr;
    return d;
}

static MscThr thr_view(const nlmc_msc *M) {
    MscThr t;
    for (int b = 0; b < kMaxBeta; ++b)
        for (int a = 0; a < 3; ++a) {
            const bool on = !M->label_mode && b < M->n_beta;
            t.t[b * 3 + a] = on ? M->h_thr[(size_t)b * 4 + a + 1] : 0u;
            t.t_odd[b * 3 + a] = on ? M->h_thr_odd[(size_t)b * 4 + a + 1] : 0u;
        }
    return t;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int kSteps>
static void launch_colour(const nlmc_msc *M, const MscDev &d, const MscThr &t, int first, int cnt, uint32_t sweep_in_batch,
                          bool pdl, bool trigger, int odd) {
    constexpr int kT = NLMC_SWEEP_THREADS;
    const dim3 blocks((unsigned)((cnt + kT - 1) / kT), (unsigned)M->n_beta, (unsigned)d.qpb);
    const uint32_t *ctr = M->d_counters;
    const int flag = (pdl ? 1 : 0) | (trigger ? 2 : 0);  // bit 0: wait for the launch before, bit 1: let the next one start early
    auto go = [&](auto kernel) { launch_maybe_pdl(kernel, blocks, dim3(kT), M->stream, pdl, d, t, first, cnt, ctr, sweep_in_batch, flag); };
    if (M->label_mode) {
        if (M->k_merged) { if (odd) go(msc_sweep_kernel<kSteps, true, NLMC_MERGED_STEPS, true>); else go(msc_sweep_kernel<kSteps, true, NLMC_MERGED_STEPS, false>); }
        else { if (odd) go(msc_sweep_kernel<kSteps, true, 0, true>); else go(msc_sweep_kernel<kSteps, true, 0, false>); }
    } else {
        if (M->k_merged) { if (odd) go(msc_sweep_kernel<kSteps, false, NLMC_MERGED_STEPS, true>); else go(msc_sweep_kernel<kSteps, false, NLMC_MERGED_STEPS, false>); }
        else { if (odd) go(msc_sweep_kernel<kSteps, false, 0, true>); else go(msc_sweep_kernel<kSteps, false, 0, false>); }
    }
}

// NLMC_MSC_BATCH = 1 selects the cooperative batch kernel (measured slower, see msc_sweep_batch_kernel; default off).
static bool batch_mode(const nlmc_msc *M) {
    if (M->k_steps != 5 || M->k_merged != NLMC_MERGED_STEPS) return false;
    if ((int)M->classes.size() > kMaxClasses) return false;
    if (const char *e = getenv("NLMC_MSC_BATCH")) return atoi(e) != 0;
    return M->batch_default;
}

static int launch_sweep_batch(nlmc_msc *M, const MscDev &d, const MscThr &t, int n_sweeps) {
    MscClasses cl;
    cl.n = 0;
    for (const auto &c : M->classes) {
        if (c.count == 0) continue;
        cl.first[cl.n] = c.first; cl.count[cl.n] = c.count; cl.odd[cl.n] = c.odd; ++cl.n;
    }
    if (cl.n == 0) return NLMC_OK;
    void *kernel = M->label_mode ? (void *)msc_sweep_batch_kernel<5, true, NLMC_MERGED_STEPS> : (void *)msc_sweep_batch_kernel<5, false, NLMC_MERGED_STEPS>;
    if (M->batch_ctas == 0) {   // co-resident CTAs of this kernel on this device, once per handle
        int per_sm = 0, sms = 0, dev = 0;
        NLMC_CUDA(cudaGetDevice(&dev));
        NLMC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        NLMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NLMC_SWEEP_THREADS, 0));
        M->batch_ctas = per_sm * sms;
        if (const char *e = getenv("NLMC_MSC_BATCH_CTAS")) M->batch_ctas = std::max(1, std::min(M->batch_ctas, atoi(e)));
        if (M->batch_ctas <= 0) M->batch_ctas = -1;
    }
    if (M->batch_ctas < 0) return NLMC_ERR_STATE;
    long long max_tiles = 0;
    for (int c = 0; c < cl.n; ++c)
        max_tiles = std::max(max_tiles, (long long)((cl.count[c] + NLMC_SWEEP_THREADS - 1) / NLMC_SWEEP_THREADS) * M->n_beta * d.qpb);
    const unsigned grid = (unsigned)std::max(1LL, std::min((long long)M->batch_ctas, max_tiles));
    uint32_t *ctr = M->d_counters;
    void *args[] = {(void *)&d, (void *)&t, (void *)&cl, (void *)&n_sweeps, (void *)&ctr};
    NLMC_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(NLMC_SWEEP_THREADS), args, 0, M->stream));
    return NLMC_OK;
}

static int launch_sweeps(nlmc_msc *M, int n_sweeps) {
    const MscDev d = dev_view(M);
    const MscThr t = thr_view(M);
    // The sweep index a kernel hashes into its random stream is counters[0] + its position in the batch.  Short site rows
    // (a block of a sharded ladder: 22 us per colour launch at 16 words) bump the counter once per batch; full rows keep
    // the 1-thread bump kernel after every sweep, which measured 2 % FASTER there (round 1, and again in round 2).
    // NLMC_MSC_PDL=1 chains the colour launches of a batch with programmatic dependent launch (the next grid is scheduled
    // while the previous one drains and waits for it just before its first load of spins): +1 % on back-to-back sweeps,
    // -2.5 % on whole C5 rounds (it implies the single bump), so it is off by default.
    const char *e_pdl = getenv("NLMC_MSC_PDL");
    const bool pdl = e_pdl ? atoi(e_pdl) != 0 : false;
    // A batch of sweeps as ONE cooperative launch (msc_sweep_batch_kernel): the default kernel shape only, on request
    if (n_sweeps >= 1 && !pdl && batch_mode(M)) {
        int rc = launch_sweep_batch(M, d, t, n_sweeps);
        if (rc != NLMC_ERR_STATE) return rc;   // NLMC_ERR_STATE: the grid does not fit / too many classes -> launch chain
    }
    const bool bump_once = pdl || (M->W < 128 ? !getenv("NLMC_MSC_BUMP_EACH") : getenv("NLMC_MSC_BUMP_ONCE") != nullptr);
    bool chained = false;  // the launch before this one in the stream is a sweep kernel of this batch
    for (int s = 0; s < n_sweeps; ++s) {
        const uint32_t off = bump_once ? (uint32_t)s : 0u;
        for (const auto &cl : M->classes) {
            const int first = cl.first, cnt = cl.count, odd = cl.odd;
            if (cnt == 0) continue;
            const bool p = pdl && chained;
            switch (M->k_steps) {
                case 4: launch_colour<4>(M, d, t, first, cnt, off, p, pdl, odd); break;
                case 6: launch_colour<6>(M, d, t, first, cnt, off, p, pdl, odd); break;
                case 7: launch_colour<7>(M, d, t, first, cnt, off, p, pdl, odd); break;
                case 8: launch_colour<8>(M, d, t, first, cnt, off, p, pdl, odd); break;
                default: launch_colour<5>(M, d, t, first, cnt, off, p, pdl, odd); break;
            }
            chained = true;
        }
        if (!bump_once) { msc_bump_kernel<<<1, 1, 0, M->stream>>>(M->d_counters, 0); chained = false; }
    }
    if (bump_once && n_sweeps > 0) msc_bump_kernel<<<1, 1, 0, M->stream>>>(M->d_counters, 0, (uint32_t)n_sweeps);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

static int launch_energy(nlmc_msc *M) {
    const MscDev d = dev_view(M);
    const int quads = M->W / 4;
    const bool bipartite = M->n_colours == 2;  // one colour class sees every bond once
    const int n_list = bipartite ? M->colour_ptr[1] : M->n;
    // sites per lane and item: as many as the 10-bit counters allow on big lattices (fewer flushes), fewer on small ones
    // so that there are still about two items per warp slot of the grid (measured at C5 size: 28 sites per lane, 256
    // threads, 296 CTAs = 124 us against 195 us with 128-thread CTAs x 1184)
    const long long slots = 148LL * 2 * 8 * 2;  // 2 CTAs of 8 warps per SM (the kernel needs 128 registers), 2 items per warp
    const long long per_item = ((long long)n_list * quads + slots - 1) / slots;          // positions per item
    int chunk = (int)std::max(16LL, std::min((long long)kEnergyChunk, (per_item + 31) / 32));
    if (const char *e = getenv("NLMC_ENERGY_CHUNK")) chunk = std::max(1, std::min(kEnergyChunk, atoi(e)));
    const int site_chunks = (n_list + 32 * chunk - 1) / (32 * chunk);
    const long long items = (long long)site_chunks * quads;
    int threads = 256, max_ctas = 148 * 2;
    if (const char *e = getenv("NLMC_ENERGY_THREADS")) threads = std::max(32, std::min(512, atoi(e) / 32 * 32));
    if (const char *e = getenv("NLMC_ENERGY_CTAS")) max_ctas = std::max(1, atoi(e));
    const int wpc = threads / 32;
    const unsigned grid = (unsigned)std::max(1LL, std::min((items + wpc - 1) / wpc, (long long)max_ctas));
    const size_t smem = sizeof(int32_t) * (size_t)M->W * 32;
    const int use_smem = smem <= 40 * 1024 ? 1 : 0;   // rows of up to 320 words; longer ones add straight into E_acc
    NLMC_CUDA(cudaMemsetAsync(M->E_acc, 0, sizeof(int32_t) * (size_t)M->W * 32, M->stream));
    msc_energy_kernel<<<grid, threads, use_smem ? smem : 0, M->stream>>>(d, M->E_acc, chunk, n_list, use_smem);
    msc_energy_finish_kernel<<<(M->W * 32 + 255) / 256, 256, 0, M->stream>>>(M->W, M->G, M->n_ladders, M->n_bonds,
                                                                            bipartite ? 2 : 1, M->E_acc, M->E);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

static int launch_thrbits(nlmc_msc *M) {
    const int planes = M->k_steps + M->k_merged;
    const int items = std::max(planes * 3, 32) * M->W;
    for (int set = 0; set < (M->has_odd ? 2 : 1); ++set) {   // set 1: the tables of the odd-degree sites
        uint32_t *tb = M->thrbits + (size_t)set * planes * 3 * M->W;
        msc_thrbits_kernel<<<(items + 127) / 128, 128, 0, M->stream>>>(
            planes, M->W, M->G, M->labels + (size_t)M->slot_begin * M->n_ladders, M->thr_total + (size_t)set * M->n_beta_total * 4,
            tb, reinterpret_cast<uint4 *>(M->thr_lane) + (size_t)set * M->W * 32);
        msc_thrnz_kernel<<<(M->W / 4 + 127) / 128, 128, 0, M->stream>>>(planes, M->W, tb, M->thr_nz + (size_t)set * (M->W / 4));
    }
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

// label mode: decisions on the energies of ALL slots (E_full [n_beta_total][n_ladders], device), then this handle's
// threshold planes are rebuilt from the new labels
static int launch_label_exchange(nlmc_msc *M, const double *E_full_dev, int num_pairs) {
    const MscDev d = dev_view(M);
    if (M->n_ladders <= 1024 && !getenv("NLMC_MSC_SPLIT_EXCHANGE")) {
        msc_label_round_kernel<<<1, 1024, 0, M->stream>>>(
            M->n_beta_total, M->n_ladders, num_pairs, M->betas_total, E_full_dev, M->labels, M->slot_of, M->accepted,
            M->accepted_rounds, d.seed_lo, d.seed_hi, M->d_counters, M->ladder_offset, M->k_steps + M->k_merged, M->W, M->G,
            M->slot_begin, M->thr_total, M->thrbits, reinterpret_cast<uint4 *>(M->thr_lane), M->thr_nz, M->has_odd ? 2 : 1);
        NLMC_CUDA(cudaGetLastError());
        return NLMC_OK;
    }
    msc_clear_round_slot_kernel<<<1, 1, 0, M->stream>>>(M->accepted_rounds, M->d_counters);
    if (M->n_beta_total >= 2 && num_pairs > 0) {
        msc_label_swap_kernel<<<(M->n_ladders + 127) / 128, 128, 0, M->stream>>>(
            M->n_beta_total, M->n_ladders, num_pairs, M->betas_total, E_full_dev, M->labels, M->slot_of, M->accepted,
            M->accepted_rounds, d.seed_lo, d.seed_hi, M->d_counters, M->ladder_offset);
        int rc = launch_thrbits(M);
        if (rc) return rc;
    }
    msc_bump_kernel<<<1, 1, 0, M->stream>>>(M->d_counters, 1);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

static int launch_swap(nlmc_msc *M, int num_pairs) {
    const MscDev d = dev_view(M);
    if (M->label_mode) {
        if (M->n_beta != M->n_beta_total) {
            set_error("nlmc_msc: this handle holds a block of a sharded ladder; gather the energies of all blocks and call "
                      "nlmc_msc_exchange_labels");
            return NLMC_ERR_STATE;
        }
        return launch_label_exchange(M, M->E, num_pairs);
    }
    msc_clear_round_slot_kernel<<<1, 1, 0, M->stream>>>(M->accepted_rounds, M->d_counters);
    if (M->n_beta >= 2 && num_pairs > 0) {
        NLMC_CUDA(cudaMemsetAsync(M->swapmask, 0, sizeof(uint32_t) * (size_t)(M->n_beta - 1) * M->G, M->stream));
        msc_swap_decide_kernel<<<(M->n_ladders + 127) / 128, 128, 0, M->stream>>>(
            M->n_beta, M->n_ladders, M->G, num_pairs, M->betas, M->E, M->swapmask, M->accepted, M->accepted_rounds,
            d.seed_lo, d.seed_hi, M->d_counters, M->ladder_offset);
        const size_t items = (size_t)M->n * M->G;
        msc_swap_apply_kernel<<<(unsigned)((items + 255) / 256), 256, 0, M->stream>>>(d, M->swapmask);
    }
    msc_bump_kernel<<<1, 1, 0, M->stream>>>(M->d_counters, 1);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

// sweeps [+ energies + exchange] as one graph launch; graphs are captured once per shape and cached
static int run_round(nlmc_msc *M, int n_sweeps, int num_pairs, bool with_energy_swap) {
    const long long launches = (long long)n_sweeps * ((long long)M->classes.size() + 1);
    if (!M->use_graphs || launches < 4 || launches > 4096) {  // tiny or huge rounds: plain launches
        int rc = launch_sweeps(M, n_sweeps);
        if (!rc && with_energy_swap) rc = launch_energy(M);
        if (!rc && with_energy_swap) rc = launch_swap(M, num_pairs);
        return rc;
    }
    for (auto &g : M->graphs)
        if (g.n_sweeps == n_sweeps && g.pairs == num_pairs && g.with_energy_swap == with_energy_swap) {
            NLMC_CUDA(cudaGraphLaunch(g.exec, M->stream));
            return NLMC_OK;
        }
    cudaGraph_t graph = nullptr;
    NLMC_CUDA(cudaStreamBeginCapture(M->stream, cudaStreamCaptureModeThreadLocal));
    int rc = launch_sweeps(M, n_sweeps);
    if (!rc && with_energy_swap) rc = launch_energy(M);
    if (!rc && with_energy_swap) rc = launch_swap(M, num_pairs);
    const cudaError_t e = cudaStreamEndCapture(M->stream, &graph);
    if (rc || e != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        if (!rc) set_error("nlmc_msc: stream capture failed: %s", cudaGetErrorString(e));
        return rc ? rc : NLMC_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) {
        set_error("nlmc_msc: cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
        return NLMC_ERR_CUDA;
    }
    if (M->graphs.size() >= 8) {  // keep the cache small
        cudaGraphExecDestroy(M->graphs.front().exec);
        M->graphs.erase(M->graphs.begin());
    }
    M->graphs.push_back({n_sweeps, num_pairs, with_energy_swap, exec});
    NLMC_CUDA(cudaGraphLaunch(exec, M->stream));
    return NLMC_OK;
}

static void drop_graphs(nlmc_msc *M) {
    for (auto &g : M->graphs) cudaGraphExecDestroy(g.exec);
    M->graphs.clear();
    for (auto &g : M->rec_graphs) cudaGraphExecDestroy(g.exec);
    M->rec_graphs.clear();
}

// one recorded sweep: sweep, the ladder's states and / or all energies into slot counters[2] of the record buffers
static int launch_recorded_sweep(nlmc_msc *M, int ladder, bool has_M, bool has_E, size_t m_stride, size_t e_stride,
                                 int n_sweeps_T) {
    const MscDev d = dev_view(M);
    int rc = launch_sweeps(M, 1);
    if (rc) return rc;
    if (has_M)
        msc_unpack_ladder_rec_kernel<<<dim3((unsigned)((M->n + 255) / 256), (unsigned)M->n_beta), 256, 0, M->stream>>>(
            d, ladder / 32, ladder % 32, M->recM, m_stride, M->d_counters, n_sweeps_T);
    if (has_E) {
        if ((rc = launch_energy(M))) return rc;
        msc_record_energy_kernel<<<(unsigned)((e_stride + 255) / 256), 256, 0, M->stream>>>(M->E, M->recE, e_stride,
                                                                                        M->d_counters);
    }
    msc_bump_kernel<<<1, 1, 0, M->stream>>>(M->d_counters, 2);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

}  // namespace nlmc

extern "C" {

int nlmc_msc_destroy(nlmc_msc *M) {
    if (!M) return NLMC_OK;
    cudaSetDevice(M->inst->device);
    nlmc::drop_graphs(M);
    void *ptrs[] = {M->S, M->recM, M->recE, M->rec, M->site_list, M->pos_of, M->thr_nz, M->rows, M->thr, M->betas, M->E_acc, M->E, M->swapmask, M->accepted,
                    M->scratch_spins, M->d_counters, M->labels, M->slot_of, M->thr_total,
                    M->betas_total, M->thrbits, M->thr_lane, M->accepted_rounds};
    if (M->stream) {   // every buffer goes back to the pool in stream order (after whatever is still queued)
        for (void *p : ptrs) nlmc::pool_free(p, M->stream);
        cudaStreamSynchronize(M->stream);
    }
    if (M->ev0) cudaEventDestroy(M->ev0);
    if (M->ev1) cudaEventDestroy(M->ev1);
    if (M->stream && M->own_stream) cudaStreamDestroy(M->stream);
    delete M;
    return NLMC_OK;
}

// label mode when betas_total != NULL: the handle owns the slots [slot_begin, slot_begin + n_beta) of a ladder of
// n_beta_total temperatures (betas = betas_total + slot_begin is what the slots start at)
static int msc_create_impl(nlmc_instance *I, int n_beta, const double *betas, int n_ladders, int ladder_offset,
                           unsigned long long seed, int n_beta_total, const double *betas_total, int slot_begin,
                           nlmc_msc **out) {
    using namespace nlmc;
    NLMC_REQUIRE(I && out && betas, "nlmc_msc_create: NULL argument");
    *out = nullptr;
    NLMC_REQUIRE(n_beta >= 1 && n_beta <= kMaxBeta, "nlmc_msc_create: n_beta must be in [1, %d]", kMaxBeta);
    NLMC_REQUIRE(n_ladders >= 1, "nlmc_msc_create: n_ladders must be >= 1");
    NLMC_REQUIRE(ladder_offset >= 0 && ladder_offset % 128 == 0, "nlmc_msc_create: ladder_offset must be a multiple of 128");
    const int n = I->n;
    // eligibility: J in {-1,+1} off the diagonal, h = 0, degrees <= 6 (site ranges on the host workers; the first
    // offending site of each range is reported in site order)
    std::vector<int32_t> nbr((size_t)n * 6, -1);
    std::vector<uint32_t> meta((size_t)n, 0u);
    std::vector<uint8_t> deg((size_t)n, 0);
    const int parts = n >= (1 << 16) ? nlmc::host_threads_shared() : 1;
    struct Issue { int kind = 0, i = -1, j = -1; double v = 0.0; };   // 1: h != 0, 2: bad value, 3: degree > 6
    std::vector<Issue> issues((size_t)parts);
    std::vector<long long> entries((size_t)parts, 0);
    nlmc::parallel_for(parts, [&](int t, int np) {
        const int per = (n + np - 1) / np, lo = std::min(n, per * t), hi = std::min(n, lo + per);
        Issue &is = issues[(size_t)t];
        for (int i = lo; i < hi && !is.kind; ++i) {
            if (I->h_h[(size_t)i] != 0.0) { is = {1, i, -1, I->h_h[(size_t)i]}; break; }
            int d = 0;
            for (int p = I->h_row_ptr[i]; p < I->h_row_ptr[i + 1]; ++p) {
                const double v = I->h_val[(size_t)p];
                const int j = I->h_col[(size_t)p];
                if (v == 0.0) continue;
                if (j == i || (v != 1.0 && v != -1.0)) { is = {2, i, j, v}; break; }
                if (d == 6) { is = {3, i, -1, 0.0}; break; }
                nbr[(size_t)i * 6 + d] = j;
                if (v < 0) meta[(size_t)i] |= 1u << d;
                ++d;
            }
            entries[(size_t)t] += d;
            deg[(size_t)i] = (uint8_t)d;
            // padding: the even slots past the last neighbour count as +1 (flip of a zeroed word), the odd ones as -1 --
            // neutral pairs, plus one slot at -1 when the degree is odd
            for (int e = d + (d & 1); e < 6; e += 2) meta[(size_t)i] |= 1u << e;
        }
    });
    long long n_entries = 0;
    for (int t = 0; t < parts; ++t) {
        const Issue &is = issues[(size_t)t];
        if (is.kind == 1) set_error("nlmc_msc_create: the bit-packed path needs h = 0 (h[%d] = %g)", is.i, is.v);
        if (is.kind == 2)
            set_error("nlmc_msc_create: the bit-packed path needs J in {-1,+1} with zero diagonal (J[%d,%d] = %g)", is.i, is.j, is.v);
        if (is.kind == 3) set_error("nlmc_msc_create: the bit-packed path supports degrees <= 6 (site %d)", is.i);
        if (is.kind) return NLMC_ERR_UNSUPPORTED;
        n_entries += entries[(size_t)t];
    }
    // symmetric adjacency is required for a valid colouring / detailed balance
    std::vector<int> asym_i((size_t)parts, -1), asym_j((size_t)parts, -1);
    nlmc::parallel_for(parts, [&](int t, int np) {
        const int per = (n + np - 1) / np, lo = std::min(n, per * t), hi = std::min(n, lo + per);
        for (int i = lo; i < hi && asym_i[(size_t)t] < 0; ++i)
            for (int d = 0; d < 6; ++d) {
                const int j = nbr[(size_t)i * 6 + d];
                if (j < 0) continue;
                bool back = false;
                for (int e = 0; e < 6; ++e) back |= nbr[(size_t)j * 6 + e] == i;
                if (!back) { asym_i[(size_t)t] = i; asym_j[(size_t)t] = j; break; }
            }
    });
    for (int t = 0; t < parts; ++t)
        NLMC_REQUIRE(asym_i[(size_t)t] < 0, "nlmc_msc_create: J must be symmetric (entry %d,%d has no transpose)",
                     asym_i[(size_t)t], asym_j[(size_t)t]);
    // greedy colouring in site order (checkerboard on bipartite lattices with even L)
    std::vector<int> colour((size_t)n, -1);
    int n_colours = 0;
    for (int i = 0; i < n; ++i) {
        unsigned used = 0;
        for (int d = 0; d < 6; ++d) {
            const int j = nbr[(size_t)i * 6 + d];
            if (j >= 0 && colour[(size_t)j] >= 0) used |= 1u << colour[(size_t)j];
        }
        int c = 0;
        while (used & (1u << c)) ++c;
        colour[(size_t)i] = c;
        n_colours = std::max(n_colours, c + 1);
    }
    // sites sorted by (colour, parity of the degree): a launch class is uniform in both
    std::vector<int32_t> site_list((size_t)n);
    std::vector<int> colour_ptr((size_t)n_colours + 1, 0);
    std::vector<int> class_cnt((size_t)n_colours * 2, 0);
    for (int i = 0; i < n; ++i) ++class_cnt[(size_t)colour[(size_t)i] * 2 + (deg[(size_t)i] & 1)];
    std::vector<nlmc_msc::SiteClass> classes;
    bool has_odd = false;
    {
        std::vector<int> fill((size_t)n_colours * 2, 0);
        int at = 0;
        for (int c = 0; c < n_colours; ++c) {
            colour_ptr[(size_t)c] = at;
            for (int par = 0; par < 2; ++par) {
                fill[(size_t)c * 2 + par] = at;
                const int cnt = class_cnt[(size_t)c * 2 + par];
                if (cnt) classes.push_back({at, cnt, par});
                if (cnt && par) has_odd = true;
                at += cnt;
            }
        }
        colour_ptr[(size_t)n_colours] = at;
        for (int i = 0; i < n; ++i) site_list[(size_t)fill[(size_t)colour[(size_t)i] * 2 + (deg[(size_t)i] & 1)]++] = i;
    }
    NLMC_CUDA(cudaSetDevice(I->device));
    auto *M = new nlmc_msc();
    M->inst = I;
    M->n = n;
    M->n_beta = n_beta;
    M->n_ladders = ((n_ladders + 127) / 128) * 128;
    M->ladder_offset = ladder_offset;
    M->G = M->n_ladders / 32;
    M->W = n_beta * M->G;
    M->n_colours = n_colours;
    M->n_bonds = n_entries / 2;
    M->colour_ptr = colour_ptr;
    M->classes = classes;
    M->has_odd = has_odd;
    M->seed = seed;
    M->h_betas.assign(betas, betas + n_beta);
    M->label_mode = betas_total ? 1 : 0;
    M->n_beta_total = betas_total ? n_beta_total : n_beta;
    M->slot_begin = betas_total ? slot_begin : 0;
    if (const char *e = getenv("NLMC_MSC_STEPS")) M->k_steps = atoi(e);
    if (M->k_steps != 4 && M->k_steps != 6 && M->k_steps != 7 && M->k_steps != 8) M->k_steps = 5;
    if (const char *e = getenv("NLMC_MSC_MERGED")) M->k_merged = atoi(e) ? NLMC_MERGED_STEPS : 0;
    if (const char *e = getenv("NLMC_MSC_GRAPHS")) M->use_graphs = atoi(e) != 0;
    const std::vector<uint32_t> thr = nlmc::msc_thresholds(n_beta, betas);
    M->h_thr = thr;
    M->h_thr_odd = nlmc::msc_thresholds(n_beta, betas, true);
    std::vector<int32_t> pos_of((size_t)n);
    for (int p = 0; p < n; ++p) pos_of[(size_t)site_list[(size_t)p]] = p;
    std::vector<int32_t> rec((size_t)n * 8, 0);  // records in site_list order: positions of the 6 neighbours, sign bits, site
    nlmc::parallel_for(parts, [&](int t, int np) {
        const int per = (n + np - 1) / np, lo = std::min(n, per * t), hi = std::min(n, lo + per);
        for (int p = lo; p < hi; ++p) {
            const int i = site_list[(size_t)p];
            for (int d = 0; d < 6; ++d) {
                const int j = nbr[(size_t)i * 6 + d];
                rec[(size_t)p * 8 + d] = j >= 0 ? pos_of[(size_t)j] : -1;
            }
            rec[(size_t)p * 8 + 6] = (int32_t)meta[(size_t)i];
            rec[(size_t)p * 8 + 7] = i;
        }
    });
    const size_t words = (size_t)n * M->W;
    // every buffer comes from the stream-ordered pool and every copy / memset is queued on the handle's stream: a
    // handle per NPT.run call costs no cudaMalloc / cudaFree (19 -> 6 ms per create at C5 size)
    bool ok = cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&M->ev0) == cudaSuccess && cudaEventCreate(&M->ev1) == cudaSuccess;
    cudaStream_t st = M->stream;
    auto alloc = [&](auto **ptr, size_t bytes) {
        return nlmc::pool_alloc(reinterpret_cast<void **>(ptr), bytes, I->device, st) == cudaSuccess;
    };
    auto put = [&](void *dst, const void *src, size_t bytes) {
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
    };
    ok = ok && alloc(&M->S, sizeof(uint32_t) * words) && alloc(&M->rec, sizeof(int32_t) * (size_t)n * 8) &&
         alloc(&M->site_list, sizeof(int32_t) * (size_t)n) && alloc(&M->pos_of, sizeof(int32_t) * (size_t)n) &&
         alloc(&M->thr, sizeof(uint32_t) * thr.size()) &&
         alloc(&M->betas, sizeof(double) * (size_t)n_beta) && alloc(&M->E_acc, sizeof(int32_t) * (size_t)M->W * 32) &&
         alloc(&M->E, sizeof(double) * (size_t)n_beta * M->n_ladders) &&
         alloc(&M->swapmask, sizeof(uint32_t) * (size_t)std::max(1, n_beta - 1) * M->G) && alloc(&M->accepted, sizeof(int32_t)) &&
         alloc(&M->accepted_rounds, sizeof(int32_t) * kRoundLog) && alloc(&M->d_counters, 4 * sizeof(uint32_t)) &&
         alloc(&M->scratch_spins, (size_t)n) &&
         cudaMemsetAsync(M->accepted_rounds, 0, sizeof(int32_t) * kRoundLog, st) == cudaSuccess &&
         cudaMemsetAsync(M->d_counters, 0, 4 * sizeof(uint32_t), st) == cudaSuccess &&
         cudaMemsetAsync(M->accepted, 0, sizeof(int32_t), st) == cudaSuccess &&
         put(M->rec, rec.data(), sizeof(int32_t) * rec.size()) &&
         put(M->site_list, site_list.data(), sizeof(int32_t) * site_list.size()) &&
         put(M->pos_of, pos_of.data(), sizeof(int32_t) * pos_of.size()) &&
         put(M->thr, thr.data(), sizeof(uint32_t) * thr.size()) && put(M->betas, betas, sizeof(double) * (size_t)n_beta);
    std::vector<uint32_t> thr_t;
    if (ok && M->label_mode) {
        // two sets of every threshold table: levels |f| = 2, 4, 6 (even degree), then |f| = 1, 3, 5 (odd degree)
        thr_t = nlmc::msc_thresholds(n_beta_total, betas_total);
        const std::vector<uint32_t> thr_o = nlmc::msc_thresholds(n_beta_total, betas_total, true);
        thr_t.insert(thr_t.end(), thr_o.begin(), thr_o.end());
        const size_t nl = (size_t)n_beta_total * M->n_ladders;
        ok = alloc(&M->labels, nl) && alloc(&M->slot_of, nl) && alloc(&M->thr_total, sizeof(uint32_t) * thr_t.size()) &&
             alloc(&M->betas_total, sizeof(double) * (size_t)n_beta_total) &&
             alloc(&M->thrbits, sizeof(uint32_t) * 2 * (size_t)(M->k_steps + M->k_merged) * 3 * M->W) &&
             alloc(&M->thr_lane, sizeof(uint32_t) * 2 * (size_t)M->W * 32 * 4) &&
             alloc(&M->thr_nz, sizeof(uint32_t) * 2 * (size_t)(M->W / 4)) &&
             put(M->thr_total, thr_t.data(), sizeof(uint32_t) * thr_t.size()) &&
             put(M->betas_total, betas_total, sizeof(double) * (size_t)n_beta_total);
    }
    // the host vectors above (and the caller's betas) must outlive the queued copies
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("nlmc_msc_create: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        nlmc_msc_destroy(M);
        return NLMC_ERR_CUDA;
    }
    if (M->label_mode) {
        const size_t nl = (size_t)n_beta_total * M->n_ladders;
        msc_labels_identity_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, M->stream>>>(n_beta_total, M->n_ladders, M->labels,
                                                                                      M->slot_of);
        const int rc = launch_thrbits(M);
        if (rc) { nlmc_msc_destroy(M); return rc; }
    }
    *out = M;
    return nlmc_msc_init_random(M, 0);
}

int nlmc_msc_create(nlmc_instance *I, int n_beta, const double *betas, int n_ladders, int ladder_offset,
                    unsigned long long seed, nlmc_msc **out) {
    return msc_create_impl(I, n_beta, betas, n_ladders, ladder_offset, seed, 0, nullptr, 0, out);
}

/* Beta-label form of the handle: slots [slot_begin, slot_begin + slot_count) of a ladder of n_beta_total temperatures
 * (slot s starts at betas_total[s]); exchanges permute the labels, never the configurations.  With slot_count ==
 * n_beta_total the handle is self-contained (nlmc_msc_round works); a block of a ladder sharded over GPUs is driven
 * with nlmc_msc_sweep / nlmc_msc_energies_dev / [all-gather by the caller] / nlmc_msc_exchange_labels. */
int nlmc_msc_create_labelled(nlmc_instance *I, int n_beta_total, const double *betas_total, int slot_begin, int slot_count,
                             int n_ladders, int ladder_offset, unsigned long long seed, nlmc_msc **out) {
    using namespace nlmc;
    NLMC_REQUIRE(out && betas_total, "nlmc_msc_create_labelled: NULL argument");
    NLMC_REQUIRE(n_beta_total >= 1 && n_beta_total <= kMaxBeta, "nlmc_msc_create_labelled: n_beta_total must be in [1, %d]", kMaxBeta);
    NLMC_REQUIRE(slot_begin >= 0 && slot_count >= 1 && slot_begin + slot_count <= n_beta_total,
                 "nlmc_msc_create_labelled: slot block [%d, %d) outside the ladder of %d", slot_begin, slot_begin + slot_count,
                 n_beta_total);
    return msc_create_impl(I, slot_count, betas_total + slot_begin, n_ladders, ladder_offset, seed, n_beta_total, betas_total,
                           slot_begin, out);
}

/* Run the handle on the caller's CUDA stream (e.g. the stream torch's NCCL collectives are ordered on), so that
 * sweeps, the energy all-gather and the label exchange queue up without host synchronisation.  NULL = back to the
 * handle's own stream. */
int nlmc_msc_set_stream(nlmc_msc *M, void *cuda_stream) {
    NLMC_REQUIRE(M, "nlmc_msc_set_stream: NULL handle");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    nlmc::drop_graphs(M);
    if (M->own_stream && M->stream) cudaStreamDestroy(M->stream);
    if (cuda_stream) {
        M->stream = static_cast<cudaStream_t>(cuda_stream);
        M->own_stream = false;
    } else {
        NLMC_CUDA(cudaStreamCreateWithFlags(&M->stream, cudaStreamNonBlocking));
        M->own_stream = true;
    }
    return NLMC_OK;
}

/* K4' into a DEVICE buffer: energies of this handle's slots, [slot_count][n_ladders] float64, queued on the handle's
 * stream without synchronisation (the send buffer of the energy all-gather). */
int nlmc_msc_energies_dev(nlmc_msc *M, double *out_E_dev) {
    NLMC_REQUIRE(M && out_E_dev, "nlmc_msc_energies_dev: NULL argument");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    int rc = nlmc::launch_energy(M);
    if (rc) return rc;
    if (out_E_dev != M->E)
        NLMC_CUDA(cudaMemcpyAsync(out_E_dev, M->E, sizeof(double) * (size_t)M->n_beta * M->n_ladders,
                                  cudaMemcpyDeviceToDevice, M->stream));
    return NLMC_OK;
}

/* K6 in label form on the gathered energies E_full_dev [n_beta_total][n_ladders] (device, slot-major): identical
 * decisions on every rank, labels permuted, this handle's thresholds rebuilt.  No synchronisation. */
int nlmc_msc_exchange_labels(nlmc_msc *M, const double *E_full_dev, int num_swapping_pairs) {
    NLMC_REQUIRE(M && E_full_dev, "nlmc_msc_exchange_labels: NULL argument");
    NLMC_REQUIRE(M->label_mode, "nlmc_msc_exchange_labels: the handle was not created with nlmc_msc_create_labelled");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    return nlmc::launch_label_exchange(M, E_full_dev, num_swapping_pairs);
}

/* labels[slot][ladder] = index of the temperature the configuration in (slot, ladder) is simulated at; identity for
 * handles without label exchange.  [n_beta_total][n_ladders_padded] uint8. */
int nlmc_msc_get_labels(nlmc_msc *M, uint8_t *out_labels) {
    NLMC_REQUIRE(M && out_labels, "nlmc_msc_get_labels: NULL argument");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    if (!M->label_mode) {
        for (int b = 0; b < M->n_beta; ++b)
            for (int l = 0; l < M->n_ladders; ++l) out_labels[(size_t)b * M->n_ladders + l] = (uint8_t)b;
        return NLMC_OK;
    }
    NLMC_CUDA(cudaMemcpyAsync(out_labels, M->labels, (size_t)M->n_beta_total * M->n_ladders, cudaMemcpyDeviceToHost, M->stream));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

/* Accepted exchanges of each of the last n_rounds rounds (oldest first), counted on the device per round. */
int nlmc_msc_swap_counts(nlmc_msc *M, int n_rounds, int *out_counts) {
    using namespace nlmc;
    NLMC_REQUIRE(M && out_counts && n_rounds >= 0 && n_rounds <= kRoundLog, "nlmc_msc_swap_counts: bad arguments");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    std::vector<int32_t> log((size_t)kRoundLog);
    uint32_t counters[4];
    NLMC_CUDA(cudaMemcpyAsync(log.data(), M->accepted_rounds, sizeof(int32_t) * kRoundLog, cudaMemcpyDeviceToHost, M->stream));
    NLMC_CUDA(cudaMemcpyAsync(counters, M->d_counters, sizeof(counters), cudaMemcpyDeviceToHost, M->stream));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    for (int k = 0; k < n_rounds; ++k) {
        const long long r = (long long)counters[1] - n_rounds + k;
        out_counts[k] = r >= 0 ? log[(size_t)(r % kRoundLog)] : 0;
    }
    return NLMC_OK;
}

int nlmc_msc_init_random(nlmc_msc *M, unsigned stream_id) {
    using namespace nlmc;
    NLMC_REQUIRE(M, "nlmc_msc_init_random: NULL handle");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    const size_t quads = (size_t)M->n * M->W / 4;
    msc_init_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, M->stream>>>(dev_view(M), stream_id);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

int nlmc_msc_info(const nlmc_msc *M, int *n_words, int *n_ladders_padded, int *n_colours, long long *n_bonds) {
    NLMC_REQUIRE(M, "nlmc_msc_info: NULL handle");
    if (n_words) *n_words = M->W;
    if (n_ladders_padded) *n_ladders_padded = M->n_ladders;
    if (n_colours) *n_colours = M->n_colours;
    if (n_bonds) *n_bonds = M->n_bonds;
    return NLMC_OK;
}

int nlmc_msc_set_betas(nlmc_msc *M, const double *betas) {
    NLMC_REQUIRE(M && betas, "nlmc_msc_set_betas: NULL argument");
    NLMC_REQUIRE(!M->label_mode, "nlmc_msc_set_betas: not available on a labelled handle (the ladder is fixed at creation)");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    const std::vector<uint32_t> thr = nlmc::msc_thresholds(M->n_beta, betas);
    M->h_betas.assign(betas, betas + M->n_beta);
    M->h_thr = thr;
    M->h_thr_odd = nlmc::msc_thresholds(M->n_beta, betas, true);
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    nlmc::drop_graphs(M);  // the thresholds are a launch argument of the captured sweeps
    NLMC_CUDA(cudaMemcpy(M->thr, thr.data(), sizeof(uint32_t) * thr.size(), cudaMemcpyHostToDevice));
    NLMC_CUDA(cudaMemcpy(M->betas, betas, sizeof(double) * (size_t)M->n_beta, cudaMemcpyHostToDevice));
    return NLMC_OK;
}

int nlmc_msc_set_seed(nlmc_msc *M, unsigned long long seed, unsigned sweep_counter) {
    NLMC_REQUIRE(M, "nlmc_msc_set_seed: NULL handle");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    M->seed = seed;
    nlmc::drop_graphs(M);  // the seed is a kernel argument of the captured launches
    const uint32_t c[2] = {sweep_counter, 0u};
    NLMC_CUDA(cudaMemcpy(M->d_counters, c, sizeof(c), cudaMemcpyHostToDevice));
    return NLMC_OK;
}

int nlmc_msc_set_spins(nlmc_msc *M, int beta_idx, int ladder, const int8_t *spins) {
    using namespace nlmc;
    NLMC_REQUIRE(M && spins && beta_idx >= 0 && beta_idx < M->n_beta && ladder >= 0 && ladder < M->n_ladders,
                 "nlmc_msc_set_spins: index out of range");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaMemcpyAsync(M->scratch_spins, spins, (size_t)M->n, cudaMemcpyHostToDevice, M->stream));
    msc_pack_kernel<<<(M->n + 255) / 256, 256, 0, M->stream>>>(dev_view(M), beta_idx * M->G + ladder / 32, ladder % 32,
                                                               M->scratch_spins);
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

int nlmc_msc_get_spins(nlmc_msc *M, int beta_idx, int ladder, int8_t *out) {
    using namespace nlmc;
    NLMC_REQUIRE(M && out && beta_idx >= 0 && beta_idx < M->n_beta && ladder >= 0 && ladder < M->n_ladders,
                 "nlmc_msc_get_spins: index out of range");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    msc_unpack_kernel<<<(M->n + 255) / 256, 256, 0, M->stream>>>(dev_view(M), beta_idx * M->G + ladder / 32, ladder % 32,
                                                                 M->scratch_spins);
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaMemcpyAsync(out, M->scratch_spins, (size_t)M->n, cudaMemcpyDeviceToHost, M->stream));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

// scratch for the site-major image of the state (the layout the C ABI exchanges)
static int msc_rows_scratch(nlmc_msc *M) {
    if (!M->rows)
        NLMC_CUDA(nlmc::pool_alloc(reinterpret_cast<void **>(&M->rows), sizeof(uint32_t) * (size_t)M->n * M->W, M->inst->device,
                                   M->stream));
    return NLMC_OK;
}

static int msc_set_packed_async(nlmc_msc *M, const uint32_t *packed) {
    using namespace nlmc;
    int rc = msc_rows_scratch(M);
    if (rc) return rc;
    const size_t quads = (size_t)M->n * (M->W / 4);
    NLMC_CUDA(cudaMemcpyAsync(M->rows, packed, sizeof(uint32_t) * (size_t)M->n * M->W, cudaMemcpyHostToDevice, M->stream));
    msc_from_rows_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, M->stream>>>(dev_view(M), reinterpret_cast<const uint4 *>(M->rows));
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

static int msc_get_packed_async(nlmc_msc *M, uint32_t *packed) {
    using namespace nlmc;
    int rc = msc_rows_scratch(M);
    if (rc) return rc;
    const size_t quads = (size_t)M->n * (M->W / 4);
    msc_to_rows_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, M->stream>>>(dev_view(M), reinterpret_cast<uint4 *>(M->rows));
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaMemcpyAsync(packed, M->rows, sizeof(uint32_t) * (size_t)M->n * M->W, cudaMemcpyDeviceToHost, M->stream));
    return NLMC_OK;
}

/* The packed state, site-major: packed[site][W] (word w = slot * G + ladder group; bit = ladder within the group).  The
 * device keeps it quad-major and colour-sorted; the transposition runs on the device. */
int nlmc_msc_set_packed(nlmc_msc *M, const uint32_t *packed) {
    NLMC_REQUIRE(M && packed, "nlmc_msc_set_packed: NULL argument");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    return msc_set_packed_async(M, packed);
}

int nlmc_msc_get_packed(nlmc_msc *M, uint32_t *packed) {
    NLMC_REQUIRE(M && packed, "nlmc_msc_get_packed: NULL argument");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    int rc = msc_get_packed_async(M, packed);
    if (rc) return rc;
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

int nlmc_msc_sweep(nlmc_msc *M, int n_sweeps) {
    NLMC_REQUIRE(M && n_sweeps >= 0, "nlmc_msc_sweep: bad arguments");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    return nlmc::run_round(M, n_sweeps, 0, false);
}

int nlmc_msc_energies(nlmc_msc *M, double *out_E) {
    NLMC_REQUIRE(M, "nlmc_msc_energies: NULL handle");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    int rc = nlmc::launch_energy(M);
    if (rc) return rc;
    if (out_E) {
        NLMC_CUDA(cudaMemcpyAsync(out_E, M->E, sizeof(double) * (size_t)M->n_beta * M->n_ladders, cudaMemcpyDeviceToHost,
                                  M->stream));
        NLMC_CUDA(cudaStreamSynchronize(M->stream));
    }
    return NLMC_OK;
}

int nlmc_msc_round(nlmc_msc *M, int n_sweeps, int num_swapping_pairs, double *out_E) {
    NLMC_REQUIRE(M && n_sweeps >= 0, "nlmc_msc_round: bad arguments");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    int rc;
    if (out_E) {  // energies of the states the sweeps produced (before the exchange), as the reference reads them
        if ((rc = nlmc::run_round(M, n_sweeps, 0, false))) return rc;
        if ((rc = nlmc::launch_energy(M))) return rc;
        NLMC_CUDA(cudaMemcpyAsync(out_E, M->E, sizeof(double) * (size_t)M->n_beta * M->n_ladders, cudaMemcpyDeviceToHost,
                                  M->stream));
        if ((rc = nlmc::launch_swap(M, num_swapping_pairs))) return rc;
        NLMC_CUDA(cudaStreamSynchronize(M->stream));
        return NLMC_OK;
    }
    return nlmc::run_round(M, n_sweeps, num_swapping_pairs, true);
}

/* n_sweeps sweeps with the state of one ladder (all betas) and/or the energies of all replicas recorded after
 * every sweep ON THE DEVICE and copied back once: the reference's M[:, jj] = m (NMC/nmc.py:89) and per-sweep energy
 * loops (NPT/npt.py:40-43, NPT/apt_preprocessor.py:107-110) without a host round trip per sweep. */
int nlmc_msc_sweep_record(nlmc_msc *M, int n_sweeps, int ladder, int8_t *out_M, double *out_E) {
    return nlmc_msc_sweep_record_layout(M, n_sweeps, ladder, out_M, out_E, 0);
}

/* m_layout 0: out_M [n_sweeps][n_beta][n]; 1: out_M [n_beta][n][n_sweeps] (rows of the reference's M). */
// keep_M_on_device: the states are recorded into M->recM and left there (the caller fetches them itself)
static int msc_record_impl(nlmc_msc *M, int n_sweeps, int ladder, int8_t *out_M, double *out_E, int m_layout,
                           bool keep_M_on_device, bool keep_E_on_device = false) {
    using namespace nlmc;
    NLMC_REQUIRE(m_layout == 0 || m_layout == 1, "nlmc_msc_sweep_record_layout: m_layout must be 0 or 1");
    const int n_sweeps_T = m_layout == 1 ? n_sweeps : 0;
    NLMC_REQUIRE(M && n_sweeps >= 0, "nlmc_msc_sweep_record: bad arguments");
    const bool has_M = out_M != nullptr || keep_M_on_device, has_E = out_E != nullptr || keep_E_on_device;
    NLMC_REQUIRE(!has_M || (ladder >= 0 && ladder < M->n_ladders), "nlmc_msc_sweep_record: ladder out of range");
    if (n_sweeps == 0) return NLMC_OK;
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    const size_t m_stride = (size_t)M->n_beta * M->n, e_stride = (size_t)M->n_beta * M->n_ladders;
    // grow-only record buffers; the captured graphs hold their addresses, so a reallocation drops the graphs
    const size_t need_M = has_M ? m_stride * (size_t)n_sweeps : 0, need_E = has_E ? e_stride * (size_t)n_sweeps : 0;
    if (need_M > M->recM_cap || need_E > M->recE_cap) {
        NLMC_CUDA(cudaStreamSynchronize(M->stream));
        for (auto &g : M->rec_graphs) cudaGraphExecDestroy(g.exec);
        M->rec_graphs.clear();
        if (need_M > M->recM_cap) {
            nlmc::pool_free(M->recM, M->stream);
            M->recM = nullptr; M->recM_cap = 0;
            NLMC_CUDA(nlmc::pool_alloc(reinterpret_cast<void **>(&M->recM), need_M, M->inst->device, M->stream));
            M->recM_cap = need_M;
        }
        if (need_E > M->recE_cap) {
            nlmc::pool_free(M->recE, M->stream);
            M->recE = nullptr; M->recE_cap = 0;
            NLMC_CUDA(nlmc::pool_alloc(reinterpret_cast<void **>(&M->recE), sizeof(double) * need_E, M->inst->device, M->stream));
            M->recE_cap = need_E;
        }
    }
    NLMC_CUDA(cudaMemsetAsync(M->d_counters + 2, 0, sizeof(uint32_t), M->stream));  // record slot 0
    int rc = NLMC_OK;
    if (M->use_graphs && n_sweeps >= 4) {
        cudaGraphExec_t exec = nullptr;
        for (auto &g : M->rec_graphs)
            if (g.ladder == (has_M ? ladder : -1) && g.has_M == has_M && g.has_E == has_E && g.n_sweeps_T == n_sweeps_T) exec = g.exec;
        if (!exec) {
            cudaGraph_t graph = nullptr;
            NLMC_CUDA(cudaStreamBeginCapture(M->stream, cudaStreamCaptureModeThreadLocal));
            rc = launch_recorded_sweep(M, ladder, has_M, has_E, m_stride, e_stride, n_sweeps_T);
            const cudaError_t e = cudaStreamEndCapture(M->stream, &graph);
            if (rc || e != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                if (!rc) set_error("nlmc_msc_sweep_record: stream capture failed: %s", cudaGetErrorString(e));
                return rc ? rc : NLMC_ERR_CUDA;
            }
            const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e2 != cudaSuccess) {
                set_error("nlmc_msc_sweep_record: cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
                return NLMC_ERR_CUDA;
            }
            if (M->rec_graphs.size() >= 4) {
                cudaGraphExecDestroy(M->rec_graphs.front().exec);
                M->rec_graphs.erase(M->rec_graphs.begin());
            }
            M->rec_graphs.push_back({has_M ? ladder : -1, has_M, has_E, n_sweeps_T, exec});
        }
        for (int s = 0; s < n_sweeps; ++s) NLMC_CUDA(cudaGraphLaunch(exec, M->stream));
    } else {
        for (int s = 0; s < n_sweeps && !rc; ++s) rc = launch_recorded_sweep(M, ladder, has_M, has_E, m_stride, e_stride, n_sweeps_T);
        if (rc) return rc;
    }
    if (has_M && !keep_M_on_device) NLMC_CUDA(cudaMemcpyAsync(out_M, M->recM, need_M, cudaMemcpyDeviceToHost, M->stream));
    if (has_E && !keep_E_on_device)
        NLMC_CUDA(cudaMemcpyAsync(out_E, M->recE, sizeof(double) * need_E, cudaMemcpyDeviceToHost, M->stream));
    if (!keep_M_on_device && !keep_E_on_device) NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

int nlmc_msc_sweep_record_layout(nlmc_msc *M, int n_sweeps, int ladder, int8_t *out_M, double *out_E, int m_layout) {
    return msc_record_impl(M, n_sweeps, ladder, out_M, out_E, m_layout, false);
}

/* nlmc_msc_sweep_record_layout with the states delivered as the float64 rows of the reference's M
 * (out_M_f64 [n_beta][n][n_sweeps], NPT/npt.py:640-644).  The int8 record stays on the device; it comes back in chunks
 * through a pinned staging buffer kept by the library (25 GB/s instead of the 4 GB/s of a copy into pageable memory), and
 * the host workers widen chunk k (non-temporal stores) while chunk k+1 is still on the wire. */
int nlmc_msc_sweep_record_f64(nlmc_msc *M, int n_sweeps, int ladder, double *out_M_f64, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(M && out_M_f64 && n_sweeps >= 0, "nlmc_msc_sweep_record_f64: bad arguments");
    if (n_sweeps == 0) return NLMC_OK;
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    // record on the device only, then fetch chunk by chunk (nlmc_host_fetch_widen_blocks)
    int rc = msc_record_impl(M, n_sweeps, ladder, nullptr, out_E, 1, true);
    if (rc) return rc;
    rc = nlmc_host_fetch_widen_blocks(M->recM, out_M_f64, 1, (uint64_t)M->n_beta * M->n * (uint64_t)n_sweeps, nullptr,
                                      M->inst->device, M->stream);
    if (rc) return rc;
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

int nlmc_msc_sweep_record_dev(nlmc_msc *M, int n_sweeps, int ladder, int m_layout, int8_t **out_M_dev, double **out_E_dev) {
    NLMC_REQUIRE(M && n_sweeps > 0 && (out_M_dev || out_E_dev), "nlmc_msc_sweep_record_dev: bad arguments");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    const int rc = msc_record_impl(M, n_sweeps, ladder, nullptr, nullptr, m_layout, out_M_dev != nullptr, out_E_dev != nullptr);
    if (rc) return rc;
    if (out_M_dev) *out_M_dev = M->recM;
    if (out_E_dev) *out_E_dev = M->recE;
    return NLMC_OK;
}

int nlmc_msc_round_host_async(nlmc_msc *M, const uint32_t *packed_in, int n_sweeps, int num_swapping_pairs,
                              uint32_t *packed_out, double *out_E) {
    NLMC_REQUIRE(M && n_sweeps >= 0, "nlmc_msc_round_host: bad arguments");
    int rc;
    if (packed_in && (rc = nlmc_msc_set_packed(M, packed_in))) return rc;
    if ((rc = nlmc_msc_round(M, n_sweeps, num_swapping_pairs, nullptr))) return rc;
    if (out_E)
        NLMC_CUDA(cudaMemcpyAsync(out_E, M->E, sizeof(double) * (size_t)M->n_beta * M->n_ladders, cudaMemcpyDeviceToHost,
                                  M->stream));
    if (packed_out && (rc = msc_get_packed_async(M, packed_out))) return rc;
    return NLMC_OK;
}

int nlmc_msc_round_host(nlmc_msc *M, const uint32_t *packed_in, int n_sweeps, int num_swapping_pairs,
                        uint32_t *packed_out, double *out_E) {
    const int rc = nlmc_msc_round_host_async(M, packed_in, n_sweeps, num_swapping_pairs, packed_out, out_E);
    if (rc) return rc;
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

int nlmc_msc_swap_count(nlmc_msc *M, int *out_accepted, int reset) {
    NLMC_REQUIRE(M && out_accepted, "nlmc_msc_swap_count: NULL argument");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaMemcpyAsync(out_accepted, M->accepted, sizeof(int32_t), cudaMemcpyDeviceToHost, M->stream));
    if (reset) NLMC_CUDA(cudaMemsetAsync(M->accepted, 0, sizeof(int32_t), M->stream));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

int nlmc_msc_sync(nlmc_msc *M) {
    NLMC_REQUIRE(M, "nlmc_msc_sync: NULL handle");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaStreamSynchronize(M->stream));
    return NLMC_OK;
}

/* CUDA-event timing on the library's own stream (bench.py): mark 0 = start, mark 1 = stop. */
int nlmc_msc_timer_mark(nlmc_msc *M, int which) {
    NLMC_REQUIRE(M && (which == 0 || which == 1), "nlmc_msc_timer_mark: bad arguments");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaEventRecord(which == 0 ? M->ev0 : M->ev1, M->stream));
    return NLMC_OK;
}

int nlmc_msc_timer_elapsed_ms(nlmc_msc *M, float *out_ms) {
    NLMC_REQUIRE(M && out_ms, "nlmc_msc_timer_elapsed_ms: NULL argument");
    NLMC_CUDA(cudaSetDevice(M->inst->device));
    NLMC_CUDA(cudaEventSynchronize(M->ev1));
    NLMC_CUDA(cudaEventElapsedTime(out_ms, M->ev0, M->ev1));
    return NLMC_OK;
}

}  // extern "C"
