/* nlmc_npmath.h -- float64 tanh / arctanh with the rounding behaviour of numpy 2.3.x on AVX-512 hosts.
 *
 * The reference's LBP (NMC/nmc.py:200-216, 230-255) calls np.tanh / np.arctanh with tolerance = machine epsilon,
 * so its iteration counts, divergence points and therefore the backbone depend on the last bit of those
 * functions.  They are not part of /root/reference: they live in numpy (the build the goldens were made with is
 * numpy 2.3.5, x86-64, AVX512_SKX dispatch):
 *   - np.tanh  float64 -> numpy/_core/src/umath/loops_hyperbolic.dispatch.c.src, simd_tanh_f64: 16 intervals
 *     selected by the exponent and the top mantissa bit of |x|, a degree-16 polynomial in (|x| - b) per
 *     interval evaluated by Horner with fused multiply-adds.
 *   - np.arctanh float64 -> Intel SVML __svml_atanh8 (numpy/_core/src/umath/svml, svml_z0_atanh_d_la.s):
 *     atanh x = (log(1+|x|) - log(1-|x|))/2, each log reduced with a 1+4-bit rounding of VRCP14PD, a 16-entry
 *     hi/lo table of log(1+i/16) and a degree-9 log1p polynomial, all steps fused multiply-adds.
 * This file restates both algorithms operation for operation; every product-sum below is a single fma(), which
 * rounds identically on x86 (FMA3/AVX-512) and in CUDA fp64.  The constants are the ones those routines load.
 * VRCP14PD itself is not restated: only its value rounded to 1+4 mantissa bits is used, and that is a step
 * function of the top 20 mantissa bits of the argument whose 16 steps were tabulated exhaustively over all 2^20
 * prefixes on a Sapphire Rapids host (the instruction is architecturally deterministic).
 * Pinned bit for bit against numpy by tests/test_npmath.py (CPU, via oracle/) and tests/test_gpu_npmath.py.
 */
#ifndef NLMC_NPMATH_H
#define NLMC_NPMATH_H
#include <stdint.h>
#ifdef __CUDACC__
#define NLMC_NPM_FN __device__ __forceinline__
#define NLMC_NPM_UNROLL _Pragma("unroll")
#define NLMC_NPM_TABLE static __device__ const
#define NLMC_NPM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define NLMC_NPM_ADD(a, b) __dadd_rn((a), (b))
#define NLMC_NPM_SUB(a, b) __dsub_rn((a), (b))
#define NLMC_NPM_MUL(a, b) __dmul_rn((a), (b))
#define NLMC_NPM_D2U(x) ((uint64_t)__double_as_longlong(x))
#define NLMC_NPM_U2D(u) __longlong_as_double((long long)(u))
#else
#include <math.h>
#include <string.h>
#define NLMC_NPM_FN static inline
#define NLMC_NPM_UNROLL
#define NLMC_NPM_TABLE static const
#define NLMC_NPM_FMA(a, b, c) fma((a), (b), (c))
#define NLMC_NPM_ADD(a, b) ((a) + (b))
#define NLMC_NPM_SUB(a, b) ((a) - (b))
#define NLMC_NPM_MUL(a, b) ((a) * (b))
static inline uint64_t nlmc_npm_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double nlmc_npm_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define NLMC_NPM_D2U(x) nlmc_npm_d2u(x)
#define NLMC_NPM_U2D(u) nlmc_npm_u2d(u)
#endif

/* tanh: row 0 = interval centres b, rows 1..17 = c0..c16, 16 intervals per row. */
NLMC_NPM_TABLE uint64_t nlmc_npm_tanh_lut[288] = {
    0x0000000000000000ull, 0x3fcc000000000000ull, 0x3fd4000000000000ull, 0x3fdc000000000000ull,
    0x3fe4000000000000ull, 0x3fec000000000000ull, 0x3ff4000000000000ull, 0x3ffc000000000000ull,
    0x4004000000000000ull, 0x400c000000000000ull, 0x4014000000000000ull, 0x401c000000000000ull,
    0x4024000000000000ull, 0x402c000000000000ull, 0x4034000000000000ull, 0x0000000000000000ull,
    0x0000000000000000ull, 0x3fcb8fd0416a7c92ull, 0x3fd35f98a0ea650eull, 0x3fda5729ee488037ull,
    0x3fe1bf47eabb8f95ull, 0x3fe686650b8c2015ull, 0x3feb2523bb6b2deeull, 0x3fee1fbf97e33527ull,
    0x3fef9258260a71c2ull, 0x3feff112c63a9077ull, 0x3fefff419668df11ull, 0x3feffffc832750f2ull,
    0x3feffffffdc96f35ull, 0x3fefffffffffcf58ull, 0x3ff0000000000000ull, 0x3ff0000000000000ull,
    0x3ff0000000000000ull, 0x3fee842ca3f08532ull, 0x3fed11574af58f1bull, 0x3fea945b9c24e4f9ull,
    0x3fe6284c3374f815ull, 0x3fe02500a09f8d6eull, 0x3fd1f25131e3a8c0ull, 0x3fbd22ca1c24a139ull,
    0x3f9b3afe1fba5c76ull, 0x3f6dd37d19b22b21ull, 0x3f27ccec13a9ef96ull, 0x3ecbe6c3f33250aeull,
    0x3e41b4865394f75full, 0x3d8853f01bda5f28ull, 0x3c73953c0197ef58ull, 0x0000000000000000ull,
    0xbbf0b3ea3fdfaa19ull, 0xbfca48aaeb53bc21ull, 0xbfd19921f4329916ull, 0xbfd5e0f09bef8011ull,
    0xbfd893b59c35c882ull, 0xbfd6ba7cb7576538ull, 0xbfce7291743d7555ull, 0xbfbb6d85a01efb80ull,
    0xbf9addae58c7141aull, 0xbf6dc59376c7aa19ull, 0xbf27cc5e74677410ull, 0xbecbe6c0e8b4cc87ull,
    0xbe41b486526b0565ull, 0xbd8853f01bef63a4ull, 0xbc73955be519be31ull, 0x0000000000000000ull,
    0xbfd5555555555555ull, 0xbfd183afc292ba11ull, 0xbfcc1a4b039c9bfaull, 0xbfc16e1e6d8d0be6ull,
    0xbf92426c751e48a2ull, 0x3fb4f152b2bad124ull, 0x3fbbba40cbef72beull, 0x3fb01ba038be6a3dull,
    0x3f916df44871efc8ull, 0x3f63c6869dfc8870ull, 0x3f1fb9aef915d828ull, 0x3ec299d1e27c6e11ull,
    0x3e379b5ddcca334cull, 0x3d8037f57bc62c9aull, 0x3c6a2d4b50a2cff7ull, 0x0000000000000000ull,
    0xbce6863ee44ed636ull, 0x3fc04dcd0476c75eull, 0x3fc43d3449a80f08ull, 0x3fc5c26f3699b7e7ull,
    0x3fc1a686f6ab2533ull, 0x3faf203c316ce730ull, 0xbf89c7a02788557cull, 0xbf98157e26e0d541ull,
    0xbf807b55c1c7d278ull, 0xbf53a18d5843190full, 0xbf0fb6bbc89b1a5bull, 0xbeb299c9c684a963ull,
    0xbe279b5dd4fb3d01ull, 0xbd7037f57ae72aa6ull, 0xbc5a2ca2bba78e86ull, 0x0000000000000000ull,
    0x3fc1111111112ab5ull, 0x3fb5c19efdfc08adull, 0x3fa74c98dc34fbacull, 0xbf790d6a8eff0a77ull,
    0xbfac3c021789a786ull, 0xbfae2196b7326859ull, 0xbf93a7a011ff8c2aull, 0x3f6e4709c7e8430eull,
    0x3f67682afa611151ull, 0x3f3ef2ee77717cbfull, 0x3ef95a4482f180b7ull, 0x3e9dc2c27da3b603ull,
    0x3e12e2afd9f7433eull, 0x3d59f320348679baull, 0x3c44b61d9bbcc940ull, 0x0000000000000000ull,
    0xbda1ea19ddddb3b4ull, 0xbfb0b8df995ce4dfull, 0xbfb2955cf41e8164ull, 0xbfaf9d05c309f7c6ull,
    0xbf987d27ccff4291ull, 0x3f8b2ca62572b098ull, 0x3f8f1cf6c7f5b00aull, 0x3f60379811e43dd5ull,
    0xbf4793826f78537eull, 0xbf2405695e36240full, 0xbee0e08de39ce756ull, 0xbe83d709ba5f714eull,
    0xbdf92e3fc5ee63e0ull, 0xbd414cc030f2110eull, 0xbc2ba022e8d82a87ull, 0x0000000000000000ull,
    0xbfaba1ba1990520bull, 0xbf96e37bba52f6fcull, 0x3ecff7df18455399ull, 0x3f97362834d33a4eull,
    0x3f9e7f8380184b45ull, 0x3f869543e7c420d4ull, 0xbf7326bd4914222aull, 0xbf5fc15b0a9d98faull,
    0x3f14cffcfa69fbb6ull, 0x3f057e48e5b79d10ull, 0x3ec33b66d7d77264ull, 0x3e66ac4e578b9b10ull,
    0x3ddcc74b8d3d5c42ull, 0x3d23c589137f92b4ull, 0x3c107f8e2c8707a1ull, 0x0000000000000000ull,
    0xbe351ca7f096011full, 0x3f9eaaf3320c3851ull, 0x3f9cf823fe761fc1ull, 0x3f9022271754ff1full,
    0xbf731fe77c9c60afull, 0xbf84a6046865ec7dull, 0xbf4ca3f1f2b9192bull, 0x3f4c77dee0afd227ull,
    0x3f04055bce68597aull, 0xbee2bf0cb4a71647ull, 0xbea31eaafe73efd5ull, 0xbe46abb02c4368edull,
    0xbdbcc749ca8079ddull, 0xbd03c5883836b9d2ull, 0xbbf07a5416264aecull, 0x0000000000000000ull,
    0x3f9664f94e6ac14eull, 0xbf94d3343bae39ddull, 0xbf7bc748e60df843ull, 0xbf8c89372b43ba85ull,
    0xbf8129a092de747aull, 0x3f60c85b4d538746ull, 0x3f5be9392199ec18ull, 0xbf2a0c68a4489f10ull,
    0xbf00462601dc2faaull, 0x3eb7b6a219dea9f4ull, 0x3e80cbcc8d4c5c8aull, 0x3e2425bb231a5e29ull,
    0x3d9992a4beac8662ull, 0x3ce191ba5ed3fb67ull, 0x3bc892450bad44c4ull, 0x0000000000000000ull,
    0xbea8c4c1fd7852feull, 0xbfccce16b1046f13ull, 0xbf81a16f224bb7b6ull, 0xbf62cbf00406bc09ull,
    0x3f75b29bb02cf69bull, 0x3f607df0f9f90c17ull, 0xbf4b852a6e0758d5ull, 0xbf0078c63d1b8445ull,
    0x3eec12eadd55be7aull, 0xbe6fa600f593181bull, 0xbe5a3c935dce3f7dull, 0xbe001c6d95e3ae96ull,
    0xbd74755a00ea1fd3ull, 0xbcbc1c6c063bb7acull, 0xbba3be9a4460fe00ull, 0x0000000000000000ull,
    0xbf822404577aa9ddull, 0x403d8b07f7a82aa3ull, 0xbf9f44ab92fbab0aull, 0x3fb2eac604473d6aull,
    0x3f45f87d903aaac8ull, 0xbf5e104671036300ull, 0x3f19bc98ddf0f340ull, 0x3f0d4304bc9246e8ull,
    0xbed13c415f7b9d41ull, 0xbe722b8d9720cdb0ull, 0x3e322666d739bec0ull, 0x3dd76a553d7e7918ull,
    0x3d4de0fa59416a39ull, 0x3c948716cf3681b4ull, 0x3b873f9f2d2fda99ull, 0x0000000000000000ull,
    0xbefdd99a221ed573ull, 0x4070593a3735bab4ull, 0xbfccab654e44835eull, 0x3fd13ed80037dbacull,
    0xbf6045b9076cc487ull, 0x3f2085ee7e8ac170ull, 0x3f23524622610430ull, 0xbeff12a6626911b4ull,
    0x3eab9008bca408afull, 0x3e634df71865f620ull, 0xbe05bb1bcf83ca73ull, 0xbdaf2ac143fb6762ull,
    0xbd23eae52a3dbf57ull, 0xbc6b5e3e9ca0955eull, 0xbb5eca68e2c1ba2eull, 0x0000000000000000ull,
    0x3f6e3be689423841ull, 0xc0d263511f5baac1ull, 0x40169f73b15ebe5cull, 0xc025c1dd41cd6cb5ull,
    0xbf58fd89fe05e0d1ull, 0x3f73f7af01d5af7aull, 0xbf1e40bdead17e6bull, 0x3ee224cd6c4513e5ull,
    0xbe24b645e68eeaa3ull, 0xbe4abfebfb72bc83ull, 0x3dd51c38f8695ed3ull, 0x3d8313ac38c6832bull,
    0x3cf7787935626685ull, 0x3c401ffc49c6bc29ull, 0xbabf0b21acfa52abull, 0x0000000000000000ull,
    0xbf2a1306713a4f3aull, 0xc1045e509116b066ull, 0x4041fab9250984ceull, 0xc0458d090ec3de95ull,
    0xbf74949d60113d63ull, 0x3f7c9fd6200d0adeull, 0x3f02cd40e0ad0a9full, 0xbe858ab8e019f311ull,
    0xbe792fa6323b7cf8ull, 0x3e2df04d67876402ull, 0xbd95c72be95e4d2cull, 0xbd55a89c30203106ull,
    0xbccad6b3bb9eff65ull, 0xbc12705ccd3dd884ull, 0xba8e0a4c47ae75f5ull, 0x0000000000000000ull,
    0xbf55d7e76dc56871ull, 0x41528c38809c90c7ull, 0xc076d57fb5190b02ull, 0x4085f09f888f8adaull,
    0x3fa246332a2fcba5ull, 0xbfb29d851a896fcdull, 0x3ed9065ae369b212ull, 0xbeb8e1ba4c98a030ull,
    0x3e6ffd0766ad4016ull, 0xbe0c63c29f505f5bull, 0xbd7fab216b9e0e49ull, 0x3d2826b62056aa27ull,
    0x3ca313e31762f523ull, 0x3bea37aa21895319ull, 0x3ae5c7f1fd871496ull, 0x0000000000000000ull,
    0x3f35e67ab76a26e7ull, 0x41848ee0627d8206ull, 0xc0a216d618b489ecull, 0x40a5b89107c8af4full,
    0x3fb69d8374520edaull, 0xbfbded519f981716ull, 0xbef02d288b5b3371ull, 0x3eb290981209c1a6ull,
    0xbe567e924bf5ff6eull, 0x3de3f7f7de6b0eb6ull, 0x3d69ed18bae3ebbcull, 0xbcf7534c4f3dfa71ull,
    0xbc730b73f1eaff20ull, 0xbbba2cff8135d462ull, 0xbab5a71b5f7d9035ull, 0x0000000000000000ull,
};

/* arctanh: log(1 + i/16), i = 0..15, as a high part and a correction. */
NLMC_NPM_TABLE uint64_t nlmc_npm_atanh_thi[16] = {
    0x0000000000000000ull, 0x3faf0a30c0120000ull, 0x3fbe27076e2b0000ull, 0x3fc5ff3070a78000ull,
    0x3fcc8ff7c79a8000ull, 0x3fd1675cababc000ull, 0x3fd4618bc21c4000ull, 0x3fd739d7f6bbc000ull,
    0x3fd9f323ecbf8000ull, 0x3fdc8ff7c79a8000ull, 0x3fdf128f5faf0000ull, 0x3fe0be72e4252000ull,
    0x3fe1e85f5e704000ull, 0x3fe307d7334f2000ull, 0x3fe41d8fe8468000ull, 0x3fe52a2d265bc000ull,
};
NLMC_NPM_TABLE uint64_t nlmc_npm_atanh_tlo[16] = {
    0x0000000000000000ull, 0xbd53ab33d066d1d2ull, 0xbd2a342c2af0003cull, 0x3d43d3c873e20a07ull,
    0x3d4a21ac25d81ef3ull, 0xbd59f1fc63382a8full, 0x3d5ec27d0b7b37b3ull, 0x3d50069ce24c53fbull,
    0x3d584bf2b68d766full, 0x3d5a21ac25d81ef3ull, 0x3d3bb2cd720ec44cull, 0x3d55056d312f7668ull,
    0x3d1a07bd8b34be7cull, 0xbd5e83c094debc15ull, 0xbd5aa33736867a17ull, 0x3d46abb9df22bc57ull,
};
/* log1p(q) = q + q^2 P(q): coefficients of P, highest degree first. */
NLMC_NPM_TABLE uint64_t nlmc_npm_atanh_poly[9] = {
    0xbfb9a9b040214368ull, 0x3fbc80666e249778ull, 0xbfbffffb8a054bc9ull, 0x3fc24922f71256f1ull,
    0xbfc55555559ba736ull, 0x3fc9999999be77afull, 0xbfcffffffffffc65ull, 0x3fd55555555554c1ull,
    0xbfe0000000000000ull,
};
/* log 2, high part and correction. */
NLMC_NPM_TABLE uint64_t nlmc_npm_atanh_ln2[2] = {
    0x3fe62e42fefa0000ull, 0x3d7cf79abc9e0000ull,
};
/* top-20-mantissa-bit positions at which the rounded VRCP14PD steps down by 1/16 */
NLMC_NPM_TABLE uint32_t nlmc_npm_rcp_steps[16] = {
    0x040f0, 0x0c980, 0x15b40, 0x1f700, 0x29e60, 0x35240, 0x41430, 0x4e600,
    0x5c990, 0x6c160, 0x7d070, 0x8f9d0, 0xa41a0, 0xbad10, 0xd41c0, 0xf0820,
};

/* np.tanh(x), float64 (simd_tanh_f64); `lut` is nlmc_npm_tanh_lut or a copy of it (e.g. in shared memory). */
NLMC_NPM_FN double nlmc_np_tanh_lut(double x, const uint64_t *lut)
{
    const uint64_t u = NLMC_NPM_D2U(x);
    const uint64_t au = u & 0x7fffffffffffffffull;
    if (au > 0x7ff0000000000000ull) return NLMC_NPM_U2D(0x7ff8000000000000ull);
    const uint64_t top = u & 0x7ff8000000000000ull;            /* exponent + first mantissa bit */
    double r;
    if (au >= 0x4038000000000000ull) {
        /* |x| >= 24: the last interval's polynomial is the constant 1 (c0 = 1, c1..c16 = 0), as is the routine's answer for
         * infinities and the top half-binade -- skip the 17 multiply-adds (frozen spins of the NMC phases sit at 1e4) */
        r = 1.0;
    } else {
        int64_t d = (int64_t)top - (int64_t)0x3fc0000000000000ull;
        int32_t hi = (int32_t)(d >> 32);
        hi = hi < 0 ? 0 : (hi > 0x780000 ? 0x780000 : hi);
        const int idx = hi >> 19;                               /* 0: |x| < 3/16 ... 15: |x| >= 24 */
        const double y = NLMC_NPM_SUB(NLMC_NPM_U2D(au), NLMC_NPM_U2D(lut[idx]));
        r = NLMC_NPM_U2D(lut[17 * 16 + idx]);
NLMC_NPM_UNROLL
        for (int k = 16; k >= 1; --k) r = NLMC_NPM_FMA(r, y, NLMC_NPM_U2D(lut[k * 16 + idx]));
    }
    return NLMC_NPM_U2D(NLMC_NPM_D2U(r) | (u & 0x8000000000000000ull));
}

NLMC_NPM_FN double nlmc_np_tanh(double x) { return nlmc_np_tanh_lut(x, nlmc_npm_tanh_lut); }

/* VRCP14PD(y) rounded to 1+4 mantissa bits (add 2^47 to the bit pattern, keep the top 16 bits), y normal > 0.
 * k = number of steps at or below the top 20 mantissa bits of y; the rounded reciprocal is
 * 2^-e for k = 0 and (1 + (16-k)/16) 2^(-e-1) otherwise. */
NLMC_NPM_FN uint64_t nlmc_npm_rcp_1p4(uint64_t ybits)
{
    const uint32_t p = (uint32_t)((ybits >> 32) & 0xfffffu);
    int k = 0;
NLMC_NPM_UNROLL
    for (int i = 0; i < 16; ++i) k += (p >= nlmc_npm_rcp_steps[i]) ? 1 : 0;
    const uint64_t e = (ybits >> 52) & 0x7ffu;
    return ((uint64_t)(2046u - e - (k > 0 ? 1u : 0u)) << 52) | ((uint64_t)((16 - k) & 15) << 48);
}

/* np.arctanh(x), float64 (__svml_atanh8_ha, the entry numpy's AVX512_SKX loop calls).  With
 * 1+|x| = yp + yp_lo, 1-|x| = ym - ym_nlo, Rp/Rm the rounded reciprocals and qp = Rp(1+|x|) - 1, qm = Rm(1-|x|) - 1:
 * 2 atanh|x| = [log Rm - log Rp] + log1p(qp) - log1p(qm); the bracket comes from the exponents and the table
 * (kh + kl), the leading terms kh + qp - qm are summed with their rounding errors carried (two-sum).
 * |x| >= 1 and NaN take the routine's scalar side path (+-inf at |x| = 1, NaN beyond). */
NLMC_NPM_FN double nlmc_np_arctanh(double x)
{
    const uint64_t u = NLMC_NPM_D2U(x);
    const double ax = NLMC_NPM_U2D(u & 0x7fffffffffffffffull);
    if (!(ax < 1.0)) {
        if (ax != ax) return NLMC_NPM_MUL(x, x);
        if (ax == 1.0) return NLMC_NPM_U2D((u & 0x8000000000000000ull) | 0x7ff0000000000000ull);
        return NLMC_NPM_U2D(0xfff8000000000000ull);
    }
    const double yp = NLMC_NPM_ADD(ax, 1.0), ym = NLMC_NPM_SUB(1.0, ax);
    const double yp_lo = NLMC_NPM_SUB(ax, NLMC_NPM_SUB(yp, 1.0));
    const double ym_nlo = NLMC_NPM_ADD(ax, NLMC_NPM_SUB(ym, 1.0));
    const uint64_t rpb = nlmc_npm_rcp_1p4(NLMC_NPM_D2U(yp)), rmb = nlmc_npm_rcp_1p4(NLMC_NPM_D2U(ym));
    const double rp = NLMC_NPM_U2D(rpb), rm = NLMC_NPM_U2D(rmb);
    double qp = NLMC_NPM_FMA(rp, yp, -1.0);
    qp = NLMC_NPM_FMA(yp_lo, rp, qp);
    double qm = NLMC_NPM_FMA(ym, rm, -1.0);
    qm = NLMC_NPM_FMA(-ym_nlo, rm, qm);
    const double dk = (double)((int)((rmb >> 52) & 0x7ffu) - (int)((rpb >> 52) & 0x7ffu));   /* VGETEXPPD difference */
    const int ip = (int)((rpb >> 48) & 15u), im = (int)((rmb >> 48) & 15u);
    const double tl = NLMC_NPM_SUB(NLMC_NPM_U2D(nlmc_npm_atanh_tlo[im]), NLMC_NPM_U2D(nlmc_npm_atanh_tlo[ip]));
    const double th = NLMC_NPM_SUB(NLMC_NPM_U2D(nlmc_npm_atanh_thi[im]), NLMC_NPM_U2D(nlmc_npm_atanh_thi[ip]));
    double pp = NLMC_NPM_FMA(NLMC_NPM_U2D(nlmc_npm_atanh_poly[0]), qp, NLMC_NPM_U2D(nlmc_npm_atanh_poly[1]));
    double pm = NLMC_NPM_FMA(NLMC_NPM_U2D(nlmc_npm_atanh_poly[0]), qm, NLMC_NPM_U2D(nlmc_npm_atanh_poly[1]));
NLMC_NPM_UNROLL
    for (int k = 2; k < 9; ++k) {
        pp = NLMC_NPM_FMA(qp, pp, NLMC_NPM_U2D(nlmc_npm_atanh_poly[k]));
        pm = NLMC_NPM_FMA(qm, pm, NLMC_NPM_U2D(nlmc_npm_atanh_poly[k]));
    }
    const double kh = NLMC_NPM_FMA(NLMC_NPM_U2D(nlmc_npm_atanh_ln2[0]), dk, th);
    const double kl = NLMC_NPM_FMA(NLMC_NPM_U2D(nlmc_npm_atanh_ln2[1]), dk, tl);
    const double s1 = NLMC_NPM_ADD(qp, kh);
    const double s2 = NLMC_NPM_SUB(s1, qm);
    const double e1 = NLMC_NPM_ADD(qp, NLMC_NPM_SUB(kh, s1));       /* qp + kh - s1 */
    const double e2 = NLMC_NPM_ADD(qm, NLMC_NPM_SUB(s2, s1));       /* qm + s2 - s1 */
    pp = NLMC_NPM_FMA(NLMC_NPM_MUL(qp, qp), pp, kl);
    pm = NLMC_NPM_FMA(-NLMC_NPM_MUL(qm, qm), pm, e1);
    const double t = NLMC_NPM_ADD(s2, NLMC_NPM_SUB(NLMC_NPM_ADD(pp, pm), e2));
    const double half = NLMC_NPM_U2D((u & 0x8000000000000000ull) | 0x3fe0000000000000ull);
    return NLMC_NPM_MUL(t, half);
}
#endif
