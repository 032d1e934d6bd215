// sde_device_rng.cuh — on-device uniform generation for sm_100a.
//
// Replaces src/rng/{pseudo,sobol}.rs of the reference.  Two generators:
//   * ChaCha8 keyed by rand's seed_from_u64 (PCG32 key expansion) — the exact stream of
//     ChaCha8Rng::seed_from_u64(seed + s) (src/rng/pseudo.rs:18,25; src/rng/sobol.rs:68-69).
//     It is counter based: block b of path s is a pure function of (seed + s, b).
//   * Gray-code Sobol by index (no shared iterator, src/rng/sobol.rs:17,41-44 disappears):
//     the map n -> x_d(n) is GF(2)-linear in the bits of n, so
//         x_d(n_cta + 32*w + lane) = x_d(n_cta) ^ x_d(32*w) ^ x_d(lane).
//     x_d(lane) (32 entries/dim) and x_d(32*w) (8 entries/dim) are tiny global tables,
//     x_d(n_cta) is folded once per CTA per time tile into shared memory, so a thread
//     pays one LDS + one (L1-resident) LDG + one XOR per dimension.
// Compiles under both nvcc and NVRTC (no host headers).
#pragma once

#ifndef SDE_TYPES_DEFINED
#define SDE_TYPES_DEFINED
typedef unsigned int sde_u32;
typedef unsigned long long sde_u64;
#endif

// ---------------------------------------------------------------- ChaCha8 -------------
__device__ __forceinline__ sde_u32 sde_rotl32(sde_u32 x, int r) { return __funnelshift_l(x, x, r); }

#define SDE_CHACHA_QR(a, b, c, d)                 \
    a += b; d ^= a; d = sde_rotl32(d, 16);        \
    c += d; b ^= c; b = sde_rotl32(b, 12);        \
    a += b; d ^= a; d = sde_rotl32(d, 8);         \
    c += d; b ^= c; b = sde_rotl32(b, 7);

// rand_core::SeedableRng::seed_from_u64 — PCG32 expansion of a u64 into the 256-bit key.
__device__ __forceinline__ void sde_seed_from_u64(sde_u64 state, sde_u32 (&key)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        state = state * 6364136223846793005ull + 11634580027462260723ull;
        sde_u32 xs = (sde_u32)(((state >> 18) ^ state) >> 27);
        sde_u32 rot = (sde_u32)(state >> 59);
        key[i] = __funnelshift_r(xs, xs, rot);
    }
}

// One ChaCha block with `ROUNDS` rounds; 64-bit block counter in words 12-13, stream id 0.
template <int ROUNDS>
__device__ __forceinline__ void sde_chacha_block(const sde_u32 (&key)[8], sde_u64 counter, sde_u32 (&out)[16]) {
    const sde_u32 c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
    sde_u32 x0 = c0, x1 = c1, x2 = c2, x3 = c3;
    sde_u32 x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3];
    sde_u32 x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    sde_u32 x12 = (sde_u32)counter, x13 = (sde_u32)(counter >> 32), x14 = 0u, x15 = 0u;
#pragma unroll
    for (int r = 0; r < ROUNDS; r += 2) {
        SDE_CHACHA_QR(x0, x4, x8, x12) SDE_CHACHA_QR(x1, x5, x9, x13)
        SDE_CHACHA_QR(x2, x6, x10, x14) SDE_CHACHA_QR(x3, x7, x11, x15)
        SDE_CHACHA_QR(x0, x5, x10, x15) SDE_CHACHA_QR(x1, x6, x11, x12)
        SDE_CHACHA_QR(x2, x7, x8, x13) SDE_CHACHA_QR(x3, x4, x9, x14)
    }
    out[0] = x0 + c0; out[1] = x1 + c1; out[2] = x2 + c2; out[3] = x3 + c3;
    out[4] = x4 + key[0]; out[5] = x5 + key[1]; out[6] = x6 + key[2]; out[7] = x7 + key[3];
    out[8] = x8 + key[4]; out[9] = x9 + key[5]; out[10] = x10 + key[6]; out[11] = x11 + key[7];
    out[12] = x12 + (sde_u32)counter; out[13] = x13 + (sde_u32)(counter >> 32); out[14] = x14; out[15] = x15;
}

// Per-path stream of rand's `random::<f64>()` draws: draw i uses words (2i, 2i+1) of block i/8.
// All indices into `buf` must be compile-time constants after unrolling (registers, not local memory).
struct SdeChaCha8Stream {
    sde_u32 key[8];
    sde_u32 buf[16];
    sde_u64 block;
    __device__ __forceinline__ void init(sde_u64 seed) { sde_seed_from_u64(seed, key); block = 0; }
    __device__ __forceinline__ void refill() { sde_chacha_block<8>(key, block, buf); ++block; }
    // 53-bit integer j of draw `slot` (0..7) in the current block: f64 = j * 2^-53 (rand: (next_u64 >> 11) * 2^-53)
    __device__ __forceinline__ sde_u64 bits53(int slot) const {
        return (((sde_u64)buf[2 * slot + 1] << 32) | (sde_u64)buf[2 * slot]) >> 11;
    }
};

// ---------------------------------------------------------------- Philox4x32-10 -------
// Counter-based generator of the generator = "philox" tier (Salmon et al., SC'11; Random123 / cuRAND Philox4_32_10): not in the
// reference — north_star asks the timed pseudo-random MC path only for STATISTICAL agreement with src/rng/pseudo.rs, and
// ChaCha8's 8 rounds over 16 words cost ~48 integer instructions per f64 draw on the half-rate ALU pipe (the bound of the
// terminal-only configs, profiles/r1_terminal_pipes.md).  One block = 4 32-bit draws, ~17 instructions per draw.
//   key = the two words of the seed; counter = (lo32(scenario), hi32(scenario), block, 0); draw #i of a path = word i & 3 of
//   block i >> 2; uniform = (word + 1/2) 2^-32 — the same 32-bit form as the digital-shift Sobol uniforms, so the integer
//   front end of the inverse normal applies unchanged.
__device__ __forceinline__ void sde_philox4x32_10(sde_u32 c0, sde_u32 c1, sde_u32 c2, sde_u32 c3, sde_u32 k0, sde_u32 k1, sde_u32 (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const sde_u32 h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const sde_u32 h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
struct SdePhiloxStream {
    sde_u32 k0, k1, s0, s1;
    sde_u32 buf[4];
    sde_u32 block;
    __device__ __forceinline__ void init(sde_u64 seed, sde_u64 scenario) {
        k0 = (sde_u32)seed; k1 = (sde_u32)(seed >> 32); s0 = (sde_u32)scenario; s1 = (sde_u32)(scenario >> 32); block = 0;
    }
    __device__ __forceinline__ void refill() { sde_philox4x32_10(s0, s1, block, 0u, k0, k1, buf); ++block; }
};

// ---------------------------------------------------------------- Sobol ---------------
// Direction numbers are stored as the top 32 bits of the 64-bit integers: exact for point
// indices n < 2^32 (such points only touch direction numbers 1..32, whose low 32 bits are 0).
// V[d][b], b = 0..31.

// x_d(n) from the raw direction numbers (used for the CTA base and by the standalone kernel).
__device__ __forceinline__ sde_u32 sde_sobol_point32(const sde_u32* __restrict__ Vd, sde_u32 n) {
    sde_u32 g = n ^ (n >> 1), x = 0;
    while (g) {
        int b = __ffs(g) - 1;
        x ^= __ldg(Vd + b);
        g &= g - 1;
    }
    return x;
}
