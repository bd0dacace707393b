// Register-resident sub-FFT engine (complex fp64) for the FFT power spectrum, v2.
//
// A sub-transform of length N = R1*R2 (R1, R2 <= 16) is carried out by a group of TG lanes of ONE
// warp.  Every lane holds up to 16 complex values in registers, runs compile-time unrolled radix-R
// codelets on them (PencilDft, twiddles are immediate constants) and meets the other lanes of its
// group exactly once, through a small padded shared-memory slab, synchronised with __syncwarp only.
// A 256-point transform therefore costs two 16-point register butterflies and ONE shared-memory
// exchange; the warps of a CTA stream through their tiles independently (no block barrier inside a
// pass), which is what hides the L2 / HBM latency of the four-step passes built on top of this.
//
// Index algebra (j = j1*R2 + j2, k = k1 + R1*k2):
//   DIT  X[k1 + R1 k2] = sum_j2 w_R2^(j2 k2) [ w_N^(j2 k1) sum_j1 x[j1 R2 + j2] w_R1^(j1 k1) ]
//        in : lane g, register i*R1 + a   holds x[a*R2 + (g + TG*i)]       ("natural" layout)
//        out: lane g, register i'*R2 + k2 holds X[(g + TG*i') + R1*k2]     ("digit" layout)
//   DIF  the mirror image: digit layout in, natural layout out, same forward roots -- so that
//        FFT -> pointwise op -> FFT needs no reordering in between.
#pragma once

#include "dp_pencil.cuh"

namespace dp {

template <int R1_, int R2_, int TG_> struct SubCfg {
    static constexpr int R1 = R1_, R2 = R2_, TG = TG_, N = R1_ * R2_;
    static constexpr int NB1 = (R2 + TG - 1) / TG;      // radix-R1 butterflies per lane (stage 1)
    static constexpr int NB2 = (R1 + TG - 1) / TG;      // radix-R2 butterflies per lane (stage 2)
    static constexpr int NS = 32 / TG;                  // sequences per warp
    static constexpr int PITCH = (R2 % 2 == 0) ? R2 + 1 : R2;          // odd row pitch: no bank conflicts
    static constexpr int ESEQ0 = R1 * PITCH;
    // Lane l of a warp works on sequence l % NS as group lane g = l / NS (sequence index fastest, so that
    // the global accesses of a quarter-warp fall into few 128-byte lines).  A quarter-warp
    // shared-memory phase then holds min(NS, 8) sequences x 8 / min(NS, 8) consecutive g: the slab stride
    // between sequences is chosen = 8 / min(NS, 8) (mod 8) to spread them over the 16-byte bank groups.
    static constexpr int NSQ = NS < 8 ? NS : 8;
    static constexpr int ESEQ = ESEQ0 + ((8 / NSQ - ESEQ0 % 8) + 8) % 8;
    static constexpr int EWARP = NS * ESEQ;             // complex elements of slab per warp
    static constexpr bool FULL1 = (R2 % TG == 0);       // every lane of the group works in stage 1
    static constexpr bool FULL2 = (R1 % TG == 0);
    static_assert(NB1 * R1 <= 16 && NB2 * R2 <= 16, "a lane holds at most 16 points");
    static_assert(32 % TG == 0, "group size must divide the warp");
};

constexpr int SUBFFT_EWARP_MAX = 640;     // >= EWARP of every configuration instantiated below

// ---- lean in-place power-of-two codelets ---------------------------------------------------------------
// The generic PencilDft codelets are out of place (input + sub-transforms + output live at once: 3N complex
// values); for the 16-point butterflies of the spectrum engine that costs spills at 255 registers.  These work in
// place on the N values plus two temporaries; the digit-reversed result order is undone by a compile-time
// renaming of the registers (no instructions).
__device__ __forceinline__ void r4_inplace(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {      // forward: w_4 = -i
    const cplx t0 = cadd(a0, a2), t1 = csub(a0, a2);
    const cplx t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = cadd(t1, t3); a3 = csub(t1, t3);
}
__device__ __forceinline__ cplx mul_w8_1(cplx a) {      // a * exp(-i pi/4) = ((x + y) + i (y - x)) / sqrt(2)
    constexpr double h = 0.70710678118654752440;
    return make_double2((a.x + a.y) * h, (a.y - a.x) * h);
}
__device__ __forceinline__ cplx mul_w8_3(cplx a) {      // a * exp(-3 i pi/4) = ((y - x) - i (x + y)) / sqrt(2)
    constexpr double h = 0.70710678118654752440;
    return make_double2((a.y - a.x) * h, -(a.x + a.y) * h);
}
__device__ __forceinline__ void dft8_lean(cplx* v) {
    // 8 = 2 x 4 (DIT): DFT4 of the even and of the odd samples, twiddle w_8^k, radix-2
    r4_inplace(v[0], v[2], v[4], v[6]);
    r4_inplace(v[1], v[3], v[5], v[7]);
    // position 2k holds E[k], position 2k+1 holds O[k]
    v[3] = mul_w8_1(v[3]);
    v[5] = cmul_mi(v[5]);
    v[7] = mul_w8_3(v[7]);
    cplx t;
    t = v[0]; v[0] = cadd(t, v[1]); v[1] = csub(t, v[1]);      // X0, X4
    t = v[2]; v[2] = cadd(t, v[3]); v[3] = csub(t, v[3]);      // X1, X5
    t = v[4]; v[4] = cadd(t, v[5]); v[5] = csub(t, v[5]);      // X2, X6
    t = v[6]; v[6] = cadd(t, v[7]); v[7] = csub(t, v[7]);      // X3, X7
    // rename: position 2k -> X[k], position 2k+1 -> X[k+4]
    const cplx x1 = v[2], x2 = v[4], x3 = v[6], x4 = v[1], x5 = v[3], x6 = v[5];
    v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6;
}
__device__ __forceinline__ void dft16_lean(cplx* v) {
    // 16 = 4 x 4 (DIT): S_r = DFT4(v[r], v[r+4], v[r+8], v[r+12]) in place (position r + 4k holds S_r[k])
    constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;     // cos, sin (pi/8)
#pragma unroll
    for (int r = 0; r < 4; r++) r4_inplace(v[r], v[r + 4], v[r + 8], v[r + 12]);
    // twiddles w_16^(r k)
    v[5] = cmul(v[5], make_double2(c1, -s1));       // r=1 k=1
    v[9] = mul_w8_1(v[9]);                          // r=1 k=2: w16^2
    v[13] = cmul(v[13], make_double2(s1, -c1));     // r=1 k=3: w16^3
    v[6] = mul_w8_1(v[6]);                          // r=2 k=1: w16^2
    v[10] = cmul_mi(v[10]);                         // r=2 k=2: w16^4
    v[14] = mul_w8_3(v[14]);                        // r=2 k=3: w16^6
    v[7] = cmul(v[7], make_double2(s1, -c1));       // r=3 k=1: w16^3
    v[11] = mul_w8_3(v[11]);                        // r=3 k=2: w16^6
    v[15] = cmul(v[15], make_double2(-c1, s1));     // r=3 k=3: w16^9
    // X[k + 4u] = sum_r w_4^(r u) (twiddled S_r[k]): radix-4 across r on positions 4k + r, result u at 4k + u
#pragma unroll
    for (int k = 0; k < 4; k++) r4_inplace(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    // rename (transpose of the 4 x 4 index grid): position 4k + u -> X[k + 4u]
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int u = k + 1; u < 4; u++) {
            const cplx t = v[4 * k + u];
            v[4 * k + u] = v[4 * u + k];
            v[4 * u + k] = t;
        }
}

template <int R> __device__ __forceinline__ void dft_inplace(cplx* v) {
#ifndef DP_NO_LEAN_CODELETS
    if constexpr (R == 16) { dft16_lean(v); return; }
    else if constexpr (R == 8) { dft8_lean(v); return; }
    else if constexpr (R == 4) { r4_inplace(v[0], v[1], v[2], v[3]); return; }
    else
#endif
    if constexpr (R > 1) {
        cplx o[R];
        PencilDft<R>::run(v, o);
#pragma unroll
        for (int u = 0; u < R; u++) v[u] = o[u];
    }
}

// tab[k1*R2 + j2] = w_N^(j2 k1) (forward roots): consecutive lanes read consecutive entries.
// sub_dit_a: first butterflies + twiddles, values parked in the slab (the registers are free afterwards:
// the place to issue the loads of the next tile); sub_dit_b: exchange + second butterflies.
template <class C>
__device__ __forceinline__ void sub_dit_a(cplx (&v)[16], cplx* __restrict__ E, const cplx* __restrict__ tab, int g) {
#pragma unroll
    for (int i = 0; i < C::NB1; i++) {
        const int j2 = g + C::TG * i;
        if (C::FULL1 || j2 < C::R2) {
            dft_inplace<C::R1>(&v[i * C::R1]);
            if constexpr (C::R2 > 1) {
                E[j2] = v[i * C::R1];
#pragma unroll
                for (int k1 = 1; k1 < C::R1; k1++)
                    E[k1 * C::PITCH + j2] = cmul(v[i * C::R1 + k1], tab[k1 * C::R2 + j2]);
            }
        }
    }
}
template <class C>
__device__ __forceinline__ void sub_dit_b(cplx (&v)[16], const cplx* __restrict__ E, int g) {
    if constexpr (C::R2 > 1) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < C::NB2; i++) {
            const int k1 = g + C::TG * i;
            if (C::FULL2 || k1 < C::R1) {
#pragma unroll
                for (int b = 0; b < C::R2; b++) v[i * C::R2 + b] = E[k1 * C::PITCH + b];
                dft_inplace<C::R2>(&v[i * C::R2]);
            }
        }
    }
}
template <class C>
__device__ __forceinline__ void sub_dit(cplx (&v)[16], cplx* __restrict__ E, const cplx* __restrict__ tab, int g) {
    sub_dit_a<C>(v, E, tab, g);
    sub_dit_b<C>(v, E, g);
}

// tab[c*R1 + k1] = w_N^(k1 c)
template <class C>
__device__ __forceinline__ void sub_dif(cplx (&v)[16], cplx* __restrict__ E, const cplx* __restrict__ tab, int g) {
    if constexpr (C::R2 > 1) {
#pragma unroll
        for (int i = 0; i < C::NB2; i++) {
            const int k1 = g + C::TG * i;
            if (C::FULL2 || k1 < C::R1) {
                dft_inplace<C::R2>(&v[i * C::R2]);
                E[k1 * C::PITCH] = v[i * C::R2];
#pragma unroll
                for (int c = 1; c < C::R2; c++)
                    E[k1 * C::PITCH + c] = cmul(v[i * C::R2 + c], tab[c * C::R1 + k1]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < C::NB1; i++) {
            const int c = g + C::TG * i;
            if (C::FULL1 || c < C::R2) {
#pragma unroll
                for (int k1 = 0; k1 < C::R1; k1++) v[i * C::R1 + k1] = E[k1 * C::PITCH + c];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < C::NB1; i++) {
        const int c = g + C::TG * i;
        if (C::FULL1 || c < C::R2) dft_inplace<C::R1>(&v[i * C::R1]);
    }
}

// w_P^idx for idx < P <= 65536 from a fine (w_P^j, j < 256) and a coarse (w_P^(256 j)) table
__device__ __forceinline__ cplx tw2level(const cplx* __restrict__ fine, const cplx* __restrict__ coarse, int idx) {
    const cplx f = fine[idx & 255];
    const int hi = idx >> 8;
    return hi ? cmul(coarse[hi], f) : f;
}

}  // namespace dp
