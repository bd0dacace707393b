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

template <int R> __device__ __forceinline__ void dft_inplace(cplx* v) {
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
