// Lytro raw bit-unpacking ("next" row) — replaces the strided numpy shuffles of LfpDecoder.comp_bayer
// (lfp_reader/lfp_decoder.py:273-326):
//   10 bit: 5 bytes -> 4 pixels,  p_k = b_k << 2 | (b_4 >> 2k) & 3
//   12 bit: 3 bytes -> 2 pixels,  p_0 = b_0 << 4 | b_1 >> 4,  p_1 = (b_1 & 15) << 8 | b_2
// (the reference's [0::4] / [2::4] / [1::2] interleave of the 12-bit halves is the plain sequence p_0, p_1).
// Pure streaming: 1.25 / 1.5 bytes in, 2 bytes out per pixel.
#pragma once
#include "common.cuh"

namespace pcb {

__global__ void __launch_bounds__(256)
unpack10_kernel(const uint8_t *__restrict__ buf, int64_t ngroups, uint16_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const uint8_t *b = buf + g * 5;
        const uint32_t b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2), b3 = __ldg(b + 3), b4 = __ldg(b + 4);
        ushort4 p;
        p.x = (uint16_t)((b0 << 2) | (b4 & 3u));
        p.y = (uint16_t)((b1 << 2) | ((b4 >> 2) & 3u));
        p.z = (uint16_t)((b2 << 2) | ((b4 >> 4) & 3u));
        p.w = (uint16_t)((b3 << 2) | ((b4 >> 6) & 3u));
        reinterpret_cast<ushort4 *>(out)[g] = p;          // out is 8-byte aligned (checked by the caller)
    }
}

__global__ void __launch_bounds__(256)
unpack12_kernel(const uint8_t *__restrict__ buf, int64_t ngroups, uint16_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const uint8_t *b = buf + g * 3;
        const uint32_t b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
        ushort2 p;
        p.x = (uint16_t)((b0 << 4) | (b1 >> 4));
        p.y = (uint16_t)(((b1 & 15u) << 8) | b2);
        reinterpret_cast<ushort2 *>(out)[g] = p;
    }
}

}  // namespace pcb
