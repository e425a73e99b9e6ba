// Vector load / store of 8 consecutive channels of a (hi, lo) fp16 activation (shared by the HBM-bound kernels).
#pragma once
#include "common.cuh"

namespace fisr {

__device__ __forceinline__ void unpack8(const uint4& v, __half (&h)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[2 * i] = __ushort_as_half(static_cast<unsigned short>(w[i] & 0xFFFF));
        h[2 * i + 1] = __ushort_as_half(static_cast<unsigned short>(w[i] >> 16));
    }
}
// 8 consecutive channels of a (hi, lo) activation -> fp32
template <int PLANES>
__device__ __forceinline__ void load8(const __half* p, size_t plane, float (&f)[8]) {
    __half h[8], l[8];
    unpack8(*reinterpret_cast<const uint4*>(p), h);
    if (PLANES == 2) {
        unpack8(*reinterpret_cast<const uint4*>(p + plane), l);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = join_f16(h[i], l[i]);
    } else if (PLANES == 3) {
        const uint2 b = *reinterpret_cast<const uint2*>(f8_row_ptr(p + plane));      // 8 e5m2 bytes of 16 * lo
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = join_f8(h[i], static_cast<uint8_t>(((i < 4 ? b.x : b.y) >> (8 * (i & 3))) & 0xFF));
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __half2float(h[i]);
    }
}
template <int PLANES>
__device__ __forceinline__ void store8(__half* p, size_t plane, const float (&f)[8]) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split2_f32(f[2 * q], f[2 * q + 1], hi[q], lo[q]);
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (PLANES == 2) *reinterpret_cast<uint4*>(p + plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (PLANES == 3) {
        uint8_t* q = f8_row_ptr(p + plane);
        *reinterpret_cast<uint2*>(q) = make_uint2(f8_pack_lo4(lo[0], lo[1]), f8_pack_lo4(lo[2], lo[3]));
        *reinterpret_cast<uint2*>(q + 64) = make_uint2(f8_pack_hi4(hi[0], hi[1]), f8_pack_hi4(hi[2], hi[3]));
    }
}

}  // namespace fisr
