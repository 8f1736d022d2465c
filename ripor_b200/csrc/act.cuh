// Activation operand buffers ("ActBuf"): what every GEMM of the engine reads as its A operand.
//
// The tensor-core GEMMs run on error-compensated splits of the fp32 activations, so producers (RMSNorm,
// attention, the ReLU epilogue) write the operand already split into planes:
//   RB200_PREC_FP32    plane 0 = raw fp32
//   RB200_PREC_TF32X3  plane 0 = tf32(x), plane 1 = tf32(x - plane0), both stored as fp32 words
//   RB200_PREC_BF16X3  plane 0 = bf16(x), plane 1 = bf16(x - plane0)
//   RB200_PREC_TF32    plane 0 = tf32(x)
//   RB200_PREC_BF16    plane 0 = bf16(x)
//   RB200_PREC_FP16X3  plane 0 = fp16(x), plane 1 = fp16(x - plane0): the 11-bit mantissa of tf32 at half the
//                      bytes and twice the MMA rate; |x| > 65504 cannot be represented, so every store checks
//                      the range and raises the engine's overflow flag (the search then returns NaN scores
//                      instead of silently wrong DocIDs)
// `plane` is the element distance between the planes (row capacity * row length).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstring>

namespace rb {

struct ActOut {
  void* base = nullptr;
  int64_t plane = 0;        // elements between plane 0 and plane 1
  int mode = 0;             // rb200_precision
  int* overflow = nullptr;  // set to 1 when an fp16 plane cannot hold a value
};

__host__ __device__ inline int prec_planes(int mode) { return (mode == 1 || mode == 2 || mode == 5) ? 2 : 1; }
__host__ __device__ inline int prec_elem_bytes(int mode) { return (mode == 2 || mode == 4 || mode == 5) ? 2 : 4; }
__host__ __device__ inline bool prec_is_fp16(int mode) { return mode == 5; }

constexpr float kFp16Limit = 65000.0f;

// round-to-nearest-even to the 10-bit tf32 mantissa, kept as an fp32 word with the low 13 bits cleared
__host__ __device__ inline float round_tf32(float x) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(x);
#else
  uint32_t b;
  memcpy(&b, &x, 4);
#endif
  if ((b & 0x7f800000u) != 0x7f800000u) b = (b + 0xfffu + ((b >> 13) & 1u)) & 0xffffe000u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float r;
  memcpy(&r, &b, 4);
  return r;
#endif
}

__device__ __forceinline__ void act_store(const ActOut& o, int64_t idx, float v) {
  switch (o.mode) {
    case 0:
      static_cast<float*>(o.base)[idx] = v;
      break;
    case 1: {
      const float hi = round_tf32(v);
      static_cast<float*>(o.base)[idx] = hi;
      static_cast<float*>(o.base)[o.plane + idx] = round_tf32(v - hi);
      break;
    }
    case 2: {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      static_cast<__nv_bfloat16*>(o.base)[idx] = hi;
      static_cast<__nv_bfloat16*>(o.base)[o.plane + idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
      break;
    }
    case 3:
      static_cast<float*>(o.base)[idx] = round_tf32(v);
      break;
    case 5: {
      if (!(fabsf(v) <= kFp16Limit) && o.overflow) *o.overflow = 1;
      const __half hi = __float2half_rn(v);
      static_cast<__half*>(o.base)[idx] = hi;
      static_cast<__half*>(o.base)[o.plane + idx] = __float2half_rn(v - __half2float(hi));
      break;
    }
    default:
      static_cast<__nv_bfloat16*>(o.base)[idx] = __float2bfloat16_rn(v);
      break;
  }
}

// two consecutive elements, idx % 2 == 0
__device__ __forceinline__ void act_store2(const ActOut& o, int64_t idx, float2 v) {
  switch (o.mode) {
    case 0:
      *reinterpret_cast<float2*>(static_cast<float*>(o.base) + idx) = v;
      break;
    case 1: {
      const float2 hi = make_float2(round_tf32(v.x), round_tf32(v.y));
      *reinterpret_cast<float2*>(static_cast<float*>(o.base) + idx) = hi;
      *reinterpret_cast<float2*>(static_cast<float*>(o.base) + o.plane + idx) =
          make_float2(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y));
      break;
    }
    case 3:
      *reinterpret_cast<float2*>(static_cast<float*>(o.base) + idx) = make_float2(round_tf32(v.x), round_tf32(v.y));
      break;
    case 5: {
      if (!(fabsf(v.x) <= kFp16Limit && fabsf(v.y) <= kFp16Limit) && o.overflow) *o.overflow = 1;
      __half* b = static_cast<__half*>(o.base);
      const __half2 hi = __floats2half2_rn(v.x, v.y);
      const float2 hf = __half22float2(hi);
      *reinterpret_cast<__half2*>(b + idx) = hi;
      *reinterpret_cast<__half2*>(b + o.plane + idx) = __floats2half2_rn(v.x - hf.x, v.y - hf.y);
      break;
    }
    default: {   // bf16 planes
      __nv_bfloat16* b = static_cast<__nv_bfloat16*>(o.base);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(v.x, v.y);
      *reinterpret_cast<__nv_bfloat162*>(b + idx) = hi;
      if (o.mode == 2) {
        const float2 hf = __bfloat1622float2(hi);
        *reinterpret_cast<__nv_bfloat162*>(b + o.plane + idx) = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
      }
      break;
    }
  }
}

// four consecutive elements, idx % 4 == 0 (vector stores)
__device__ __forceinline__ void act_store4(const ActOut& o, int64_t idx, float4 v) {
  switch (o.mode) {
    case 0:
      *reinterpret_cast<float4*>(static_cast<float*>(o.base) + idx) = v;
      break;
    case 1: {
      float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
      float4 lo = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z),
                              round_tf32(v.w - hi.w));
      *reinterpret_cast<float4*>(static_cast<float*>(o.base) + idx) = hi;
      *reinterpret_cast<float4*>(static_cast<float*>(o.base) + o.plane + idx) = lo;
      break;
    }
    case 3: {
      float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
      *reinterpret_cast<float4*>(static_cast<float*>(o.base) + idx) = hi;
      break;
    }
    case 5: {
      __half* b = static_cast<__half*>(o.base);
      const float f[4] = {v.x, v.y, v.z, v.w};
      __half hi[4], lo[4];
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bad |= !(fabsf(f[i]) <= kFp16Limit);
        hi[i] = __float2half_rn(f[i]);
        lo[i] = __float2half_rn(f[i] - __half2float(hi[i]));
      }
      if (bad && o.overflow) *o.overflow = 1;
      *reinterpret_cast<uint2*>(b + idx) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(b + o.plane + idx) = *reinterpret_cast<uint2*>(lo);
      break;
    }
    default: {   // bf16 planes
      __nv_bfloat16* b = static_cast<__nv_bfloat16*>(o.base);
      const float f[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hi[i] = __float2bfloat16_rn(f[i]);
        lo[i] = __float2bfloat16_rn(f[i] - __bfloat162float(hi[i]));
      }
      *reinterpret_cast<uint2*>(b + idx) = *reinterpret_cast<uint2*>(hi);
      if (o.mode == 2) *reinterpret_cast<uint2*>(b + o.plane + idx) = *reinterpret_cast<uint2*>(lo);
      break;
    }
  }
}

}  // namespace rb
