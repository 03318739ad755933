#pragma once
// Component-wise arithmetic on the 16-byte CUDA vector types, the device
// header shared by the 1-D kernels (counterpart of the reference's
// utils/cuda_vectors.h: +=, -=, *=, +, * for float4/double2 against a vector
// or a scalar, plus the `cg` namespace alias).  Generated from two macros
// instead of written out per operator.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define B200FE_VEC_APPLY2(OP, a, b) { (a).x OP (b).x; (a).y OP (b).y; }
#define B200FE_VEC_APPLY4(OP, a, b) { (a).x OP (b).x; (a).y OP (b).y; (a).z OP (b).z; (a).w OP (b).w; }
#define B200FE_SCL_APPLY2(OP, a, s) { (a).x OP (s); (a).y OP (s); }
#define B200FE_SCL_APPLY4(OP, a, s) { (a).x OP (s); (a).y OP (s); (a).z OP (s); (a).w OP (s); }

#define B200FE_VEC_COMPOUND(OP)                                                                              \
    __host__ __device__ inline double2 &operator OP(double2 &a, const double2 b) { B200FE_VEC_APPLY2(OP, a, b) return a; } \
    __host__ __device__ inline float4 &operator OP(float4 &a, const float4 b) { B200FE_VEC_APPLY4(OP, a, b) return a; }    \
    __host__ __device__ inline double2 &operator OP(double2 &a, const double s) { B200FE_SCL_APPLY2(OP, a, s) return a; }  \
    __host__ __device__ inline float4 &operator OP(float4 &a, const float s) { B200FE_SCL_APPLY4(OP, a, s) return a; }

B200FE_VEC_COMPOUND(+=)
B200FE_VEC_COMPOUND(-=)
B200FE_VEC_COMPOUND(*=)

#define B200FE_VEC_BINARY(OP, COMPOUND)                                                                      \
    __host__ __device__ inline double2 operator OP(double2 a, const double2 b) { return a COMPOUND b; }     \
    __host__ __device__ inline float4 operator OP(float4 a, const float4 b) { return a COMPOUND b; }        \
    __host__ __device__ inline double2 operator OP(double2 a, const double s) { return a COMPOUND s; }      \
    __host__ __device__ inline float4 operator OP(float4 a, const float s) { return a COMPOUND s; }

B200FE_VEC_BINARY(+, +=)
B200FE_VEC_BINARY(-, -=)
B200FE_VEC_BINARY(*, *=)

#undef B200FE_VEC_BINARY
#undef B200FE_VEC_COMPOUND
#undef B200FE_SCL_APPLY4
#undef B200FE_SCL_APPLY2
#undef B200FE_VEC_APPLY4
#undef B200FE_VEC_APPLY2
