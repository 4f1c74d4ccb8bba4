// Small vector / quaternion helpers for the step kernel (row-major 3x3, quaternion w,x,y,z).
#pragma once
#include <cuda_runtime.h>

#ifndef MW_REAL_DOUBLE
typedef float real;
#define R_(x) x##f
#define MW_EPS 1e-12f
#else
typedef double real;
#define R_(x) x
#define MW_EPS 1e-15
#endif

#define DEV __device__ __forceinline__
#define FULLMASK 0xffffffffu

DEV real rsqrt_(real x) { return (real)1 / sqrt(x); }
DEV void v3zero(real* a) { a[0] = a[1] = a[2] = 0; }
DEV void v3copy(real* a, const real* b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
DEV void v3add(real* r, const real* a, const real* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
DEV void v3sub(real* r, const real* a, const real* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
DEV void v3scl(real* r, const real* a, real s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
DEV void v3addscl(real* r, const real* a, const real* b, real s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
DEV real v3dot(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DEV real v3norm(const real* a) { return sqrt(v3dot(a, a)); }
DEV void v3cross(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
DEV real v3normalize(real* a) {
  real n = v3norm(a);
  if (n < MW_EPS) { a[0] = 1; a[1] = a[2] = 0; return 0; }
  real s = (real)1 / n; a[0] *= s; a[1] *= s; a[2] *= s; return n;
}
DEV void mat_mulvec(real* r, const real* M, const real* v) {
  real x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2], z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
DEV void mat_tmulvec(real* r, const real* M, const real* v) {
  real x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2], z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
DEV void mat_col(real* r, const real* M, int c) { r[0] = M[c]; r[1] = M[3 + c]; r[2] = M[6 + c]; }
DEV void mat_mul(real* r, const real* A, const real* B) {
  real t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) r[i] = t[i];
}
DEV void quat_mul(real* r, const real* a, const real* b) {
  real w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
DEV void quat_normalize(real* q) {
  real n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MW_EPS) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  real s = (real)1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
DEV void quat2mat(real* R, const real* q) {
  real w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
DEV void quat_rot(real* r, const real* q, const real* v) { real R[9]; quat2mat(R, q); mat_mulvec(r, R, v); }
DEV void quat_axisangle(real* q, const real* axis, real ang) {
  real s, c; sincos((real)0.5 * ang, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
DEV real warp_sum(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
DEV real bcast(real v, int src) { return __shfl_sync(FULLMASK, v, src); }
